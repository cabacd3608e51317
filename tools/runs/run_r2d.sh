set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2d_parity.log 2>&1; pe=$?; tail -5 gpurun_out/r2d_parity.log
if [ $pe -eq 0 ]; then
  timeout 1500 python tools/ab.py --tag r2d_hdl64 --repeats 2 tma: lws:lws old:tune:NSC_WS=0 d4:d4 d6:d6 s23:s23 s25:s25 xnocomp:xnocomp xskiptail:xskiptail 2>&1 | tee gpurun_out/r2d_ab_hdl64.txt
  timeout 1200 python tools/ab.py --tag r2d_hdl32 --repeats 2 --args "--shape hdl32 --scans 4096" tma: lws:lws old:tune:NSC_WS=0 d4:d4 d6:d6 s23:s23 s25:s25 xskiptail:xskiptail 2>&1 | tee gpurun_out/r2d_ab_hdl32.txt
  timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2d_gpu_tests.log 2>&1; tail -5 gpurun_out/r2d_gpu_tests.log
fi
