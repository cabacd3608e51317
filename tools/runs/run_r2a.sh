set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2a_parity.log 2>&1; pe=$?; tail -5 gpurun_out/r2a_parity.log
if [ $pe -eq 0 ]; then
  timeout 1200 python tools/ab.py --tag r2a_hdl64 --repeats 2 ws: old:tune:NSC_WS=0 d4:d4 d6:d6 s28:s28 s20:s20 2>&1 | tee gpurun_out/r2a_ab_hdl64.txt
  timeout 900 python tools/ab.py --tag r2a_hdl32 --repeats 2 --args "--shape hdl32 --scans 4096" ws: old:tune:NSC_WS=0 d4:d4 d6:d6 s28:s28 s20:s20 2>&1 | tee gpurun_out/r2a_ab_hdl32.txt
  timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_gpu_tests.log 2>&1; tail -5 gpurun_out/r2a_gpu_tests.log
fi
