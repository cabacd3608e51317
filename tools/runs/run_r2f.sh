set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2f_parity.log 2>&1; pe=$?; tail -3 gpurun_out/r2f_parity.log
if [ $pe -eq 0 ]; then
  timeout 1800 python tools/ab.py --tag r2f_hdl64 --repeats 2 d6: d5:d5 p1d12:p1d12 p1d10:p1d10 p3d4:p3d4 s26:s26 s22:s22 old:tune:NSC_WS=0 2>&1 | tee gpurun_out/r2f_ab_hdl64.txt
  timeout 1200 python tools/ab.py --tag r2f_hdl32 --repeats 2 --args "--shape hdl32 --scans 4096" d6: d5:d5 p1d12:p1d12 p1d10:p1d10 p3d4:p3d4 s26:s26 s22:s22 old:tune:NSC_WS=0 2>&1 | tee gpurun_out/r2f_ab_hdl32.txt
fi
