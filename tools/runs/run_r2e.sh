set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2e_parity.log 2>&1; pe=$?; tail -3 gpurun_out/r2e_parity.log
if [ $pe -eq 0 ]; then
  timeout 1800 python tools/ab.py --tag r2e_hdl64 --repeats 2 pk: nopk:nopk d6:d6 old:tune:NSC_WS=0 xskiptail:xskiptail xt1:xt1 xt2:xt2 xt4:xt4 xt8:xt8 xt14:xt14 xt7:xt7 2>&1 | tee gpurun_out/r2e_ab_hdl64.txt
  timeout 1200 python tools/ab.py --tag r2e_hdl32 --repeats 1 --args "--shape hdl32 --scans 4096" pk: nopk:nopk d6:d6 old:tune:NSC_WS=0 xskiptail:xskiptail xt1:xt1 xt2:xt2 xt4:xt4 xt8:xt8 2>&1 | tee gpurun_out/r2e_ab_hdl32.txt
fi
