set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2m_gpu_tests.log 2>&1; pe=$?; tail -3 gpurun_out/r2m_gpu_tests.log
if [ $pe -eq 0 ]; then
  timeout 600 python tools/peer_store_cost.py 2>&1 | tee gpurun_out/r2m_peer_store_cost.txt
  timeout 600 python tools/ab.py --tag r2m_hdl64 --repeats 2 bulk: old:tune:NSC_WS=0 2>&1 | tee gpurun_out/r2m_ab_hdl64.txt
  timeout 600 python tools/ab.py --tag r2m_hdl32 --repeats 1 --args "--shape hdl32 --scans 4096" bulk: 2>&1 | tee gpurun_out/r2m_ab_hdl32.txt
fi
