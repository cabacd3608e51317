set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2b_parity.log 2>&1; pe=$?; tail -5 gpurun_out/r2b_parity.log
if [ $pe -eq 0 ]; then
  timeout 120 tools/_build/membw 2>&1 | tee gpurun_out/r2b_membw.txt
  timeout 1500 python tools/ab.py --tag r2b_hdl64 --repeats 2 ws: old:tune:NSC_WS=0 ef:ef d4:d4 xnocomp:xnocomp xnocompd6:xnocompd6 xskiptail:xskiptail 2>&1 | tee gpurun_out/r2b_ab_hdl64.txt
  timeout 1200 python tools/ab.py --tag r2b_hdl32 --repeats 2 --args "--shape hdl32 --scans 4096" ws: old:tune:NSC_WS=0 ef:ef d4:d4 xnocomp:xnocomp xskiptail:xskiptail 2>&1 | tee gpurun_out/r2b_ab_hdl32.txt
  timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench rc $?"; tail -3 gpurun_out/r2b_bench_n1.err
  timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; echo "ref rc $?"
fi
