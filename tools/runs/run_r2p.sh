set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2p_gpu_tests.log 2>&1; pe=$?; tail -3 gpurun_out/r2p_gpu_tests.log
if [ $pe -eq 0 ]; then
  timeout 900 python tools/ab.py --tag r2p_hdl64 --repeats 2 hint: nohint:nohint old:tune:NSC_WS=0 2>&1 | tee gpurun_out/r2p_ab_hdl64.txt
  timeout 600 python tools/ab.py --tag r2p_hdl32 --repeats 2 --args "--shape hdl32 --scans 4096" hint: nohint:nohint 2>&1 | tee gpurun_out/r2p_ab_hdl32.txt
  for q in 1 8; do timeout 300 python bench.py --workload retrieval --queries $q --steps 50 > gpurun_out/r2p_retrieval_q$q.json 2> gpurun_out/r2p_retrieval_q$q.err; echo "retrieval q=$q rc $?"; done
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err; echo "bench rc $?"; tail -2 gpurun_out/r2p_bench_n1.err
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extras"
  $CMD > gpurun_out/r2p_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nsc -c 200 --csv --log-file gpurun_out/r2p_launches.csv $CMD > gpurun_out/r2p_ncu1.log 2>&1; echo "ncu launches rc $?"
  $CMD > gpurun_out/r2p_plain.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:encode_points_ws -s 3 -c 1 -o gpurun_out/r2p_ws $CMD > gpurun_out/r2p_ncu2.log 2>&1; echo "ncu full rc $?"
fi
