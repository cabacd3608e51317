set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_gpu_tests.log 2>&1; pe=$?; tail -5 gpurun_out/r2h_gpu_tests.log
if [ $pe -eq 0 ]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; echo "bench rc $?"; tail -2 gpurun_out/r2h_bench_n1.err
  timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err; echo "ref rc $?"
  for q in 1 8; do timeout 300 python bench.py --workload retrieval --queries $q --steps 50 > gpurun_out/r2h_retrieval_q$q.json 2> gpurun_out/r2h_retrieval_q$q.err; echo "retrieval q=$q rc $?"; done
  timeout 300 python tools/latency.py > gpurun_out/r2h_latency.txt 2>&1; tail -5 gpurun_out/r2h_latency.txt
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extras"
  $CMD > gpurun_out/r2h_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches.csv $CMD > gpurun_out/r2h_ncu1.log 2>&1; echo "ncu launches rc $?"
  $CMD > gpurun_out/r2h_plain.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:encode_points_ws -s 3 -c 1 -o gpurun_out/r2h_ws $CMD > gpurun_out/r2h_ncu2.log 2>&1; echo "ncu full rc $?"
  RCMD="python bench.py --workload retrieval --queries 1 --steps 3 --no-cpu"
  $RCMD > gpurun_out/r2h_plain_r.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2h_launches_retrieval.csv $RCMD > gpurun_out/r2h_ncu3.log 2>&1; echo "ncu retrieval rc $?"
fi
