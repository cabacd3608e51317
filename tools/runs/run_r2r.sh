set -x
timeout 900 python -m pytest tests/test_gpu_retrieval.py -x -q -m gpu > gpurun_out/r2r_retrieval_tests.log 2>&1; pe=$?; tail -5 gpurun_out/r2r_retrieval_tests.log
if [ $pe -eq 0 ]; then
  for q in 1 8; do timeout 300 python bench.py --workload retrieval --queries $q --steps 100 > gpurun_out/r2r_retrieval_q$q.json 2> gpurun_out/r2r_retrieval_q$q.err; echo "retrieval q=$q rc $?"; done
  python -c "
import json
for q in (1,8):
    r=json.loads(open('gpurun_out/r2r_retrieval_q%d.json'%q).read().strip().splitlines()[-1]); print('retr',q,r['ms_per_step'],r['config']['ms_per_query'],r['roofline']['frac'])"
  RCMD="python bench.py --workload retrieval --queries 1 --steps 3 --no-cpu"
  $RCMD > gpurun_out/r2r_plain_r.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wasserstein\|select\|topk\|cdf_rows -c 100 --csv --log-file gpurun_out/r2r_launches_retrieval.csv $RCMD > gpurun_out/r2r_ncu.log 2>&1; echo "ncu rc $?"
fi
