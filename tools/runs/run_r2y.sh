set -x
timeout 900 python -m pytest tests/test_gpu_retrieval.py -x -q -m gpu > gpurun_out/r2y_tests.log 2>&1; pe=$?; tail -3 gpurun_out/r2y_tests.log
if [ $pe -eq 0 ]; then
  for q in 1 8; do timeout 300 python bench.py --workload retrieval --queries $q --steps 200 > gpurun_out/r2y_retrieval_q$q.json 2> gpurun_out/r2y_retrieval_q$q.err; echo "retrieval q=$q rc $?"; done
  python -c "
import json
for q in (1,8):
    r=json.loads(open('gpurun_out/r2y_retrieval_q%d.json'%q).read().strip().splitlines()[-1]); print('retr',q,r['ms_per_step'],r['config']['ms_per_query'],r['roofline']['frac'])"
fi
