set -x
timeout 900 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_keyframe.py -x -q -m gpu > gpurun_out/r2t_tests.log 2>&1; pe=$?; tail -3 gpurun_out/r2t_tests.log
if [ $pe -eq 0 ]; then
  for q in 1 8; do timeout 300 python bench.py --workload retrieval --queries $q --steps 100 > gpurun_out/r2t_retrieval_q$q.json 2> gpurun_out/r2t_retrieval_q$q.err; echo "retrieval q=$q rc $?"; done
  timeout 300 python bench.py --workload keyframe --steps 20 > gpurun_out/r2t_keyframe.json 2> gpurun_out/r2t_keyframe.err; echo "keyframe rc $?"; tail -3 gpurun_out/r2t_keyframe.err
  python -c "
import json
for q in (1,8):
    r=json.loads(open('gpurun_out/r2t_retrieval_q%d.json'%q).read().strip().splitlines()[-1]); print('retr',q,r['ms_per_step'],r['config']['ms_per_query'],r['roofline']['frac'])
r=json.loads(open('gpurun_out/r2t_keyframe.json').read().strip().splitlines()[-1]); print('kf',r['value'],r['ms_per_step'],r['e2e']['value'],r.get('cpu_baseline'),r.get('checks'))"
fi
