set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2s_gpu_tests.log 2>&1; pe=$?; tail -3 gpurun_out/r2s_gpu_tests.log
if [ $pe -eq 0 ]; then
  timeout 900 python tools/ab.py --tag r2s_hdl64 --repeats 2 split: nosplit:tune:NSC_TAILSPLIT=0 2>&1 | tee gpurun_out/r2s_ab_hdl64.txt
  timeout 600 python tools/ab.py --tag r2s_hdl32 --repeats 2 --args "--shape hdl32 --scans 4096" split: nosplit:tune:NSC_TAILSPLIT=0 2>&1 | tee gpurun_out/r2s_ab_hdl32.txt
  timeout 600 python tools/ab.py --tag r2s_b128 --repeats 2 --args "--shape beam128 --scans 2048" split: nosplit:tune:NSC_TAILSPLIT=0 2>&1 | tee gpurun_out/r2s_ab_b128.txt
  timeout 600 python tools/ab.py --tag r2s_hdl64_600 --repeats 2 --args "--scans 600" split: nosplit:tune:NSC_TAILSPLIT=0 2>&1 | tee gpurun_out/r2s_ab_hdl64_600.txt
  for q in 1 8; do timeout 300 python bench.py --workload retrieval --queries $q --steps 100 > gpurun_out/r2s_retrieval_q$q.json 2> gpurun_out/r2s_retrieval_q$q.err; echo "retrieval q=$q rc $?"; done
  python -c "
import json
for q in (1,8):
    r=json.loads(open('gpurun_out/r2s_retrieval_q%d.json'%q).read().strip().splitlines()[-1]); print('retr',q,r['ms_per_step'],r['config']['ms_per_query'],r['roofline']['frac'])"
fi
