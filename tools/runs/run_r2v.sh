set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2v_gpu_tests.log 2>&1; pe=$?; tail -3 gpurun_out/r2v_gpu_tests.log
if [ $pe -eq 0 ]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err; echo "bench rc $?"; tail -2 gpurun_out/r2v_bench_n1.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2v_bench_n1.json').read().strip().splitlines()[-1])
print('frac',d['roofline']['frac'],'per_scan',d['e2e']['per_scan']['ms_per_scan'],d['e2e']['per_scan']['equals_batch_path'],'pageable',d['e2e']['pageable_list']['value'],'checks',d['checks']['ok'])"
  timeout 300 python tools/latency.py > gpurun_out/r2v_latency.txt 2>&1; tail -3 gpurun_out/r2v_latency.txt
fi
