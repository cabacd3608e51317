set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_gpu_tests.log 2>&1; pe=$?; tail -3 gpurun_out/r2q_gpu_tests.log
if [ $pe -eq 0 ]; then
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err; echo "bench rc $?"; tail -2 gpurun_out/r2q_bench_n1.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2q_bench_n1.json').read().strip().splitlines()[-1])
print('frac',d['roofline']['frac'],'per_scan',d['e2e']['per_scan'],'pageable',d['e2e']['pageable_list']['value'])"
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extras"
  $CMD > gpurun_out/r2q_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:encode_points -c 200 --csv --log-file gpurun_out/r2q_launches.csv $CMD > gpurun_out/r2q_ncu1.log 2>&1; echo "ncu launches rc $?"
fi
