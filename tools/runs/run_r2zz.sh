set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2zz_gpu_tests.log 2>&1; echo "tests rc $?"; tail -2 gpurun_out/r2zz_gpu_tests.log
timeout 300 python bench.py --workload quantize --steps 20 > gpurun_out/r2zz_quantize.json 2> gpurun_out/r2zz_quantize.err; echo "quantize rc $?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2zz_bench_n1.json 2> gpurun_out/r2zz_bench_n1.err; echo "bench rc $?"; tail -2 gpurun_out/r2zz_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2zz_quantize.json').read().strip().splitlines()[-1]); print('quant', d['ms_per_step'], d['config']['dequantise_ms'], d['roofline']['frac'], d['checks'])
d=json.loads(open('gpurun_out/r2zz_bench_n1.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['per_scan']['ms_per_scan'], d['checks']['ok'])"
