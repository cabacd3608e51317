set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_gpu_tests.log 2>&1; echo "tests rc $?"; tail -2 gpurun_out/r2z_gpu_tests.log
timeout 600 $TR --nproc-per-node 2 --master-port 29511 tools/check_sharded.py --scans 61 > gpurun_out/r2z_check_sharded_n2_61.log 2>&1; echo "check61 rc $?"; tail -1 gpurun_out/r2z_check_sharded_n2_61.log
timeout 600 $TR --nproc-per-node 2 --master-port 29512 tools/check_sharded.py --scans 1500 > gpurun_out/r2z_check_sharded_n2_1500.log 2>&1; echo "check1500 rc $?"; tail -1 gpurun_out/r2z_check_sharded_n2_1500.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; echo "n1 rc $?"
timeout 900 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r2z_bench_n2.json 2> gpurun_out/r2z_bench_n2.err; echo "n2 rc $?"; tail -2 gpurun_out/r2z_bench_n2.err
python - <<'PY'
import json
def load(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return None
a,b=load('gpurun_out/r2z_bench_n1.json'),load('gpurun_out/r2z_bench_n2.json')
print('n1',a['value'],a['roofline']['frac'])
if b: print('n2',b['value'],'eff',b['value']/(2*a['value']),'step',b['ms_per_step'],b['checks']['db_identical'],b['checks']['ok'])
PY
