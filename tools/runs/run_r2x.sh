set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > gpurun_out/r2x_bench_n1.json 2> gpurun_out/r2x_bench_n1.err; echo "n1 rc $?"
timeout 1500 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2x_bench_n8.json 2> gpurun_out/r2x_bench_n8.err; echo "n8 rc $?"; tail -2 gpurun_out/r2x_bench_n8.err
timeout 900 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras > gpurun_out/r2x_bench_n4.json 2> gpurun_out/r2x_bench_n4.err; echo "n4 rc $?"
timeout 900 $TR --nproc-per-node 2 --master-port 29525 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r2x_bench_n2.json 2> gpurun_out/r2x_bench_n2.err; echo "n2 rc $?"
python - <<'PY'
import json
def load(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return None
a=load('gpurun_out/r2x_bench_n1.json')
print('n1',a['value'],a['roofline']['frac'])
for n in (2,4,8):
    b=load(f'gpurun_out/r2x_bench_n{n}.json')
    if b: print(n,b['value'],'eff',b['value']/(n*a['value']),'step',b['ms_per_step'],b['checks']['db_identical'],b['checks']['ok'], b.get('c5',{}).get('value'), b['e2e']['value'])
PY
