set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29521 tools/check_sharded.py --scans 1500 > gpurun_out/r2j_check_sharded_n8.log 2>&1; echo "check rc $?"; tail -2 gpurun_out/r2j_check_sharded_n8.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; echo "n1 rc $?"
timeout 1500 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.err; echo "n8 rc $?"; tail -3 gpurun_out/r2j_bench_n8.err
timeout 900 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras > gpurun_out/r2j_bench_n4.json 2> gpurun_out/r2j_bench_n4.err; echo "n4 rc $?"
python - <<'PY'
import json
def load(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return None
a,b,c=load('gpurun_out/r2j_bench_n1.json'),load('gpurun_out/r2j_bench_n8.json'),load('gpurun_out/r2j_bench_n4.json')
if a: print('n1',a['value'],a['roofline']['frac'])
if b: print('n8 fused',b['value'], 'eff', b['value']/(8*a['value']) if a else None, {k:v for k,v in b['checks'].items() if k!='what'}, b.get('c5'), b['e2e'])
if c: print('n4 fused',c['value'], 'eff', c['value']/(4*a['value']) if a else None, c['checks']['db_identical'])
PY
