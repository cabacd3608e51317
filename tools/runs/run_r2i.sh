set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29511 tools/check_sharded.py --scans 61 > gpurun_out/r2i_check_sharded_n2_61.log 2>&1; echo "check61 rc $?"; tail -3 gpurun_out/r2i_check_sharded_n2_61.log
timeout 600 $TR --nproc-per-node 2 --master-port 29512 tools/check_sharded.py --scans 400 > gpurun_out/r2i_check_sharded_n2_400.log 2>&1; echo "check400 rc $?"; tail -3 gpurun_out/r2i_check_sharded_n2_400.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo "n1 rc $?"
timeout 1200 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --c5-scans 20000 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; echo "n2 rc $?"; tail -3 gpurun_out/r2i_bench_n2.err
timeout 900 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --gather nccl --no-extras > gpurun_out/r2i_bench_n2_nccl.json 2> gpurun_out/r2i_bench_n2_nccl.err; echo "n2 nccl rc $?"
python - <<'PY'
import json
def load(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return None
a,b,c=load('gpurun_out/r2i_bench_n1.json'),load('gpurun_out/r2i_bench_n2.json'),load('gpurun_out/r2i_bench_n2_nccl.json')
if a: print('n1',a['value'],a['roofline']['frac'])
if b: print('n2 fused',b['value'], 'eff', b['value']/(2*a['value']) if a else None, b['checks'], b.get('c5'), b['e2e']['value'])
if c: print('n2 nccl',c['value'], 'eff', c['value']/(2*a['value']) if a else None, c['checks']['db_identical'])
PY
