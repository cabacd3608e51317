set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29511 tools/check_sharded.py --scans 1500 > gpurun_out/r2o_check_sharded_n8.log 2>&1; echo "check rc $?"; tail -1 gpurun_out/r2o_check_sharded_n8.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err; echo "n1 rc $?"
timeout 900 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > gpurun_out/r2o_bench_n8_bulk.json 2> gpurun_out/r2o_bench_n8_bulk.err; echo "bulk rc $?"
NSC_LIB=$PWD/neural_spectral_codec_b200/libnsc_b200_h1.so timeout 900 $TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > gpurun_out/r2o_bench_n8_bulk_h1.json 2> gpurun_out/r2o_bench_n8_bulk_h1.err; echo "h1 rc $?"
python - <<'PY'
import json
def load(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return None
a=load('gpurun_out/r2o_bench_n1.json')
print('n1',a['value'],a['roofline']['frac'])
for t in ('bulk','bulk_h1'):
    b=load(f'gpurun_out/r2o_bench_n8_{t}.json')
    if b: print(t,b['value'],'eff',b['value']/(8*a['value']),'step',b['ms_per_step'],b['checks']['db_identical'],b['checks']['ok'],json.dumps(b['scaling_breakdown']))
PY
