set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
NP=${NP:-2}
timeout 600 $TR --nproc-per-node $NP --master-port 29511 tools/check_sharded.py --scans 61 > gpurun_out/r2l_check_sharded_n${NP}_61.log 2>&1; echo "check61 rc $?"; tail -2 gpurun_out/r2l_check_sharded_n${NP}_61.log
timeout 600 $TR --nproc-per-node $NP --master-port 29512 tools/check_sharded.py --scans 1500 > gpurun_out/r2l_check_sharded_n${NP}_1500.log 2>&1; echo "check1500 rc $?"; tail -2 gpurun_out/r2l_check_sharded_n${NP}_1500.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; echo "n1 rc $?"
timeout 900 $TR --nproc-per-node $NP --master-port 29513 bench.py --gpus $NP --steps 20 --warmup 5 --no-extras > gpurun_out/r2l_bench_n${NP}_lag1.json 2> gpurun_out/r2l_bench_n${NP}_lag1.err; echo "lag1 rc $?"; tail -3 gpurun_out/r2l_bench_n${NP}_lag1.err
timeout 900 $TR --nproc-per-node $NP --master-port 29514 bench.py --gpus $NP --steps 20 --warmup 5 --no-extras --gather-lag 0 > gpurun_out/r2l_bench_n${NP}_lag0.json 2> gpurun_out/r2l_bench_n${NP}_lag0.err; echo "lag0 rc $?"
NP=$NP python - <<'PY'
import json, os
NP=int(os.environ['NP'])
def load(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return None
a=load('gpurun_out/r2l_bench_n1.json')
print('n1',a['value'],a['roofline']['frac'])
for lag in (1,0):
    b=load(f'gpurun_out/r2l_bench_n{NP}_lag{lag}.json')
    if b: print('lag',lag,b['value'],'eff',b['value']/(NP*a['value']),'step',b['ms_per_step'],b['checks']['db_identical'],b['checks']['ok'],json.dumps(b['scaling_breakdown']))
PY
