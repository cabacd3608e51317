set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
SECONDS=0
timeout 1500 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n2_bench_n2_full.json 2> gpurun_out/r2n2_bench_n2_full.err; echo "n2 rc $? after $SECONDS s"; tail -3 gpurun_out/r2n2_bench_n2_full.err | cut -c1-300
python -c "
import json
b=json.loads(open('gpurun_out/r2n2_bench_n2_full.json').read().strip().splitlines()[-1]); print(b['value'], b['checks']['ok'], b['c5'], b['e2e']['value'])"
