set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err; echo "n2 rc $?"; tail -3 gpurun_out/r2k_bench_n2.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2k_bench_n2.json').read().strip().splitlines()[-1])
print(b['value'], b['ms_per_step'], b['scaling_breakdown'], b['checks']['db_identical'])
PY
