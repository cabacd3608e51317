# Round-end sequence of the driver, on one fresh box: reference arm, GPU tests, smoke(), bench.
set -x
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc $?"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_gpu_tests.log 2>&1; echo "tests rc $?"; tail -2 gpurun_out/final_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_b200.json 2> gpurun_out/final_b200.err; echo "bench rc $?"; tail -2 gpurun_out/final_b200.err
timeout 300 python examples/encode_and_retrieve.py > gpurun_out/final_example.log 2>&1; echo "example rc $?"; tail -4 gpurun_out/final_example.log
python - <<'PY'
import json
a=json.loads(open('gpurun_out/final_b200.json').read().strip().splitlines()[-1]); r=json.loads(open('gpurun_out/final_ref.json').read().strip().splitlines()[-1])
print('value',a['value'],'frac',a['roofline']['frac'],'e2e',a['e2e']['value'],'ref',r['value'],r['cpu_baseline']['kind'],r['cpu_baseline']['cores'],'ratio',a['value']/r['value'],'e2e ratio',a['e2e']['value']/r['value'])
print('same config', a['config']==r['config'], 'steps', a['steps'], r['steps'], 'checks', a['checks']['ok'])
print([ (c['shape'],c['shuffled'],round(c['frac'],3)) for c in a['configs']])
PY
