#!/bin/bash
# Final check after the quantiser rewrite: the whole GPU suite, the quantise bench line, one ncu capture.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r2zt_gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2zt_gpu_tests.log
timeout 120 python bench.py --workload quantize --steps 50 > gpurun_out/r2zt_bench_quantize.json 2>gpurun_out/r2zt_bench_quantize.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:quantize_kernel -s 8 -c 2 \
  -f -o gpurun_out/r2zt_quantize python bench.py --workload quantize --steps 2 --no-cpu > gpurun_out/r2zt_ncu.log 2>&1
tail -3 gpurun_out/r2zt_gpu_tests.log; cat gpurun_out/r2zt_bench_quantize.json | cut -c1-400
