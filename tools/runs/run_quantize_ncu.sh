#!/bin/bash
# One ncu --set full capture of the quantise and the dequantise kernel (bench workload, 100k x 800).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python bench.py --workload quantize --steps 2 --no-cpu > gpurun_out/r2zq_pre.json 2>gpurun_out/r2zq_pre.err || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:quantize_kernel -s 4 -c 2 \
  -f -o gpurun_out/r2zq_quantize python bench.py --workload quantize --steps 2 --no-cpu > gpurun_out/r2zq_ncu.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/r2zq_quantize.ncu-rep
