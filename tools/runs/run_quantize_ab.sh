#!/bin/bash
# Same-box A/B of the quantise / dequantise kernels: library variants built with
#   make -C neural_spectral_codec_b200/csrc VARIANT=<v> DEFS="-DNSC_Q_MIN_BLOCKS=.. -DNSC_Q_BATCH=.."
# (qold = the round-1 scalar kernels). Two interleaved rounds.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_quantization.py -x -q -m gpu > gpurun_out/r2zs_quantize_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2zs_quantize_tests.log
: > gpurun_out/r2zs_quantize_ab.txt
for rep in 0 1; do
  for v in "" qold q4 q5 q7; do
    if [ -z "$v" ]; then lib=""; label=product; else lib="$PWD/neural_spectral_codec_b200/libnsc_b200_$v.so"; label=$v; fi
    out=$(NSC_LIB="$lib" timeout 200 python bench.py --workload quantize --steps 50 --no-cpu 2>gpurun_out/r2zs_err_$label.log | tail -1)
    echo "$label rep $rep $out" | python -c "
import sys, json
l = sys.stdin.read().strip(); lab, _, rep, js = l.split(' ', 3); d = json.loads(js)
print(f\"{lab:8s} rep {rep}  quantise {d['ms_per_step']:.4f} ms frac {d['roofline']['frac']:.3f}   dequantise {d['config']['dequantise_ms']:.4f} ms frac {d['roofline']['dequantise_frac']:.3f}  sums {d['checks']['row_sums_65535']}\")
" >> gpurun_out/r2zs_quantize_ab.txt 2>&1
  done
done
python bench.py --workload quantize --steps 50 > gpurun_out/r2zs_bench_quantize.json 2>gpurun_out/r2zs_bench_quantize.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:quantize_kernel -s 4 -c 2 \
  -f -o gpurun_out/r2zs_quantize python bench.py --workload quantize --steps 2 --no-cpu > gpurun_out/r2zs_ncu.log 2>&1
cat gpurun_out/r2zs_quantize_tests.log | tail -3; cat gpurun_out/r2zs_quantize_ab.txt
