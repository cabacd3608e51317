#!/bin/bash
# A/B of tuning builds and feeds on the GPU box: prints scans/s, roofline fraction, kernel ms.
#   tools/ab.sh "<lib-variant>:<feed> ..."     e.g. tools/ab.sh ":ldg :cpasync t256:ldg t256:cpasync"
cd "$(dirname "$0")/.."
for spec in $1; do
  v="${spec%%:*}"; f="${spec##*:}"
  lib="$PWD/neural_spectral_codec_b200/libnsc_b200${v:+_$v}.so"
  out="gpurun_out/ab_${v:-base}_$f.json"
  NSC_LIB="$lib" NSC_FEED="$f" python bench.py --steps ${STEPS:-30} --warmup 3 --no-cpu > "$out" 2> "${out%.json}.err" || { echo "$spec FAILED"; tail -3 "${out%.json}.err"; continue; }
  python - "$out" "$spec" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(f"{sys.argv[2]:24s} {d['value']/1e6:7.3f} M scans/s  frac {d['roofline']['frac']:.4f}  kernel {d['roofline']['kernel_ms']:.4f} ms")
PY
done
