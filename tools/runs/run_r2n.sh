set -x
for v in "" h1 h2 "" h1 h2; do
  if [ -z "$v" ]; then L=neural_spectral_codec_b200/libnsc_b200.so; else L=neural_spectral_codec_b200/libnsc_b200_$v.so; fi
  echo "variant ${v:-base}"; NSC_LIB=$PWD/$L timeout 300 python tools/peer_store_cost.py --steps 20
done 2>&1 | tee gpurun_out/r2n_peer_store_cost_hints.txt
