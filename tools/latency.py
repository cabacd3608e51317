#!/usr/bin/env python
"""Latency of small encode batches (device-resident points, CUDA events): one thread-block
cluster per scan vs. one CTA per scan (NSC_SPLIT=0)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_spectral_codec_b200 import SpectralEncoder, synth  # noqa: E402

enc = SpectralEncoder(n_elevation=16, target_elevation_bins=16).to("cuda")
for n in (1, 4, 16, 32, 64, 148):
    pts, offs = synth.make_batch(synth.HDL64, 0, n, device="cuda")
    out = torch.empty((n, 800), device="cuda")
    for _ in range(5):
        enc.encode_points_batch(pts, offs, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        enc.encode_points_batch(pts, offs, out=out)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print(f"NSC_SPLIT={os.environ.get('NSC_SPLIT', '1')} scans={n:4d} median {ts[len(ts)//2]:8.1f} us  min {ts[0]:8.1f} us")
