#!/usr/bin/env python
"""Host cost of one encode_points_batch call (tiny scan, so the GPU is never the bottleneck)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_spectral_codec_b200 import SpectralEncoder, synth  # noqa: E402

enc = SpectralEncoder(n_elevation=16, target_elevation_bins=16).to("cuda")
pts, offs = synth.make_batch(synth.SensorShape("s", 16, -24.8, 2.0, 100), 0, 1, device="cuda")
out = torch.empty((1, 800), device="cuda")
for _ in range(200):
    enc.encode_points_batch(pts, offs, out=out)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 5000
for _ in range(n):
    enc.encode_points_batch(pts, offs, out=out)
torch.cuda.synchronize()
print(f"encode_points_batch: {(time.perf_counter() - t0) / n * 1e6:.1f} us per call (1 tiny scan, device-resident)")
host = synth.make_scan(synth.HDL64, 0).numpy()
for _ in range(20):
    enc.encode_points(host)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    d = enc.encode_points(host).detach().cpu().numpy()
print(f"encode_points(np 120k pts) -> numpy: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us per scan (the reference's per-scan call pattern)")
scans = [synth.make_scan(synth.HDL64, i).numpy() for i in range(256)]   # separate pageable arrays
enc.encode_scans(scans)          # first call allocates the pinned staging buffers
t0 = time.perf_counter()
for _ in range(3):
    enc.encode_scans(scans)
dt = (time.perf_counter() - t0) / 3
nbytes = sum(s.nbytes for s in scans)
print(f"encode_scans(list of 256 pageable arrays): {256 / dt:.0f} scans/s ({nbytes / dt / 1e9:.1f} GB/s of points)")
