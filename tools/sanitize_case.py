#!/usr/bin/env python
"""Smallest run that touches every kernel once (for compute-sanitizer: one tool per gpurun call)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_spectral_codec_b200 import SpectralEncoder, interpolate_range_image, synth  # noqa: E402
from neural_spectral_codec_b200.quantization import HistogramQuantizer  # noqa: E402
from neural_spectral_codec_b200.retrieval import WassersteinRetriever  # noqa: E402

small = synth.SensorShape("s", 32, -24.8, 2.0, 150)
enc = SpectralEncoder(n_elevation=16, target_elevation_bins=16).to("cuda")
pts, offs = synth.make_batch(small, 0, 80, device="cuda")
d = enc.encode_points_batch(pts, offs)                       # persistent kernel, cp.async feed
o = offs.cpu()
d1 = enc.encode_points_batch(pts[: o[3]], offs[:4])         # cluster kernel
assert torch.equal(d[:3], d1)
d3 = enc.encode_points_batch(pts[:, :3].contiguous(), offs)  # 12-byte points, LDG feed
assert torch.equal(d, d3)
img = enc.projector.project_batch(pts, offs, interpolate=False)
rng_img, inten = enc.projector.project_intensity_batch(pts, offs)
assert torch.equal(img, rng_img)
filled = interpolate_range_image(img)
enc64 = SpectralEncoder(n_elevation=64, target_elevation_bins=16).to("cuda")
enc64(torch.rand(3, 64, 360, device="cuda"))
r = WassersteinRetriever(device="cuda")
r.add_to_database(d, positions=np.arange(240, dtype=np.float64).reshape(80, 3))
idx, top, cnt = r.query_batch(d[:9], top_k=5, query_positions=np.zeros((9, 3)), spatial_filter_distance=30.0)
q = HistogramQuantizer(n_bins=800)
back = q.dequantize(q.quantize(d))
torch.cuda.synchronize()
print("sanitize case ok", float(d.sum()), int(cnt.sum()), float(back.sum()))
