#!/usr/bin/env python
"""Golden vectors for NON-default constructor arguments of the unmodified reference encoder, all on
the points of hdl64_small_shuffled.npz / hdl32_small.npz (only outputs are stored here).
Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_params.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")
sys.dont_write_bytecode = True
from encoding.range_image import interpolate_range_image  # noqa: E402  (reference)
from encoding.spectral_encoder import SpectralEncoder  # noqa: E402  (reference)

CASES = [
    dict(points="hdl64_small_shuffled", n_elevation=16, n_bins=30, alpha=1.3, target_elevation_bins=16),
    dict(points="hdl64_small_shuffled", n_elevation=32, n_bins=50, alpha=2.0, target_elevation_bins=8),
    dict(points="hdl32_small", n_elevation=32, n_bins=50, alpha=2.0, target_elevation_bins=16,
         elevation_range=(-30.67, 10.67)),
    dict(points="hdl32_small", n_elevation=16, n_bins=64, alpha=3.0, target_elevation_bins=16,
         elevation_range=(-15.0, 15.0)),                              # training_helipr_to_kitti.yaml:62
    dict(points="hdl64_small_shuffled", n_elevation=64, n_bins=50, alpha=0.5, target_elevation_bins=64,
         learnable_alpha=False),
    dict(points="hdl64_small_shuffled", n_elevation=16, n_bins=50, alpha=2.0, target_elevation_bins=16,
         epsilon=1e-6),
    dict(points="hdl32_small", n_elevation=24, n_bins=50, alpha=2.0, target_elevation_bins=16,
         elevation_range=(-60.0, 60.0)),                              # wide field of view: threshold rows
]


def main():
    out = {"cases": json.dumps(CASES)}
    for i, c in enumerate(CASES):
        kw = {k: v for k, v in c.items() if k != "points"}
        pts = np.load(os.path.join(HERE, c["points"] + ".npz"))["points"]
        enc = SpectralEncoder(n_azimuth=360, **kw)
        img, _ = enc.projector.project(pts, keep_intensity=False)
        filled = interpolate_range_image(img, method="linear")
        desc = enc.encode_points(pts).detach().numpy()
        k = torch.arange(enc.n_freqs, dtype=torch.float32)
        lut = torch.clamp(torch.searchsorted(enc._compute_bin_edges(enc.alpha).detach(), k, right=True) - 1,
                          0, enc.n_bins - 1).numpy()
        out[f"range_image{i}"], out[f"interpolated{i}"] = img, filled
        out[f"descriptor{i}"], out[f"freq_to_bin{i}"] = desc, lut.astype(np.int64)
        print(i, kw, desc.shape, float(desc.sum()))
    np.savez_compressed(os.path.join(HERE, "ctor_params.npz"), **out)


if __name__ == "__main__":
    main()
