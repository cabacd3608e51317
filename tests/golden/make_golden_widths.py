#!/usr/bin/env python
"""Golden vectors of the UNMODIFIED reference for image widths other than 360 columns
(n_azimuth is a constructor argument of both reference classes, spectral_encoder.py:35-47,
range_image.py:102-127). Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_widths.py

Per configuration: a synthetic cloud, the projected and the interpolated image, the descriptor of
encode_points, the freq->bin table, forward() on a small batch of images, and both interpolation
methods on random sparse images. Writes tests/golden/widths.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
sys.dont_write_bytecode = True
from encoding.range_image import interpolate_range_image  # noqa: E402  (reference)
from encoding.spectral_encoder import SpectralEncoder  # noqa: E402  (reference)

from neural_spectral_codec_b200 import synth  # noqa: E402

CONFIGS = [
    dict(n_elevation=16, n_azimuth=180, n_bins=50, target_elevation_bins=16),
    dict(n_elevation=16, n_azimuth=500, n_bins=50, target_elevation_bins=16),       # 500 = 2^2 5^3
    dict(n_elevation=16, n_azimuth=1024, n_bins=64, target_elevation_bins=16, alpha=1.5),
    dict(n_elevation=32, n_azimuth=90, n_bins=20, target_elevation_bins=16),         # pooled 32 -> 16
    dict(n_elevation=8, n_azimuth=77, n_bins=10, target_elevation_bins=8, interpolate_empty=False),   # odd, prime factors 7 11
]


def main():
    out = {"n_configs": np.int64(len(CONFIGS))}
    shape = synth.SensorShape("w", 32, -24.8, 2.0, 600)
    rng = np.random.default_rng(21)
    for i, kw in enumerate(CONFIGS):
        enc = SpectralEncoder(**kw)
        pts = synth.make_scan(shape, 900 + i).numpy()
        img, _ = enc.projector.project(pts, keep_intensity=False)
        filled = interpolate_range_image(img, method="linear")
        with torch.no_grad():
            desc = enc.encode_points(pts).numpy()
            edges = enc._compute_bin_edges(enc.alpha)
            k = torch.arange(enc.n_freqs, dtype=torch.float32)
            lut = torch.clamp(torch.searchsorted(edges, k, right=True) - 1, 0, enc.n_bins - 1).numpy()
            W = kw["n_azimuth"]
            imgs = (rng.uniform(1, 60, (3, kw["n_elevation"], W)) * (rng.uniform(0, 1, (3, kw["n_elevation"], W)) < 0.4)).astype(np.float32)
            imgs[1, 2] = 0
            imgs[2, :2] = 0
            fwd = enc(torch.from_numpy(imgs)).numpy()
        lin = np.stack([interpolate_range_image(m, method="linear") for m in imgs])
        near = np.stack([interpolate_range_image(m, method="nearest") for m in imgs])
        out[f"c{i}_kw"] = np.array(repr(kw))
        out[f"c{i}_points"], out[f"c{i}_image"], out[f"c{i}_filled"] = pts, img, filled
        out[f"c{i}_desc"], out[f"c{i}_lut"] = desc, lut.astype(np.int32)
        out[f"c{i}_imgs"], out[f"c{i}_forward"], out[f"c{i}_linear"], out[f"c{i}_nearest"] = imgs, fwd, lin, near
        print(kw, "points", len(pts), "desc sum", float(desc.sum()))
    np.savez_compressed(os.path.join(HERE, "widths.npz"), **out)
    print("wrote widths.npz", os.path.getsize(os.path.join(HERE, "widths.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
