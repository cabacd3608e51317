#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (the reference is mounted read-only at
/root/reference and does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports ``encoding.spectral_encoder.SpectralEncoder`` and
``encoding.range_image`` from /root/reference/src, feeds them seeded synthetic
scans and hand-built edge cases, and stores inputs + the reference's outputs
(range image, interpolated image, descriptor, freq->bin LUT, bin edges) as
compressed ``.npz`` files. Nothing of the reference's source is copied; only
its outputs are recorded. ``tests/test_oracle_golden.py`` pins ``oracle/`` to
these vectors, ``tests/test_gpu_parity.py`` checks the CUDA path against them.
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


def reference_outputs(points: np.ndarray, **ctor):
    kw = dict(n_elevation=16, n_azimuth=360, n_bins=50, alpha=2.0, learnable_alpha=True,
              target_elevation_bins=16)
    kw.update(ctor)
    enc = SpectralEncoder(**kw)
    img, _ = enc.projector.project(points, keep_intensity=False)
    filled = interpolate_range_image(img, method="linear") if enc.interpolate_empty else img
    desc = enc.encode_points(points).detach().numpy()
    edges = enc._compute_bin_edges(enc.alpha).detach().numpy()
    k = torch.arange(enc.n_freqs, dtype=torch.float32)
    lut = torch.clamp(torch.searchsorted(enc._compute_bin_edges(enc.alpha).detach(), k, right=True) - 1,
                      0, enc.n_bins - 1).numpy()
    return dict(range_image=img, interpolated=filled, descriptor=desc, bin_edges=edges,
                freq_to_bin=lut.astype(np.int64))


def small(shape: synth.SensorShape, az_steps: int) -> synth.SensorShape:
    return synth.SensorShape(shape.name + "_small", shape.rings, shape.el_lo_deg, shape.el_hi_deg,
                             az_steps, shape.dropout)


def edge_cases():
    f = np.float32
    rng = np.random.default_rng(7)
    cases = {}
    cases["empty"] = np.zeros((0, 4), f)
    cases["all_out_of_range"] = np.array([[0.1, 0.2, 0.1, 0], [200, 0, 0, 0], [0, -90, 5, 1],
                                          [0, 0, 0, 0]], f)
    cases["single_point"] = np.array([[7.5, -3.25, -1.0, 0.5]], f)
    cases["axes"] = np.array([[5, 0, 0, 0], [0, 5, 0, 0], [0, -5, 0, 0], [-5, 0.0, 0, 0],
                              [-5, -0.0, 0, 0], [0, 0, 5, 0], [0, 0, -5, 0], [3, 3, 0, 0],
                              [-3, 3, 0.1, 0], [-3, -3, -0.1, 0], [3, -3, -1, 0]], f)
    nonfinite = (rng.standard_normal((400, 4)) * 12).astype(f)
    nonfinite[::7, 0] = np.nan
    nonfinite[3::11, 1] = np.inf
    nonfinite[5::13, 2] = -np.inf
    nonfinite[::17, 3] = np.nan          # intensity NaN must not drop the point
    cases["nonfinite"] = nonfinite
    # out-of-FOV points clamp into rows 0 / 15
    az = rng.uniform(-np.pi, np.pi, 600)
    el = np.concatenate([rng.uniform(-1.2, -0.5, 300), rng.uniform(0.1, 1.2, 300)])
    r = rng.uniform(2, 60, 600)
    cases["fov_clamp"] = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az),
                                   r * np.sin(el), np.zeros(600)], 1).astype(f)
    # range limits: exactly 1.0, 80.0 and neighbours
    lim = []
    for rr in (1.0, np.nextafter(f(1.0), f(0)), np.nextafter(f(1.0), f(2)), 80.0,
               np.nextafter(f(80.0), f(0)), np.nextafter(f(80.0), f(100)), 79.99999, 80.00001):
        lim.append([rr, 0, 0, 0]); lim.append([0, rr, 0, 0])
        lim.append([rr * 0.6, rr * 0.8, 0, 0]); lim.append([rr * 0.48, rr * 0.64, -rr * 0.6, 0])
    cases["range_limits"] = np.array(lim, f)
    # sparse sensor: 5 rings only -> empty rows (leading, interior, trailing) + 1-pixel row
    pts = []
    for e_deg, n in ((-18.0, 300), (-9.0, 180), (-8.5, 40), (-3.0, 1)):
        a = np.sort(rng.uniform(-np.pi, np.pi, n)); e = np.deg2rad(e_deg)
        rr = 10 + 5 * np.sin(2 * a) + rng.normal(0, 0.02, n)
        pts.append(np.stack([rr * np.cos(e) * np.cos(a), rr * np.cos(e) * np.sin(a),
                             rr * np.sin(e) * np.ones(n), np.zeros(n)], 1))
    cases["sparse_rows"] = np.concatenate(pts).astype(f)
    cases["xyz_only"] = np.ascontiguousarray(cases["fov_clamp"][:, :3])
    return cases


def main():
    out = {}
    # full-size HDL-64 scan (config C1)
    out["hdl64_full"] = synth.make_scan(synth.HDL64, 0).numpy()
    out["hdl64_small_shuffled"] = synth.make_scan(small(synth.HDL64, 521), 1, shuffle=True).numpy()
    out["hdl32_small"] = synth.make_scan(small(synth.HDL32, 600), 2).numpy()
    out["beam128_small"] = synth.make_scan(small(synth.BEAM128, 550), 3).numpy()
    out.update(edge_cases())
    for name, pts in out.items():
        ref = reference_outputs(pts)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), points=pts, **ref)
        print(f"{name:24s} N={pts.shape[0]:7d} sum={ref['descriptor'].sum():.8f} "
              f"empty_px={(ref['range_image'] == 0).sum():5d}")

    # default-constructor path: n_elevation=64 projected, pooled to 16 rows (next-row §8(f).2)
    pts = out["hdl64_small_shuffled"]
    ref = reference_outputs(pts, n_elevation=64)
    np.savez_compressed(os.path.join(HERE, "elev64_pooled.npz"), points=pts, **ref)
    ref = reference_outputs(out["sparse_rows"], n_elevation=64)
    np.savez_compressed(os.path.join(HERE, "elev64_sparse.npz"), points=out["sparse_rows"], **ref)
    ref = reference_outputs(pts, interpolate_empty=False)
    np.savez_compressed(os.path.join(HERE, "no_interp.npz"), points=pts, **ref)

    # forward()/encode_batch on range images: 16-row and 64-row (pooled) batches
    enc16 = SpectralEncoder(n_elevation=16, target_elevation_bins=16)
    imgs16 = np.stack([np.load(os.path.join(HERE, f"{n}.npz"))["range_image"]
                       for n in ("hdl64_full", "hdl32_small", "sparse_rows", "empty")])
    d16 = enc16(torch.from_numpy(imgs16)).detach().numpy()
    enc64 = SpectralEncoder(n_elevation=64, target_elevation_bins=16)
    imgs64 = np.stack([np.load(os.path.join(HERE, f"{n}.npz"))["range_image"]
                       for n in ("elev64_pooled", "elev64_sparse")])
    d64 = enc64(torch.from_numpy(imgs64)).detach().numpy()
    rng = np.random.default_rng(11)
    imgs40 = (rng.uniform(0, 60, (3, 40, 360)) * (rng.uniform(0, 1, (3, 40, 360)) > 0.2)).astype(np.float32)
    enc40 = SpectralEncoder(n_elevation=40, target_elevation_bins=16)
    d40 = enc40(torch.from_numpy(imgs40)).detach().numpy()
    np.savez_compressed(os.path.join(HERE, "forward_batches.npz"), imgs16=imgs16, desc16=d16,
                        imgs64=imgs64, desc64=d64, imgs40=imgs40, desc40=d40)

    # rotation property of the reference itself (spectral_encoder.py:365-415), for the record
    base = out["hdl64_small_shuffled"]
    descs = []
    for k in range(8):
        a = 2 * np.pi * k / 8
        R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        p = base.copy(); p[:, :2] = (base[:, :2].astype(np.float64) @ R.T).astype(np.float32)
        descs.append(enc16.encode_points(p).detach().numpy())
    d = np.array(descs)
    print("reference rotation max|diff| over 8 yaw steps:",
          max(np.abs(d[i] - d[j]).max() for i in range(8) for j in range(i + 1, 8)))
    with open(os.path.join(HERE, "PROVENANCE.txt"), "w") as fh:
        fh.write("generated by tests/golden/make_golden.py from /root/reference (unmodified)\n")
        fh.write(f"numpy {np.__version__} torch {torch.__version__}\n")
        fh.write(f"cpu flags: {'avx512' if 'avx512f' in open('/proc/cpuinfo').read() else 'no-avx512'}\n")


if __name__ == "__main__":
    main()
