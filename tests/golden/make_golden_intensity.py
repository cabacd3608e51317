#!/usr/bin/env python
"""Golden vectors for RangeImageProjector.project(points, keep_intensity=True) from the
UNMODIFIED reference. Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_intensity.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")
sys.dont_write_bytecode = True
from encoding.range_image import RangeImageProjector  # noqa: E402  (reference)


def main():
    out = {}
    proj = RangeImageProjector(n_elevation=16, n_azimuth=360)
    cases = {n: np.load(os.path.join(HERE, n + ".npz"))["points"]
             for n in ("hdl64_small_shuffled", "beam128_small", "hdl32_small", "nonfinite")}
    # ties: several points with the same rounded range in one pixel, different intensities,
    # negative and zero intensities, a pixel whose only point has negative intensity
    rng = np.random.default_rng(3)
    base = np.array([[10.0, 0.5, -1.0], [-7.0, 3.0, -0.5], [2.0, -9.0, -2.0]], np.float32)
    ties = []
    for b in base:
        for inten in (0.25, 0.75, 0.5, -0.3, 0.0):
            ties.append([b[0], b[1], b[2], inten])
        far = b * np.float32(1.5)
        ties.append([far[0], far[1], far[2], 9.0])          # farther point with larger intensity: ignored
    ties.append([20.0, 20.0, -3.0, -0.7])
    ties = np.array(ties, np.float32)
    cases["intensity_ties"] = ties[rng.permutation(len(ties))]
    for name, pts in cases.items():
        r, i = proj.project(pts, keep_intensity=True)
        out[name + "_points"], out[name + "_range"], out[name + "_intensity"] = pts, r, i
        print(name, pts.shape, "nonzero intensity px", int((i != 0).sum()))
    np.savez_compressed(os.path.join(HERE, "intensity.npz"), **out)


if __name__ == "__main__":
    main()
