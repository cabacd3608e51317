#!/usr/bin/env python
"""Golden vectors for interpolate_range_image(img, 'linear' | 'nearest') of the UNMODIFIED reference on random
sparse images (hole patterns the projected scans do not produce). Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_interp.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")
sys.dont_write_bytecode = True
from encoding.range_image import interpolate_range_image  # noqa: E402  (reference)


def main():
    rng = np.random.default_rng(5)
    imgs = []
    for density in (0.02, 0.1, 0.3, 0.7, 0.95):
        for _ in range(4):
            imgs.append((rng.uniform(1, 60, (16, 360)) * (rng.uniform(0, 1, (16, 360)) < density)).astype(np.float32))
    a = imgs[3]; a[4:9] = 0                      # interior empty rows
    a = imgs[5]; a[:3] = 0                       # leading empty rows
    a = imgs[6]; a[13:] = 0                      # trailing empty rows
    imgs[7][:] = 0                               # all empty
    a = imgs[8]; a[2] = 0; a[2, 77] = 5.5        # single valid pixel -> constant row
    a = imgs[9]; a[5] = 0; a[5, 0] = 3.0; a[5, 359] = 9.0      # wrap-around neighbours
    a = imgs[10]; a[:] = 0; a[7, 100:200] = 12.5  # one non-empty row feeds all the others
    imgs = np.stack(imgs)
    out = np.stack([interpolate_range_image(i, method="linear") for i in imgs])
    near = np.stack([interpolate_range_image(i, method="nearest") for i in imgs])
    np.savez_compressed(os.path.join(HERE, "interp_random.npz"), images=imgs, interpolated=out, nearest=near)
    print(imgs.shape, "holes filled:", int(((imgs == 0) & (out != 0)).sum()))


if __name__ == "__main__":
    main()
