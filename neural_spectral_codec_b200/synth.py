"""Seeded synthetic LiDAR scans.

Scans are float32 AoS ``xyzi`` rows, ring-major then azimuth-ordered like a
spinning sensor, which is the layout every reference loader produces
(reference ``src/data/kitti_loader.py:100-115``: ``np.fromfile(float32).reshape(-1, 4)``).
Content is a function of ``(shape, scan_index)`` only -- ``seed = 1234 + scan_index`` --
so a batch is independent of how it is sharded over ranks.

The generator is written in torch ops so the same code fills HBM directly on a
GPU box (bench) or runs on the CPU (tests, golden fixtures). CPU and CUDA RNG
streams differ; parity is always checked on the *same bytes* (copied D2H).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

import torch

BASE_SEED = 1234


@dataclass(frozen=True)
class SensorShape:
    name: str
    rings: int
    el_lo_deg: float
    el_hi_deg: float
    az_steps: int
    dropout: float = 0.08


# C1/C2/C5: KITTI HDL-64E-like; C3: NCLT HDL-32E-like; C4: dense 128-beam.
HDL64 = SensorShape("hdl64", 64, -24.8, 2.0, 2083)
HDL32 = SensorShape("hdl32", 32, -30.67, 10.67, 2400)
BEAM128 = SensorShape("beam128", 128, -25.0, 15.0, 2200)
SHAPES = {s.name: s for s in (HDL64, HDL32, BEAM128)}


def make_scan(shape: SensorShape, scan_index: int, device="cpu", shuffle: bool = False,
              nan_frac: float = 0.001, far_frac: float = 0.005) -> torch.Tensor:
    """One synthetic scan -> float32 ``(N, 4)`` xyzi on ``device``.

    Scene: ground plane at z = -1.73 m, a closed wall whose horizontal distance
    is sinusoidal + stepped in azimuth (8-40 m), Gaussian range noise (2 cm),
    azimuth jitter (1e-4 rad), ~8 % dropout, an open-sky sector with no returns, 0.1 % NaN rows and 0.5 % rows
    pushed beyond max range so both filters of the projector are exercised.
    """
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(BASE_SEED + int(scan_index))
    R, A = shape.rings, shape.az_steps
    f32 = torch.float32

    ph = torch.rand(4, generator=g, device=dev, dtype=f32) * (2 * math.pi)
    el = torch.linspace(math.radians(shape.el_lo_deg), math.radians(shape.el_hi_deg), R,
                        device=dev, dtype=f32).view(R, 1)
    az = (torch.arange(A, device=dev, dtype=f32) * (2 * math.pi / A) - math.pi).view(1, A)
    az = az + 1e-4 * torch.randn(R, A, generator=g, device=dev, dtype=f32)

    wall = (22.0 + 9.0 * torch.sin(3.0 * az + ph[0]) + 5.0 * torch.sign(torch.sin(7.0 * az + ph[1]))
            + 2.5 * torch.sin(13.0 * az + ph[2]) + 1.0 * torch.sin(29.0 * az + ph[3]))
    wall = wall.clamp(4.0, 60.0)
    cos_el, sin_el = torch.cos(el), torch.sin(el)
    r_wall = wall / cos_el
    r_ground = torch.where(sin_el < -1e-3, 1.73 / (-sin_el).clamp_min(1e-3),
                           torch.full_like(sin_el, 1e6))
    rng = torch.minimum(r_wall, r_ground.expand(R, A))
    rng = rng + 0.02 * torch.randn(R, A, generator=g, device=dev, dtype=f32)

    u = torch.rand(R, A, generator=g, device=dev, dtype=f32)
    far = u < far_frac
    rng = torch.where(far, rng * 10.0 + 80.0, rng)

    x = rng * cos_el * torch.cos(az)
    y = rng * cos_el * torch.sin(az)
    z = (rng * sin_el).expand(R, A)
    inten = torch.rand(R, A, generator=g, device=dev, dtype=f32)
    nanrow = (u > 1.0 - nan_frac)
    x = torch.where(nanrow, torch.full_like(x, float("nan")), x)

    keep = torch.rand(R, A, generator=g, device=dev, dtype=f32) >= shape.dropout
    # open-sky sector (~10 % of azimuth): rays above -2 deg give no return at all, so the
    # upper rows of the range image have holes for the interpolation stage to fill.
    sky = (torch.sin(az + ph[3]) > 0.95) & (el > math.radians(-2.0))
    keep = keep & ~sky
    pts = torch.stack([x, y, z, inten], dim=-1).view(R * A, 4)[keep.view(-1)]
    if shuffle:
        perm = torch.randperm(pts.shape[0], generator=g, device=dev)
        pts = pts[perm]
    return pts.contiguous()


def make_batch(shape: SensorShape, first_scan: int, n_scans: int, device="cpu",
               shuffle: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Concatenated scans + CSR offsets: ``(sum N, 4) f32``, ``(n_scans+1,) int64``."""
    scans: List[torch.Tensor] = [make_scan(shape, first_scan + i, device, shuffle)
                                 for i in range(n_scans)]
    counts = torch.tensor([0] + [s.shape[0] for s in scans], dtype=torch.int64)
    offsets = torch.cumsum(counts, 0).to(device)
    if scans:
        pts = torch.cat(scans, 0)
    else:
        pts = torch.zeros(0, 4, dtype=torch.float32, device=device)
    return pts, offsets


def make_batch_resident(shape: SensorShape, first_scan: int, n_scans: int, device,
                        shuffle: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """``make_batch`` for batches that fill a large part of the device: every scan is written
    straight into ONE preallocated buffer (upper bound ``rings * az_steps`` points per scan), so
    the peak is one copy of the batch instead of the list of scans plus their concatenation
    (a 50 000-scan shard of the 100 k-scan config is 96 GB). Same bytes as ``make_batch``."""
    dev = torch.device(device)
    cap = int(n_scans) * shape.rings * shape.az_steps
    buf = torch.empty((cap, 4), dtype=torch.float32, device=dev)
    counts = [0]
    pos = 0
    for i in range(n_scans):
        s = make_scan(shape, first_scan + i, dev, shuffle)
        buf[pos:pos + s.shape[0]] = s
        pos += s.shape[0]
        counts.append(pos)
    offsets = torch.tensor(counts, dtype=torch.int64).to(dev)
    return buf[:pos], offsets


def shard_range(n_scans: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous scan block owned by ``rank``: ceil(B/G) per rank."""
    per = -(-n_scans // world_size)
    lo = min(n_scans, rank * per)
    hi = min(n_scans, lo + per)
    return lo, hi
