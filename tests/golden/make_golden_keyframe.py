#!/usr/bin/env python
"""Records outputs of the UNMODIFIED reference for the keyframe gate geometry:
``data.pose_utils.compute_overlap`` and ``keyframe.criteria.KeyframeSelectionCriteria``.

    python tests/golden/make_golden_keyframe.py        # needs /root/reference (build container)

The reference subsamples clouds above ``max_points`` with the global NumPy generator, so every
case is run under ``np.random.seed(case_seed)`` and the seed is stored: a replay under the same
seed draws the same points. Clouds are small synthetic scans (float32 xyzi, as the loaders give).
Writes ``tests/golden/keyframe.npz``.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
sys.dont_write_bytecode = True

from data.pose_utils import compute_overlap  # noqa: E402  (the reference)
from keyframe.criteria import KeyframeSelectionCriteria  # noqa: E402

from neural_spectral_codec_b200 import synth  # noqa: E402


def pose(x, y, z, yaw_deg, pitch_deg=0.0):
    a, b = np.radians(yaw_deg), np.radians(pitch_deg)
    Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry
    T[:3, 3] = [x, y, z]
    return T


def main():
    out = {}
    shape = synth.SensorShape("kf", 32, -24.8, 2.0, 400)            # ~11 k points per scan
    small = synth.SensorShape("kf_small", 16, -24.8, 2.0, 250)      # ~3.5 k points: no subsample
    rng = np.random.default_rng(11)

    # ---- compute_overlap cases -----------------------------------------------------------
    cases = []
    a = synth.make_scan(small, 1).numpy()
    b = synth.make_scan(small, 2).numpy()
    big_a = synth.make_scan(shape, 3).numpy()
    big_b = synth.make_scan(shape, 4).numpy()
    cases.append(("small_identity", a, a.copy(), np.eye(4), 0.2, 5000))
    cases.append(("small_shift", a, b, pose(0.3, -0.1, 0.0, 2.0), 0.2, 5000))
    cases.append(("small_xyz_only", a[:, :3].copy(), b[:, :3].copy(), pose(0.05, 0.0, 0.0, 0.5), 0.2, 5000))
    cases.append(("small_far", a, b, pose(500.0, 0.0, 0.0, 0.0), 0.2, 5000))
    cases.append(("big_subsampled", big_a, big_b, pose(0.2, 0.1, 0.0, 1.0), 0.2, 5000))
    cases.append(("big_other_limits", big_a, big_b, pose(0.2, 0.1, 0.0, 1.0, 0.3), 0.5, 2000))
    cases.append(("float64_input", a.astype(np.float64), b.astype(np.float64), pose(0.1, 0.1, 0.0, 1.0), 0.2, 5000))
    huge = (rng.standard_normal((600, 4)) * np.array([3e6, 3e6, 10.0, 1.0])).astype(np.float32)   # clipped at +-1e6
    cases.append(("clipped", huge, huge[::-1].copy(), pose(1.0, 2.0, 0.0, 10.0), 0.2, 5000))
    empty = np.zeros((0, 4), np.float32)
    cases.append(("one_empty", a, empty, np.eye(4), 0.2, 5000))
    cases.append(("both_empty", empty, empty, np.eye(4), 0.2, 5000))
    nf = a.copy()
    nf[::7, 0] = np.nan
    nf[3::11, 3] = np.inf                                             # a non-finite INTENSITY drops the point too
    cases.append(("nonfinite", nf, nf[::-1].copy(), pose(0.02, 0.0, 0.0, 0.1), 0.2, 5000))
    grid = (np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(4), indexing="ij"), -1)
            .reshape(-1, 3) * 0.2 + 0.1).astype(np.float32)           # one point per voxel, far from edges
    cases.append(("grid_half", grid, grid + np.float32([1.2, 0, 0]), np.eye(4), 0.2, 5000))
    out["overlap_names"] = np.array([c[0] for c in cases])
    for i, (name, p1, p2, T, vs, mp) in enumerate(cases):
        seed = 1000 + i
        np.random.seed(seed)
        iou = compute_overlap(p1, p2, T, voxel_size=vs, max_points=mp)
        out[f"ov{i}_p1"], out[f"ov{i}_p2"], out[f"ov{i}_T"] = p1, p2, T
        out[f"ov{i}_meta"] = np.array([vs, mp, seed], np.float64)
        out[f"ov{i}_iou"] = np.float64(iou)
        print(f"{name:18s} iou {iou:.6f}")

    # ---- the gate over a short drive -----------------------------------------------------
    crit = KeyframeSelectionCriteria()        # reference defaults: 0.5 m, 15 deg, IoU 0.7, 5 s, 0.2 m
    # a stop-and-go drive of small clouds (no subsample): the sensor stands still for a few scans
    # (same cloud, same pose -> IoU 1, not selected), creeps, and jumps in position / yaw / time.
    unique = [synth.make_scan(small, 50 + i).numpy() for i in range(8)]
    which = [0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 4, 4, 4, 4, 5, 5, 6, 6, 6, 7, 7, 7]
    n = len(which)
    scans = [unique[w] for w in which]
    poses, stamps = [], []
    x, yaw, t = 0.0, 0.0, 0.0
    for i in range(n):
        moved = i > 0 and which[i] != which[i - 1]
        x += 0.6 if i == 9 else (0.04 if moved else 0.0)
        yaw += 20.0 if i == 16 else (0.2 if moved else 0.0)
        t += 6.0 if i == 21 else 0.1
        poses.append(pose(x, 0.0, 0.0, yaw))
        stamps.append(t)
    poses = np.stack(poses)
    stamps = np.array(stamps)
    np.random.seed(4242)
    last = 0
    sel, dist, rot, dt, ov = [True], [0.0], [0.0], [0.0], [np.nan]
    for i in range(1, n):
        s, d = crit.should_select_keyframe(poses[i], stamps[i], scans[i], poses[last], stamps[last], scans[last])
        sel.append(bool(s))
        dist.append(d["distance"]["value"])
        rot.append(d["rotation"]["value"])
        dt.append(d["temporal"]["value"])
        ov.append(np.nan if d["geometric"]["value"] is None else d["geometric"]["value"])
        if s:
            last = i
    out["seq_points"] = np.concatenate(scans)
    out["seq_offsets"] = np.cumsum([0] + [len(s) for s in scans])
    out["seq_poses"], out["seq_stamps"] = poses, stamps
    out["seq_seed"] = np.int64(4242)
    out["seq_selected"] = np.array(sel)
    out["seq_distance"], out["seq_rotation"], out["seq_temporal"] = np.array(dist), np.array(rot), np.array(dt)
    out["seq_overlap"] = np.array(ov)
    print("selected:", np.nonzero(sel)[0].tolist(), " geometric checks run:", int(np.isfinite(ov).sum()))

    # require_all = True on a few pairs (every criterion evaluated, no early exit)
    np.random.seed(99)
    ra = []
    for i, j in ((1, 0), (9, 0), (16, 9), (21, 16), (3, 0)):
        s, d = crit.should_select_keyframe(poses[i], stamps[i], scans[i], poses[j], stamps[j], scans[j], require_all=True)
        ra.append([i, j, float(s), d["geometric"]["value"]])
    out["require_all"] = np.array(ra, np.float64)
    np.savez_compressed(os.path.join(HERE, "keyframe.npz"), **out)
    print("wrote keyframe.npz", os.path.getsize(os.path.join(HERE, "keyframe.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
