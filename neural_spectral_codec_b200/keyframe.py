"""Keyframe gate on the GPU: the step that decides WHICH scans reach the encoder.

Host mirror of the reference's gate geometry:

* ``compute_overlap`` -- ``data.pose_utils.compute_overlap`` (reference
  ``src/data/pose_utils.py:323-389``): voxel-set IoU of two clouds, one of them moved by ``T_12``;
* ``KeyframeSelectionCriteria`` -- the class of ``src/keyframe/criteria.py:18-249`` with the same
  constructor, ``check_*`` methods, ``should_select_keyframe`` signature, early exit and
  ``details`` dictionary;
* ``compute_overlap_batch`` / ``select_keyframes`` -- what the reference does not have: the IoU
  of many cloud pairs in ONE kernel launch, and the gate run over a whole recorded sequence
  (``train_multi_dataset.py:152-190`` runs it scan by scan) with the pairs (last keyframe, scan i)
  batched speculatively.

The three pose criteria are a handful of float64 operations and stay on the host, written with
the same NumPy calls as the reference so their values are identical. The voxel IoU runs in
``libnsc_b200.so`` (``nsc_voxel_overlap_batch``); there is no CPU path for it.

Randomness: the reference subsamples clouds above ``max_points`` with the global NumPy generator
(``np.random.choice``, first cloud first). The mirror draws the same numbers in the same order, so
``np.random.seed(s)`` before either implementation gives the same IoU. ``select_keyframes`` keeps
that order too: speculated pairs that turn out not to be needed have their draws rolled back
(``np.random.set_state``).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


# ---- SE(3) helpers (reference src/data/pose_utils.py:58-76, :90-103, :136-187), host float64 ----
def inverse_pose(T: np.ndarray) -> np.ndarray:
    out = np.eye(4)
    R = T[:3, :3]
    out[:3, :3] = R.T
    out[:3, 3] = -R.T @ T[:3, 3]
    return out


def relative_pose(T_source: np.ndarray, T_target: np.ndarray) -> np.ndarray:
    return inverse_pose(T_source) @ T_target


def euclidean_distance(T1: np.ndarray, T2: np.ndarray) -> float:
    return np.linalg.norm(T2[:3, 3] - T1[:3, 3])


def rotation_angle_degrees(T1: np.ndarray, T2: np.ndarray) -> float:
    cos_theta = np.clip((np.trace(T1[:3, :3].T @ T2[:3, :3]) - 1) / 2, -1, 1)
    return np.degrees(np.arccos(cos_theta))


# ---- voxel IoU ---------------------------------------------------------------------------------
def _subsample(points: np.ndarray, max_points: int) -> np.ndarray:
    if len(points) > max_points:                                    # pose_utils.py:343-350
        return points[np.random.choice(len(points), max_points, replace=False)]
    return points


def _as_cloud(points) -> np.ndarray:
    a = np.asarray(points)
    if a.ndim != 2 or a.shape[1] not in (3, 4):
        raise ValueError(f"Invalid point cloud shape: {a.shape}")    # transform_points, :133
    return a


def compute_overlap_batch(pairs: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray]], voxel_size: float = 0.2,
                          device="cuda", return_counts: bool = False):
    """IoU of many ``(points1, points2, T_12)`` pairs in one kernel launch (no subsampling here).

    All clouds of a call must share dtype (float32 or float64) and width (3 or 4 columns).
    Returns float64 ``(n_pairs,)`` on the host; with ``return_counts`` also int32 ``(n_pairs, 3)``:
    voxels of cloud 1, of cloud 2, and of both."""
    lib = _lib.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("compute_overlap: no CPU implementation; pass a CUDA device")
    n_pairs = len(pairs)
    if n_pairs == 0:
        return (np.zeros(0), np.zeros((0, 3), np.int32)) if return_counts else np.zeros(0)
    clouds: List[np.ndarray] = []
    for p1, p2, _ in pairs:
        clouds += [_as_cloud(p1), _as_cloud(p2)]
    widths = {c.shape[1] for c in clouds}
    if len(widths) != 1:
        raise ValueError("all clouds of one call must have the same number of columns")
    f64 = any(c.dtype == np.float64 for c in clouds)
    dt = np.float64 if f64 else np.float32
    if f64 and any(c.dtype != np.float64 for c in clouds):
        raise ValueError("mixed float32 / float64 clouds: the reference's arithmetic depends on the dtype; convert first")
    clouds = [np.ascontiguousarray(c, dtype=dt) for c in clouds]
    offs = np.zeros(2 * n_pairs + 1, np.int64)
    np.cumsum([len(c) for c in clouds], out=offs[1:])
    total = int(offs[-1])
    max_pair = int((offs[2::2] - offs[:-1:2]).max())
    stride = widths.pop()
    pts = np.concatenate(clouds) if total else np.zeros((0, stride), dt)
    T = np.ascontiguousarray(np.stack([np.asarray(t, np.float64).reshape(4, 4) for _, _, t in pairs]))
    d_pts = torch.from_numpy(pts).to(dev)
    d_offs = torch.from_numpy(offs).to(dev)
    d_T = torch.from_numpy(T).to(dev)
    d_cnt = torch.empty((n_pairs, 3), dtype=torch.int32, device=dev)
    d_iou = torch.empty((n_pairs,), dtype=torch.float64, device=dev)
    ws_bytes = int(lib.nsc_voxel_overlap_workspace_bytes(total, max_pair, n_pairs))
    ws = torch.empty((ws_bytes + 3) // 4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = lib.nsc_voxel_overlap_batch(d_pts.data_ptr(), stride, 1 if f64 else 0, d_offs.data_ptr(), total,
                                         max_pair, d_T.data_ptr(), n_pairs, float(voxel_size), d_cnt.data_ptr(),
                                         d_iou.data_ptr(), ws.data_ptr(), ws.numel() * 4,
                                         torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(st, "nsc_voxel_overlap_batch")
    iou = d_iou.cpu().numpy()
    return (iou, d_cnt.cpu().numpy()) if return_counts else iou


def compute_overlap(points1: np.ndarray, points2: np.ndarray, T_12: np.ndarray, voxel_size: float = 0.2,
                    max_points: int = 5000, device="cuda") -> float:
    """``compute_overlap(points1, points2, T_12, voxel_size, max_points) -> IoU`` of the reference
    (pose_utils.py:323-389): random subsample of both clouds (global NumPy generator), cloud 1
    moved into cloud 2's frame, 0.2 m voxels, intersection over union."""
    p1 = _subsample(_as_cloud(points1), max_points)
    p2 = _subsample(_as_cloud(points2), max_points)
    return float(compute_overlap_batch([(p1, p2, T_12)], voxel_size, device)[0])


class KeyframeSelectionCriteria:
    """The 4-criterion gate of the reference (criteria.py:18-249): distance > 0.5 m OR rotation >
    15 deg OR time > 5 s OR voxel IoU with the last keyframe < 0.7, the IoU only evaluated when the
    three cheap criteria fail (or always with ``require_all``)."""

    def __init__(self, distance_threshold: float = 0.5, rotation_threshold: float = 15.0,
                 overlap_threshold: float = 0.7, temporal_threshold: float = 5.0, voxel_size: float = 0.2,
                 device="cuda"):
        self.distance_threshold = distance_threshold
        self.rotation_threshold = rotation_threshold
        self.overlap_threshold = overlap_threshold
        self.temporal_threshold = temporal_threshold
        self.voxel_size = voxel_size
        self.device = device

    def check_distance(self, pose_current, pose_last):
        distance = euclidean_distance(pose_current, pose_last)
        return distance > self.distance_threshold, distance

    def check_rotation(self, pose_current, pose_last):
        rotation = rotation_angle_degrees(pose_current, pose_last)
        return rotation > self.rotation_threshold, rotation

    def check_temporal(self, timestamp_current, timestamp_last):
        time_diff = abs(timestamp_current - timestamp_last)
        return time_diff > self.temporal_threshold, time_diff

    def check_geometric_novelty(self, points_current, points_last, pose_current, pose_last):
        T_rel = relative_pose(pose_last, pose_current)                # criteria.py:115-119
        overlap = compute_overlap(points_last, points_current, T_rel, voxel_size=self.voxel_size,
                                  device=self.device)
        return overlap < self.overlap_threshold, overlap

    def should_select_keyframe(self, pose_current, timestamp_current, points_current, pose_last,
                               timestamp_last, points_last, require_all: bool = False):
        details = {}
        d_ok, d = self.check_distance(pose_current, pose_last)
        details["distance"] = {"satisfied": d_ok, "value": d, "threshold": self.distance_threshold}
        r_ok, r = self.check_rotation(pose_current, pose_last)
        details["rotation"] = {"satisfied": r_ok, "value": r, "threshold": self.rotation_threshold}
        t_ok, t = self.check_temporal(timestamp_current, timestamp_last)
        details["temporal"] = {"satisfied": t_ok, "value": t, "threshold": self.temporal_threshold}
        if not require_all and (d_ok or r_ok or t_ok):                # early exit, criteria.py:204-212
            details["geometric"] = {"satisfied": None, "value": None, "threshold": self.overlap_threshold,
                                    "note": "Skipped (early termination)"}
            details["selected"] = True
            return True, details
        have = points_current is not None and points_last is not None
        if have:
            g_ok, ov = self.check_geometric_novelty(points_current, points_last, pose_current, pose_last)
            details["geometric"] = {"satisfied": g_ok, "value": ov, "threshold": self.overlap_threshold}
        else:
            g_ok = False
            details["geometric"] = {"satisfied": None, "value": None, "threshold": self.overlap_threshold,
                                    "note": "Point clouds not provided"}
        if require_all:
            selected = all([d_ok, r_ok, t_ok] + ([g_ok] if have else []))
        else:
            selected = g_ok
        details["selected"] = selected
        return selected, details


def select_keyframes(scans: Sequence[np.ndarray], poses: np.ndarray, timestamps: Sequence[float],
                     criteria: Optional[KeyframeSelectionCriteria] = None, max_points: int = 5000,
                     window: int = 16):
    """The gate over a recorded sequence: what looping ``KeyframeSelector.process_scan``
    (reference ``src/keyframe/selector.py:96-167``, first scan forced) decides, with the voxel IoUs
    of up to ``window`` consecutive undecided scans against the current last keyframe computed in
    one launch. Returns ``(selected bool (n,), overlap float64 (n,) with NaN where the IoU was not
    needed)``. The decisions -- and, under a seed, the subsample draws -- are those of the
    sequential loop: a speculated pair after the first selected scan of a window is discarded and
    its draws are rolled back."""
    crit = criteria or KeyframeSelectionCriteria()
    n = len(scans)
    selected = np.zeros(n, bool)
    overlap = np.full(n, np.nan)
    if n == 0:
        return selected, overlap
    selected[0] = True
    last, i = 0, 1
    while i < n:
        # scans from i on, against keyframe `last`, until one is selected by a pose criterion
        batch, states = [], []
        j = i
        while j < n and len(batch) < window:
            cheap = (crit.check_distance(poses[j], poses[last])[0] or crit.check_rotation(poses[j], poses[last])[0]
                     or crit.check_temporal(timestamps[j], timestamps[last])[0])
            if cheap:
                break
            p_last = _subsample(_as_cloud(scans[last]), max_points)
            p_cur = _subsample(_as_cloud(scans[j]), max_points)
            batch.append((p_last, p_cur, relative_pose(poses[last], poses[j])))
            states.append(np.random.get_state())
            j += 1
        hit = None
        if batch:
            ious = compute_overlap_batch(batch, crit.voxel_size, crit.device)
            for k, ov in enumerate(ious):
                overlap[i + k] = ov
                if ov < crit.overlap_threshold:
                    hit = i + k
                    break
            if hit is not None and hit + 1 < j:          # later pairs were speculation: undo their draws
                overlap[hit + 1:j] = np.nan
                np.random.set_state(states[hit - i])
        if hit is not None:
            selected[hit] = True
            last, i = hit, hit + 1
        elif j < n and len(batch) < window:               # stopped at a scan selected by a pose criterion
            selected[j] = True
            last, i = j, j + 1
        else:
            i = j
    return selected, overlap
