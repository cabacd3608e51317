"""CPU oracle for the keyframe gate geometry -- TEST INFRASTRUCTURE ONLY.

Restates, call for call, the reference's voxel-IoU overlap (``src/data/pose_utils.py:323-389``),
the SE(3) helpers it uses (``:58-76, 90-103, 107-133, 136-187``) and the 4-criterion gate with
early termination (``src/keyframe/criteria.py:53-249``). Only ``tests/`` may import it. Pinned by
``tests/golden/keyframe.npz`` (outputs of the unmodified reference on seeded inputs, recorded by
``tests/golden/make_golden_keyframe.py``); ``tests/test_oracle_keyframe.py`` requires equality.

Arithmetic types (NumPy >= 2, what the in-container reference does):
  * ``transform_points`` stacks float32 points with a float64 column of ones -> float64, then a
    float64 4x4 matmul; the transformed cloud is voxelised in float64;
  * the untransformed cloud keeps its dtype: float32 points are clipped and divided by the Python
    float ``voxel_size`` in float32 (weak scalar promotion), then floored and cast to int32;
  * the random subsample (``np.random.choice(n, max_points, replace=False)``, first cloud first)
    is drawn from NumPy's global generator: equal seeds give equal draws.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def inverse_pose(T: np.ndarray) -> np.ndarray:                        # pose_utils.py:58-76
    out = np.eye(4)
    R, t = T[:3, :3], T[:3, 3]
    out[:3, :3] = R.T
    out[:3, 3] = -R.T @ t
    return out


def relative_pose(T_source: np.ndarray, T_target: np.ndarray) -> np.ndarray:   # :90-103
    return inverse_pose(T_source) @ T_target


def transform_points(points: np.ndarray, T: np.ndarray) -> np.ndarray:          # :107-133
    if points.shape[1] == 3:
        hom = np.hstack([points, np.ones((len(points), 1))])
        return (T @ hom.T).T[:, :3]
    if points.shape[1] == 4:
        hom = np.hstack([points[:, :3], np.ones((len(points), 1))])
        return np.hstack([(T @ hom.T).T[:, :3], points[:, 3:4]])
    raise ValueError(f"Invalid point cloud shape: {points.shape}")


def euclidean_distance(T1: np.ndarray, T2: np.ndarray) -> float:                # :136-149
    return np.linalg.norm(T2[:3, 3] - T1[:3, 3])


def rotation_angle_degrees(T1: np.ndarray, T2: np.ndarray) -> float:            # :152-187
    trace = np.trace(T1[:3, :3].T @ T2[:3, :3])
    return np.degrees(np.arccos(np.clip((trace - 1) / 2, -1, 1)))


def subsample(points1: np.ndarray, points2: np.ndarray, max_points: int = 5000):
    """The two draws of pose_utils.py:343-350, in the reference's order."""
    if len(points1) > max_points:
        points1 = points1[np.random.choice(len(points1), max_points, replace=False)]
    if len(points2) > max_points:
        points2 = points2[np.random.choice(len(points2), max_points, replace=False)]
    return points1, points2


def voxel_keys(points: np.ndarray, voxel_size: float) -> np.ndarray:
    """Unique int32 voxel coordinates ``(n_unique, 3)`` of one cloud (pose_utils.py:356-377)."""
    points = points[np.isfinite(points).all(axis=1)]                  # :358-359 (every column)
    if len(points) == 0:
        return np.zeros((0, 3), np.int32)
    points = np.clip(points, -1e6, 1e6)                                # :367
    coords = np.floor(points / voxel_size).astype(np.int32)[:, :3]    # :369 (dtype of `points` decides)
    return np.unique(coords, axis=0)                                   # :371-377


def overlap_counts(points1: np.ndarray, points2: np.ndarray, T_12: np.ndarray,
                   voxel_size: float = 0.2) -> Tuple[int, int, int]:
    """``(|V1|, |V2|, |V1 & V2|)`` for clouds that are already subsampled."""
    v1 = voxel_keys(transform_points(points1, T_12), voxel_size)       # :353, :379
    v2 = voxel_keys(points2, voxel_size)                               # :380
    s1 = set(map(tuple, v1.tolist()))
    s2 = set(map(tuple, v2.tolist()))
    return len(s1), len(s2), len(s1 & s2)


def compute_overlap(points1: np.ndarray, points2: np.ndarray, T_12: np.ndarray,
                    voxel_size: float = 0.2, max_points: int = 5000) -> float:
    """``compute_overlap`` (pose_utils.py:323-389): IoU of the voxel sets, 0.0 for an empty union."""
    points1, points2 = subsample(points1, points2, max_points)
    n1, n2, inter = overlap_counts(points1, points2, T_12, voxel_size)
    union = n1 + n2 - inter
    return inter / union if union > 0 else 0.0


def should_select_keyframe(pose_current, timestamp_current, points_current, pose_last, timestamp_last,
                           points_last, require_all: bool = False, distance_threshold: float = 0.5,
                           rotation_threshold: float = 15.0, overlap_threshold: float = 0.7,
                           temporal_threshold: float = 5.0, voxel_size: float = 0.2) -> Tuple[bool, dict]:
    """``KeyframeSelectionCriteria.should_select_keyframe`` (criteria.py:156-249): returns
    ``(selected, {"distance", "rotation", "temporal", "overlap"})`` with ``overlap`` None when the
    geometric check is skipped by the early exit or for lack of clouds."""
    dist = euclidean_distance(pose_current, pose_last)                 # :65-69
    rot = rotation_angle_degrees(pose_current, pose_last)              # :87-91
    dt = abs(timestamp_current - timestamp_last)                       # :150-153
    d_ok, r_ok, t_ok = dist > distance_threshold, rot > rotation_threshold, dt > temporal_threshold
    vals = {"distance": float(dist), "rotation": float(rot), "temporal": float(dt), "overlap": None}
    if not require_all and (d_ok or r_ok or t_ok):                     # :204-212
        return True, vals
    g_ok = False
    have = points_current is not None and points_last is not None
    if have:                                                           # :215-224, :115-131
        T_rel = relative_pose(pose_last, pose_current)
        ov = compute_overlap(points_last, points_current, T_rel, voxel_size=voxel_size)
        vals["overlap"] = float(ov)
        g_ok = ov < overlap_threshold
    if require_all:                                                    # :236-240
        crit = [d_ok, r_ok, t_ok] + ([g_ok] if have else [])
        return bool(all(crit)), vals
    return bool(g_ok), vals                                            # :242
