"""CUDA keyframe-gate geometry (nsc_voxel_overlap_batch, neural_spectral_codec_b200/keyframe.py)
against outputs of the unmodified reference (tests/golden/keyframe.npz) and the oracle.

The bar is EXACT equality: voxel counts are integers and the IoU is one float64 division of them.
(The transform is evaluated in float64 on both sides; a different summation order inside the
reference's BLAS call could move a coordinate by one ulp, which changes a voxel only for a point
within 1e-15 m of a voxel face.)"""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import keyframe_oracle as ko

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(GOLDEN_DIR, "keyframe.npz"))


def overlap_cases():
    for i, name in enumerate(G["overlap_names"]):
        vs, mp, seed = G[f"ov{i}_meta"]
        yield str(name), G[f"ov{i}_p1"], G[f"ov{i}_p2"], G[f"ov{i}_T"], float(vs), int(mp), int(seed), float(G[f"ov{i}_iou"])


def test_compute_overlap_equals_the_reference_vectors():
    from neural_spectral_codec_b200 import keyframe as kf
    for name, p1, p2, T, vs, mp, seed, want in overlap_cases():
        np.random.seed(seed)
        got = kf.compute_overlap(p1, p2, T, voxel_size=vs, max_points=mp)
        assert got == want, (name, got, want)


def test_batch_counts_equal_the_oracle():
    from neural_spectral_codec_b200 import keyframe as kf
    rng = np.random.default_rng(3)
    pairs = []
    for i in range(40):
        n1, n2 = int(rng.integers(0, 6000)), int(rng.integers(0, 6000))
        a = (rng.standard_normal((n1, 4)) * [15, 15, 2, 1]).astype(np.float32)
        b = (rng.standard_normal((n2, 4)) * [15, 15, 2, 1]).astype(np.float32)
        if i % 5 == 0 and n1:
            b = np.concatenate([b, a[: n1 // 2]])                      # guaranteed common voxels
        ang = rng.uniform(-0.2, 0.2)
        T = np.eye(4)
        T[:2, :2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
        T[:3, 3] = rng.uniform(-1, 1, 3)
        if i % 5 == 0:
            T = np.eye(4)
        pairs.append((a, b, T))
    iou, cnt = kf.compute_overlap_batch(pairs, 0.2, return_counts=True)
    for i, (a, b, T) in enumerate(pairs):
        want = ko.overlap_counts(a, b, T, 0.2)
        assert tuple(cnt[i]) == want, (i, cnt[i], want)
        uni = want[0] + want[1] - want[2]
        assert iou[i] == (want[2] / uni if uni else 0.0)
    # xyz-only and float64 clouds, another voxel size
    for dt, w in ((np.float32, 3), (np.float64, 4), (np.float64, 3)):
        p = [(a[:, :w].astype(dt), b[:, :w].astype(dt), T) for a, b, T in pairs[:8]]
        iou, cnt = kf.compute_overlap_batch(p, 0.35, return_counts=True)
        for i, (a, b, T) in enumerate(p):
            assert tuple(cnt[i]) == ko.overlap_counts(a, b, T, 0.35)


def test_pairs_too_large_for_shared_memory_use_the_global_table():
    from neural_spectral_codec_b200 import keyframe as kf
    rng = np.random.default_rng(5)
    a = (rng.standard_normal((30000, 3)) * [25, 25, 3]).astype(np.float32)
    b = np.concatenate([a[:12000] + np.float32(0.01), (rng.standard_normal((9000, 3)) * [25, 25, 3]).astype(np.float32)])
    T = np.eye(4)
    T[0, 3] = 0.05
    iou, cnt = kf.compute_overlap_batch([(a, b, T), (b, a, np.eye(4))], 0.2, return_counts=True)
    assert tuple(cnt[0]) == ko.overlap_counts(a, b, T, 0.2)
    assert tuple(cnt[1]) == ko.overlap_counts(b, a, np.eye(4), 0.2)
    np.random.seed(8)
    got = kf.compute_overlap(a, b, T, max_points=20000)               # the reference's argument, beyond its default
    np.random.seed(8)
    assert got == ko.compute_overlap(a, b, T, max_points=20000)


def test_gate_sequence_equals_the_reference():
    from neural_spectral_codec_b200 import keyframe as kf
    pts, offs = G["seq_points"], G["seq_offsets"]
    scans = [pts[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    poses, stamps = G["seq_poses"], G["seq_stamps"]
    crit = kf.KeyframeSelectionCriteria()
    np.random.seed(int(G["seq_seed"]))
    last = 0
    for i in range(1, len(scans)):
        sel, d = crit.should_select_keyframe(poses[i], stamps[i], scans[i], poses[last], stamps[last], scans[last])
        assert bool(sel) == bool(G["seq_selected"][i]), i
        assert d["distance"]["value"] == G["seq_distance"][i] and d["rotation"]["value"] == G["seq_rotation"][i]
        assert d["temporal"]["value"] == G["seq_temporal"][i]
        want = G["seq_overlap"][i]
        got = d["geometric"]["value"]
        assert (got is None and np.isnan(want)) or got == want, (i, got, want)
        if sel:
            last = i
    np.random.seed(99)
    for i, j, want_sel, want_ov in G["require_all"]:
        i, j = int(i), int(j)
        sel, d = crit.should_select_keyframe(poses[i], stamps[i], scans[i], poses[j], stamps[j], scans[j], require_all=True)
        assert bool(sel) == bool(want_sel) and d["geometric"]["value"] == want_ov
    # the batched sequence gate takes the same decisions with the same IoUs
    for window in (1, 4, 16):
        np.random.seed(int(G["seq_seed"]))
        sel, ov = kf.select_keyframes(scans, poses, stamps, window=window)
        np.testing.assert_array_equal(sel, G["seq_selected"])
        np.testing.assert_array_equal(ov, G["seq_overlap"])


def test_batched_gate_keeps_the_draw_order_of_the_sequential_loop():
    """Clouds above max_points are subsampled with the global generator: the windowed gate must
    consume exactly the draws the scan-by-scan loop consumes (speculated pairs rolled back)."""
    from neural_spectral_codec_b200 import keyframe as kf, synth
    shape = synth.SensorShape("kf", 32, -24.8, 2.0, 400)
    base = [synth.make_scan(shape, 70 + i).numpy() for i in range(3)]
    which = [0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 2, 2]
    scans = [base[w] for w in which]
    poses = np.stack([np.eye(4)] * len(which))
    stamps = np.arange(len(which)) * 0.1
    crit = kf.KeyframeSelectionCriteria(overlap_threshold=0.2)       # subsampled self-overlap is ~0.25
    np.random.seed(21)
    last, want_sel, want_ov = 0, [True], [np.nan]
    for i in range(1, len(scans)):
        s, v = ko.should_select_keyframe(poses[i], stamps[i], scans[i], poses[last], stamps[last], scans[last],
                                         overlap_threshold=0.2)
        want_sel.append(s)
        want_ov.append(np.nan if v["overlap"] is None else v["overlap"])
        if s:
            last = i
    tail_draw = np.random.random()
    for window in (1, 3, 8):
        np.random.seed(21)
        sel, ov = kf.select_keyframes(scans, poses, stamps, criteria=crit, window=window)
        np.testing.assert_array_equal(sel, np.array(want_sel))
        np.testing.assert_array_equal(ov, np.array(want_ov))
        assert np.random.random() == tail_draw                       # generator left in the same state
