"""The keyframe-gate oracle (oracle/keyframe_oracle.py) against outputs of the unmodified reference
recorded in tests/golden/keyframe.npz (tests/golden/make_golden_keyframe.py). CPU only."""
import os

import numpy as np

from conftest import GOLDEN_DIR
from oracle import keyframe_oracle as ko

G = np.load(os.path.join(GOLDEN_DIR, "keyframe.npz"))


def overlap_cases():
    for i, name in enumerate(G["overlap_names"]):
        vs, mp, seed = G[f"ov{i}_meta"]
        yield str(name), G[f"ov{i}_p1"], G[f"ov{i}_p2"], G[f"ov{i}_T"], float(vs), int(mp), int(seed), float(G[f"ov{i}_iou"])


def test_compute_overlap_equals_the_reference_under_the_recorded_seeds():
    for name, p1, p2, T, vs, mp, seed, want in overlap_cases():
        np.random.seed(seed)
        got = ko.compute_overlap(p1, p2, T, voxel_size=vs, max_points=mp)
        assert got == want, (name, got, want)


def test_gate_sequence_equals_the_reference():
    pts, offs = G["seq_points"], G["seq_offsets"]
    scans = [pts[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    poses, stamps = G["seq_poses"], G["seq_stamps"]
    np.random.seed(int(G["seq_seed"]))
    last = 0
    for i in range(1, len(scans)):
        sel, v = ko.should_select_keyframe(poses[i], stamps[i], scans[i], poses[last], stamps[last], scans[last])
        assert sel == bool(G["seq_selected"][i]), i
        assert v["distance"] == G["seq_distance"][i] and v["rotation"] == G["seq_rotation"][i]
        assert v["temporal"] == G["seq_temporal"][i]
        want = G["seq_overlap"][i]
        assert (v["overlap"] is None and np.isnan(want)) or v["overlap"] == want, i
        if sel:
            last = i
    np.random.seed(99)
    for i, j, want_sel, want_ov in G["require_all"]:
        i, j = int(i), int(j)
        sel, v = ko.should_select_keyframe(poses[i], stamps[i], scans[i], poses[j], stamps[j], scans[j], require_all=True)
        assert sel == bool(want_sel) and v["overlap"] == want_ov


def test_properties():
    rng = np.random.default_rng(0)
    a = rng.uniform(-20, 20, (2000, 3)).astype(np.float32)
    assert ko.compute_overlap(a, a, np.eye(4)) == 1.0
    assert ko.compute_overlap(a, a + np.float32(1000.0), np.eye(4)) == 0.0
    n1, n2, inter = ko.overlap_counts(a, a[:1000], np.eye(4))
    assert inter == n2 <= n1
