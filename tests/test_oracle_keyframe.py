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


def test_windowed_gate_logic_on_the_cpu(monkeypatch):
    """`keyframe.select_keyframes` batches the IoUs of a window of scans and rolls back the
    subsample draws of speculated pairs. Its host logic is checked here without a GPU: the kernel
    call is replaced by the oracle's exact voxel counts, and the decisions, IoUs and the state of
    NumPy's global generator must equal those of the scan-by-scan loop."""
    from neural_spectral_codec_b200 import keyframe as kf

    def fake_batch(pairs, voxel_size=0.2, device="cuda", return_counts=False):
        out = []
        for p1, p2, T in pairs:
            n1, n2, inter = ko.overlap_counts(p1, p2, T, voxel_size)
            uni = n1 + n2 - inter
            out.append(inter / uni if uni else 0.0)
        return np.array(out)

    monkeypatch.setattr(kf, "compute_overlap_batch", fake_batch)
    rng = np.random.default_rng(12)
    base = [(rng.standard_normal((7000, 4)) * [12, 12, 1.5, 1]).astype(np.float32) for _ in range(4)]
    which = [0, 0, 0, 1, 1, 2, 2, 2, 2, 3, 3, 0, 0, 0, 1]
    scans = [base[w] for w in which]                       # 7000 > max_points: every check subsamples
    n = len(scans)
    poses = np.stack([np.eye(4)] * n)
    for i in range(n):
        poses[i, 0, 3] = 0.02 * i + (0.7 if i >= 9 else 0.0)   # one jump beyond the distance threshold
    stamps = np.arange(n) * 0.1
    crit = kf.KeyframeSelectionCriteria(overlap_threshold=0.3)
    np.random.seed(5)
    last, want_sel, want_ov = 0, [True], [np.nan]
    for i in range(1, n):
        s, v = ko.should_select_keyframe(poses[i], stamps[i], scans[i], poses[last], stamps[last], scans[last],
                                         overlap_threshold=0.3)
        want_sel.append(s)
        want_ov.append(np.nan if v["overlap"] is None else v["overlap"])
        if s:
            last = i
    after = np.random.random()
    assert 2 < sum(want_sel) < n and np.isfinite(want_ov).sum() > 3      # a mix of outcomes
    for window in (1, 2, 5, 32):
        np.random.seed(5)
        sel, ov = kf.select_keyframes(scans, poses, stamps, criteria=crit, window=window)
        np.testing.assert_array_equal(sel, np.array(want_sel))
        np.testing.assert_array_equal(ov, np.array(want_ov))
        assert np.random.random() == after
    # the single-pair mirror under the same patch: same draws, same IoU as the oracle
    np.random.seed(3)
    a = kf.compute_overlap(base[0], base[1], np.eye(4))
    np.random.seed(3)
    assert a == ko.compute_overlap(base[0], base[1], np.eye(4))
