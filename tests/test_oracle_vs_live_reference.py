"""The oracles against the reference ITSELF, imported live, on fresh random inputs -- beyond the
committed golden vectors. Runs wherever the reference's sources are reachable (/root/reference in
the build container, or the git-ignored copy oracle/make_ref.py makes); skipped elsewhere. CPU only.
"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import keyframe_oracle as ko
from oracle import nsc_oracle as orc
from oracle import retrieval_oracle as ro

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import make_ref  # noqa: E402

SRC = make_ref.ref_src_path()
pytestmark = pytest.mark.skipif(SRC is None, reason="reference sources not reachable")


@pytest.fixture(scope="module")
def ref():
    sys.dont_write_bytecode = True
    if SRC not in sys.path:
        sys.path.insert(0, SRC)
    import importlib
    mods = {name: importlib.import_module(name) for name in
            ("encoding.spectral_encoder", "encoding.range_image", "data.pose_utils", "keyframe.criteria",
             "retrieval.wasserstein")}
    return mods


def random_cloud(rng, n, spread=(30.0, 30.0, 3.0)):
    p = (rng.standard_normal((n, 4)) * [spread[0], spread[1], spread[2], 1.0]).astype(np.float32)
    p[:, 2] -= 1.0
    if n > 10:
        p[rng.integers(0, n, 3), rng.integers(0, 3, 3)] = np.nan
        p[rng.integers(0, n, 3), 0] *= 40.0          # beyond max range
    return p


@pytest.mark.parametrize("kw", [
    dict(n_elevation=16, n_azimuth=360, n_bins=50, target_elevation_bins=16),
    dict(n_elevation=64, n_azimuth=360, n_bins=50, target_elevation_bins=16),
    dict(n_elevation=16, n_azimuth=250, n_bins=30, target_elevation_bins=16, alpha=1.2),
    dict(n_elevation=16, n_azimuth=360, n_bins=50, target_elevation_bins=16, interpolate_empty=False,
         elevation_range=(-15.0, 15.0)),
])
def test_encoder_oracle_equals_the_live_reference(ref, kw):
    enc = ref["encoding.spectral_encoder"].SpectralEncoder(**kw)
    cfg = orc.OracleConfig(**kw)
    rng = np.random.default_rng(hash(repr(sorted(kw.items()))) % (2 ** 32))
    for n in (0, 1, 57, 4000, 30000):
        pts = random_cloud(rng, n)
        img = enc.projector.project(pts, keep_intensity=False)[0]
        st = orc.stages(pts, cfg)
        np.testing.assert_array_equal(st["range_image"], img)
        if cfg.interpolate_empty:
            np.testing.assert_array_equal(st["interpolated"],
                                          ref["encoding.range_image"].interpolate_range_image(img, method="linear"))
            np.testing.assert_array_equal(orc.interpolate_range_image(img, method="nearest"),
                                          ref["encoding.range_image"].interpolate_range_image(img, method="nearest"))
        with torch.no_grad():
            want = enc.encode_points(pts).numpy()
        np.testing.assert_allclose(st["descriptor"], want, rtol=1e-6, atol=1e-9)


def test_keyframe_oracle_equals_the_live_reference(ref):
    rng = np.random.default_rng(77)
    crit = ref["keyframe.criteria"].KeyframeSelectionCriteria()
    for trial in range(6):
        a = random_cloud(rng, int(rng.integers(100, 9000)), (8.0, 8.0, 1.0))
        b = a[rng.permutation(len(a))[: max(1, len(a) // 2)]] + np.float32(0.01 * trial)
        T = np.eye(4)
        ang = rng.uniform(-0.05, 0.05)
        T[:2, :2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
        T[:3, 3] = rng.uniform(-0.3, 0.3, 3)
        np.random.seed(trial)
        want = ref["data.pose_utils"].compute_overlap(a, b, T)
        np.random.seed(trial)
        assert ko.compute_overlap(a, b, T) == want
        pose_b = T.copy()
        np.random.seed(100 + trial)
        ws, wd = crit.should_select_keyframe(pose_b, 0.3 * trial, b, np.eye(4), 0.0, a, require_all=bool(trial % 2))
        np.random.seed(100 + trial)
        gs, gv = ko.should_select_keyframe(pose_b, 0.3 * trial, b, np.eye(4), 0.0, a, require_all=bool(trial % 2))
        assert gs == ws and gv["distance"] == wd["distance"]["value"] and gv["rotation"] == wd["rotation"]["value"]
        assert gv["overlap"] == wd["geometric"]["value"]


def test_retrieval_oracle_equals_the_live_reference(ref):
    rng = np.random.default_rng(5)
    db = rng.gamma(0.5, 1.0, (700, 800)).astype(np.float32)
    db /= db.sum(1, keepdims=True)
    W = ref["retrieval.wasserstein"]
    r = W.WassersteinRetriever(use_torch=True, device="cpu")
    r.add_to_database(torch.from_numpy(db))
    for i in (0, 13, 699):
        q = torch.from_numpy(db[i] * np.float32(1.7))
        want = W.wasserstein_distance_batch_torch(q, torch.from_numpy(db))
        np.testing.assert_array_equal(ro.wasserstein_distance_batch(q, torch.from_numpy(db)).numpy(), want.numpy())
        wi, wd = r.query(q, top_k=10)
        gi, gd = ro.query_topk(q, torch.from_numpy(db), 10)
        np.testing.assert_array_equal(gi, wi)
        np.testing.assert_array_equal(gd, wd)
