"""Pin oracle/retrieval_oracle.py to the vectors recorded from the unmodified reference."""
import os

import numpy as np
import torch

from conftest import GOLDEN_DIR
from oracle import retrieval_oracle as ro


def test_distances_and_topk_match_reference_vectors():
    g = np.load(os.path.join(GOLDEN_DIR, "retrieval.npz"))
    db = torch.from_numpy(g["database"])
    for i, q in enumerate(g["queries"]):
        d = ro.wasserstein_distance_batch(torch.from_numpy(q), db).numpy()
        np.testing.assert_allclose(d, g["distances"][i], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(d, g["distances_numpy"][i], rtol=2e-5, atol=5e-3)  # reference torch vs numpy path
        idx, dist = ro.query_topk(torch.from_numpy(q), db, 10)
        np.testing.assert_array_equal(idx, g["top10_idx"][i])
        np.testing.assert_allclose(dist, g["top10_dist"][i], rtol=1e-6, atol=1e-7)


def test_spatial_filter_semantics():
    """two_stage_retrieval.py:158-202: strictly closer than the threshold is excluded, the rest
    ranked; fewer valid rows than top_k shortens the answer."""
    rng = np.random.default_rng(0)
    db = torch.from_numpy(rng.random((50, 800)).astype(np.float32))
    xyz = np.stack([np.arange(50) * 10.0, np.zeros(50), np.zeros(50)], 1)
    q = db[7] * 1.0
    idx, d = ro.global_retrieval(q, xyz[7], db, xyz, top_k=5, spatial_filter_distance=50.0)
    assert 7 not in idx and all(abs(i - 7) >= 5 for i in idx) and len(idx) == 5
    assert np.all(np.diff(d) >= 0)
    idx2, _ = ro.global_retrieval(q, xyz[7], db, xyz, top_k=5, spatial_filter_distance=0.0)
    assert idx2[0] == 7
    idx3, _ = ro.global_retrieval(q, xyz[7], db, xyz, top_k=60, spatial_filter_distance=225.0)
    assert len(idx3) == 50 - len([i for i in range(50) if abs(i - 7) * 10 < 225.0])
    idx4, _ = ro.global_retrieval(q, xyz[7], db, xyz, top_k=5, spatial_filter_distance=1e9)
    assert len(idx4) == 0
