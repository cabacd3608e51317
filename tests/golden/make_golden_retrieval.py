#!/usr/bin/env python
"""Golden vectors for stage-1 retrieval from the UNMODIFIED reference
(/root/reference/src/retrieval/wasserstein.py, loaded by file path so that the package's
open3d-dependent siblings are not imported). Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_retrieval.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
spec = importlib.util.spec_from_file_location("ref_wasserstein", "/root/reference/src/retrieval/wasserstein.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def main():
    rng = np.random.default_rng(2024)
    # database: the reference descriptors of the golden scans + random histograms with the
    # descriptor's structure (heavy DC bin per row), + degenerate rows
    descs = [np.load(os.path.join(HERE, f"{n}.npz"))["descriptor"] for n in
             ("hdl64_full", "hdl32_small", "beam128_small", "hdl64_small_shuffled", "sparse_rows",
              "single_point", "empty", "fov_clamp")]
    rand = rng.gamma(0.3, 1.0, (500, 800)).astype(np.float32)
    rand[:, ::50] += 20 * rng.random((500, 16)).astype(np.float32)
    rand /= rand.sum(1, keepdims=True)
    unnorm = (rng.random((20, 800)) * 7).astype(np.float32)          # rows that are not normalised
    zeros = np.zeros((3, 800), np.float32)                           # sum <= eps -> left unnormalised
    tiny = np.full((2, 800), 1e-12, np.float32)
    db = np.concatenate([np.stack(descs), rand, unnorm, zeros, tiny]).astype(np.float32)
    queries = np.concatenate([db[[0, 3, 9, 100, 510]] * np.float32(1.0),
                              (db[[1, 50]] + 0.02 * rng.random((2, 800)).astype(np.float32) / 800),
                              (rng.random((2, 800)) * 3).astype(np.float32),     # unnormalised query
                              np.zeros((1, 800), np.float32)]).astype(np.float32)
    retr = ref.WassersteinRetriever(use_torch=True, device="cpu")
    retr.add_to_database(db[:100])
    retr.add_to_database(db[100:])          # grown in two pieces like repeated add_keyframe calls
    dist = np.stack([ref.wasserstein_distance_batch_torch(torch.from_numpy(q), torch.from_numpy(db)).numpy()
                     for q in queries])
    dist_np = np.stack([ref.wasserstein_distance_batch_numpy(q, db) for q in queries])
    top = [retr.query(q, top_k=10) for q in queries]
    idx10 = np.stack([t[0] for t in top])
    d10 = np.stack([t[1] for t in top])
    np.savez_compressed(os.path.join(HERE, "retrieval.npz"), database=db, queries=queries,
                        distances=dist, distances_numpy=dist_np, top10_idx=idx10, top10_dist=d10)
    print("db", db.shape, "queries", queries.shape, "dist range", dist.min(), dist.max(),
          "max |torch - numpy| =", np.abs(dist - dist_np).max())


if __name__ == "__main__":
    main()
