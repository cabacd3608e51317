"""CPU oracle for stage-1 retrieval (1-D Wasserstein top-K) -- TEST INFRASTRUCTURE ONLY.

Restates the reference's ``src/retrieval/wasserstein.py`` (batch distance :134-172, retriever
:276-389) and the spatial filter of ``src/retrieval/two_stage_retrieval.py:145-202``. Only
``tests/`` and the cpu_baseline leg of ``bench.py --workload retrieval`` may import it. Pinned by
``tests/golden/retrieval.npz`` (outputs of the unmodified reference, recorded by
``tests/golden/make_golden_retrieval.py``).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch


def wasserstein_distance_batch(query: torch.Tensor, database: torch.Tensor,
                               epsilon: float = 1e-8) -> torch.Tensor:
    """``wasserstein_distance_batch_torch`` (wasserstein.py:134-172): float32 ``(n_db,)``."""
    qs = query.sum()                                                    # :152
    if qs > epsilon:                                                    # :153
        query = query / qs                                              # :154  (no epsilon here)
    sums = database.sum(dim=1, keepdim=True)                            # :157
    database = torch.where(sums > epsilon, database / (sums + epsilon), database)   # :158-162
    q_cdf = torch.cumsum(query, dim=0)                                  # :165
    d_cdf = torch.cumsum(database, dim=1)                               # :166
    return torch.abs(d_cdf - q_cdf.unsqueeze(0)).sum(dim=1)            # :169


def query_topk(query: torch.Tensor, database: torch.Tensor, top_k: int = 10,
               epsilon: float = 1e-8) -> Tuple[np.ndarray, np.ndarray]:
    """``WassersteinRetriever.query`` torch branch (wasserstein.py:328-367): indices and
    ascending distances of the ``top_k`` nearest database rows."""
    if database.shape[0] == 0:                                          # :343-344
        return np.array([]), np.array([])
    d = wasserstein_distance_batch(query, database, epsilon)
    k = min(top_k, database.shape[0])                                   # :359
    dist, idx = torch.topk(d, k=k, largest=False)                       # :360-364
    return idx.numpy(), dist.numpy()


def global_retrieval(query: torch.Tensor, query_xyz: Optional[np.ndarray], database: torch.Tensor,
                     database_xyz: Optional[np.ndarray], top_k: int, spatial_filter_distance: float,
                     epsilon: float = 1e-8) -> Tuple[np.ndarray, np.ndarray]:
    """``TwoStageRetrieval._global_retrieval`` (two_stage_retrieval.py:145-202): keyframes closer
    than ``spatial_filter_distance`` to the query position are excluded (strict ``<``, :163),
    then the ``top_k`` nearest remaining ones by Wasserstein distance, ascending."""
    n = database.shape[0]
    valid = np.ones(n, bool)
    if query_xyz is not None and database_xyz is not None:
        dist = np.linalg.norm(np.asarray(database_xyz, np.float64) - np.asarray(query_xyz, np.float64)[None, :], axis=1)
        valid = ~(dist < spatial_filter_distance)
    if valid.sum() == 0:
        return np.array([], np.int64), np.array([], np.float32)
    k = min(top_k, int(valid.sum()))
    idx, d = query_topk(query, database, top_k=n, epsilon=epsilon)     # :182-185 "get all"
    keep = valid[idx]
    return idx[keep][:k], d[keep][:k]
