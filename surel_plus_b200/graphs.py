"""Synthetic CSR graphs of the BASELINE.json shapes.

Undirected simple graphs, symmetrised CSR with sorted columns and no self-loops (what
dataloader.py:119-129 asserts of the real datasets).  Heavy-tailed degrees: one endpoint is
floor(N * U^gamma), the other uniform.  Deterministic in (N, E, seed).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

#: name -> (N nodes, E undirected edges, generator seed)
SHAPES = {
    "collab": (235_868, 1_285_465, 1),
    "ppa": (576_289, 30_326_273, 2),
    "citation2": (2_927_963, 30_561_187, 3),
    "dblp": (1_924_991, 7_904_336, 4),
    "twitter": (41_652_230, 1_468_365_182 // 2, 5),
}


def synthetic_graph(N: int, E: int, seed: int = 0, gamma: float = 2.0, isolated: int = 0) -> sp.csr_matrix:
    """Symmetric boolean CSR with about 2E stored entries (duplicates coalesced).
    The last `isolated` node ids are left without edges (edge-case coverage)."""
    rng = np.random.default_rng(seed)
    n_live = N - isolated
    chunks_r, chunks_c = [], []
    left = E
    while left > 0:
        b = min(left, 1 << 26)
        src = np.minimum((n_live * rng.random(b) ** gamma).astype(np.int64), n_live - 1)
        dst = rng.integers(0, n_live, b)
        keep = src != dst
        chunks_r.append(src[keep].astype(np.int32))
        chunks_c.append(dst[keep].astype(np.int32))
        left -= b
    r = np.concatenate(chunks_r)
    c = np.concatenate(chunks_c)
    rows = np.concatenate([r, c])
    cols = np.concatenate([c, r])
    del r, c
    A = sp.csr_matrix((np.ones(rows.size, dtype=bool), (rows, cols)), shape=(N, N))
    A.sum_duplicates()
    A.data[:] = True
    A.sort_indices()
    return A


def named_graph(name: str, scale: float = 1.0) -> sp.csr_matrix:
    N, E, seed = SHAPES[name]
    return synthetic_graph(max(int(N * scale), 16), max(int(E * scale), 16), seed)
