"""Host mirror of sampler/random_walks.py:74-82 (`subg_matrix`) that keeps the SpG in HBM."""
from __future__ import annotations

import numpy as np

from . import _capi
from .spg import DeviceGraph, SpG


def subg_matrix(G, train_idx, num_walks=200, num_steps=4, device="cuda", seed=111413, rng_mode=None, graph=None):
    """Same call as the reference's subg_matrix(G, train_idx, num_walks, num_steps):
    returns (z, enc) where `z` is a device-resident SpG (in place of the scipy CSR; every
    gather/pgather/hgather of surel_plus_b200.train accepts it) and `enc` is the int16 LP table
    with the all-zero row 0 prepended (random_walks.py:81).  Walk length is num_steps-1
    (random_walks.py:78)."""
    print(f'Start sampling for #{len(train_idx)} nodes with {num_walks} {num_steps}-step walks')
    idx = np.asarray(train_idx)
    if idx.shape[0] != G.shape[0] or not np.array_equal(idx, np.arange(G.shape[0])):
        raise NotImplementedError("subg_matrix expects train_idx == arange(G.shape[0]) (as every reference caller passes)")
    own = graph is None
    if own:
        graph = DeviceGraph.from_scipy(G, device)
    z = SpG.sample(graph, idx, num_walks=num_walks, num_steps=num_steps - 1, seed=seed,
                   rng_mode=_capi.SUBG_RNG_PHILOX if rng_mode is None else rng_mode, first_visit_ranks=False)
    if own:
        graph.close()
    return z, z.enc_table()
