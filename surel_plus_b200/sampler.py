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
    own = graph is None
    if own:
        graph = DeviceGraph.from_scipy(G, device)
    z = SpG.sample(graph, idx, num_walks=num_walks, num_steps=num_steps - 1, seed=seed,
                   rng_mode=_capi.SUBG_RNG_PHILOX if rng_mode is None else rng_mode, first_visit_ranks=False)
    if idx.shape[0] != G.shape[0] or not np.array_equal(idx, np.arange(G.shape[0])):
        # a subset / permutation of the nodes: the reference's (N, N) matrix has row idx[i] = set i and empty rows
        # elsewhere (random_walks.py:79); only the row table is rebuilt, the entries stay in place
        z.expand_rows(G.shape[0])
    if own:
        graph.close()
    return z, z.enc_table()


def gen_batch(iterable, n=1, keep=False):
    """sampler/random_walks.py:25-33."""
    length = len(iterable)
    if keep:
        for ndx in range(0, length, n):
            yield iterable[ndx:min(ndx + n, length)]
    else:
        for ndx in range(0, length - n, n):
            yield iterable[ndx:min(ndx + n, length)]


def np_sampling(ptr, neighs, bsize, target, num_walks=200, num_steps=4, device="cuda", nthread=-1, seed=111413):
    """sampler/random_walks.py:36-47: walk_sampler over batches of seeds (first hop without replacement);
    returns (object array of per-seed node ids, stacked landing counts [T, num_steps+1])."""
    from .subg_acc import walk_sampler
    key, freq = [], []
    for batch in gen_batch(target, bsize, True):
        _, freqs = walk_sampler(ptr, neighs, batch, num_walks=num_walks, num_steps=num_steps, replacement=True,
                                nthread=nthread, seed=seed, device=device)
        key.append(freqs[:, 0])
        freq.append(freqs[:, 1])
    return np.concatenate(key), np.vstack(np.hstack(freq))


def rw_matrix(G, train_idx, num_walks=200, num_steps=4, batch_size=2000, reduced=True, device="cuda", nthread=-1,
              seed=111413):
    """Legacy SUREL-v1 preprocessing (sampler/random_walks.py:58-73): same returns (z scipy CSR of
    landing-count row pointers + 1, freqs with the all-zero row 0).  `fastremap.unique/remap` of the
    reference is numpy's unique(return_index/return_inverse): ids follow the sorted projected keys."""
    import scipy.sparse as sp
    gsize = G.shape[0]
    neighbors, freqs = np_sampling(G.indptr, G.indices, batch_size, train_idx, num_walks=num_walks,
                                   num_steps=num_steps - 1, device=device, nthread=nthread, seed=seed)
    if reduced:
        proj = np.array([(num_walks + 1) ** i for i in reversed(range(num_steps))], dtype=np.int64)
        idy = freqs.astype(np.int64) @ proj
        _, idx, idy = np.unique(idy, return_index=True, return_inverse=True)
        freqs = freqs[idx]
    else:
        idy = np.arange(len(freqs))
    i = np.repeat(np.arange(len(neighbors)), np.fromiter(map(len, neighbors), dtype=int))
    j = np.concatenate(neighbors)
    z = sp.csr_matrix((idy + 1, (i, j)), (gsize, gsize))
    freqs = np.insert(freqs, 0, np.zeros((1, num_steps)), axis=0)
    return z, freqs
