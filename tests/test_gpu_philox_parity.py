"""GPU: bit parity of the kernel that is BENCHMARKED (gset_sample_kernel<.., PARITY=false>, Philox draws).

The rand_r-replay and trace modes run the PARITY=true instantiation; this file pins the fast one: the kernel keeps the
walks it drew (SUBG_SAMPLE_DUMP_WALKS), the oracle (trace-driven restatement of subg_acc.c:778-1000, itself pinned
against the compiled reference) is fed exactly those walks, and everything the kernel produced from them -- set sizes,
first-visit order, LP counts, 64-bit keys, first-occurrence ids, sorted SpG rows -- must be identical, with and
without first-visit ranks (the bench runs without), for the keys-per-lane instantiations of the three LP workloads
(collab EPL 13, ppa EPL 19, dblp EPL 8 (bitonic sort, 4 walks per lane)) and for 32- and 64-bit sort keys.  Feeding the dump back as SUBG_RNG_TRACE
closes the loop inside the library."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po


def _shard(graph, q, lo, hi, M, m, seed, flags, bucket=-1):
    from surel_plus_b200 import SpG, _capi
    from surel_plus_b200.spg import _ptr, _stream
    h = C.c_void_p()
    _capi.check(_capi.load().subg_gset_sample_shard(graph._h, _ptr(q), q.size, lo, hi, M, m, bucket, seed,
                                                     _capi.SUBG_RNG_PHILOX, None, flags, _stream(graph.device), C.byref(h)))
    return SpG(h, graph.device, n_nodes=graph.N, num_walks=M)


def _check_window(A, g, q, lo, hi, M, m, seed, bucket=-1):
    from surel_plus_b200 import SpG, _capi
    # with first-visit ranks: the reference's return list, array for array
    a = _shard(g, q, lo, hi, M, m, seed, _capi.SAMPLE_DUMP_WALKS, bucket)
    walks = a.walks().cpu().numpy()
    assert walks.shape == (hi - lo, M, m)
    # the dumped walks are walks of the graph: every hop follows an edge (or stays put at a dead end)
    prev = np.broadcast_to(q[lo:hi, None], (hi - lo, M)).astype(np.int64)
    for s in range(m):
        cur = walks[:, :, s].astype(np.int64)
        deg = A.indptr[prev + 1] - A.indptr[prev]
        moved = deg > 0
        assert np.array_equal(cur[~moved], prev[~moved])
        samp = np.flatnonzero(moved.ravel())[:: max(1, moved.sum() // 4000)]
        pu, cu = prev.ravel()[samp], cur.ravel()[samp]
        for u, v in zip(pu.tolist(), cu.tolist()):
            row = A.indices[A.indptr[u]:A.indptr[u + 1]]
            assert row[np.searchsorted(row, v)] == v
        prev = cur
    want = po.gset_from_walks(q[lo:hi], walks, M, m, bucket)
    nsize, remap, enc = a.export_reference()
    assert np.array_equal(nsize, want["nsize"])
    iso = np.diff(A.indptr)[q[lo:hi]] == 0
    assert np.array_equal(remap[1], want["remap"][1])
    keep = np.repeat(~iso, nsize)                                 # the reference leaves an isolated seed's id unwritten
    assert np.array_equal(remap[0][keep], want["remap"][0][keep])
    assert np.array_equal(enc, want["enc"])
    # without ranks (what bench.py times): same kernel instantiation minus the rank bitmap; the SpG must be the same
    b = _shard(g, q, lo, hi, M, m, seed, _capi.SAMPLE_DUMP_WALKS | _capi.SAMPLE_NO_RANKS, bucket)
    assert torch.equal(b.walks(), a.walks())
    va, vb = a.views(), b.views()
    for k in ("indptr", "indices", "data", "enc"):
        assert torch.equal(va[k], vb[k]), k
    rows = np.repeat(np.arange(hi - lo), nsize)
    order = np.lexsort((remap[0], rows))                          # ascending node id inside every set (random_walks.py:79-80)
    assert np.array_equal(vb["indices"].cpu().numpy(), remap[0][order])
    assert np.array_equal(vb["data"].cpu().numpy(), remap[1][order] + 1)
    # and the dump fed back as a trace reproduces the same arrays inside the library
    t = SpG.sample(g, q[lo:hi], M, m, bucket=bucket, rng_mode=_capi.SUBG_RNG_TRACE, walks=a.walks())
    tn, tr, te = t.export_reference()
    assert np.array_equal(tn, nsize) and np.array_equal(tr, remap) and np.array_equal(te, enc)
    for x in (a, b, t):
        x.close()
    return int(nsize.sum())


@pytest.mark.parametrize("shape,M,m", [("collab", 200, 2), ("dblp", 100, 2), ("ppa", 200, 3)])
def test_philox_kernel_bit_parity_on_named_shapes(shape, M, m):
    """Full-size graphs of the LP workloads, windows of the all-nodes query (hubs at the low ids, the tail, the
    middle): the global seed index keys the Philox counters, so these are the walks of the full-size pass."""
    from surel_plus_b200 import DeviceGraph
    from surel_plus_b200.graphs import named_graph
    A = named_graph(shape)
    g = DeviceGraph.from_scipy(A, "cuda:0")
    n = A.shape[0]
    q = np.arange(n, dtype=np.int32)
    w = 1500 if shape == "ppa" else 3000
    total = 0
    for lo in (0, n // 2, n - w):
        total += _check_window(A, g, q, lo, lo + w, M, m, 111413)
    assert total > 3 * w
    g.close()


def test_philox_kernel_bit_parity_key64_bucket_and_isolated():
    """64-bit sort keys (more nodes than a 32-bit (node, order) key holds), a bucket smaller than the sets, isolated seeds."""
    from surel_plus_b200 import DeviceGraph
    from surel_plus_b200.graphs import synthetic_graph
    A = synthetic_graph(6_000_000, 9_000_000, seed=11, gamma=2.0, isolated=1000)
    g = DeviceGraph.from_scipy(A, "cuda:0")
    n = A.shape[0]
    rng = np.random.default_rng(5)
    q = np.concatenate([np.arange(2000), n - 1 - np.arange(500), rng.integers(0, n, 1500)]).astype(np.int32)
    q = q[np.sort(np.unique(q, return_index=True)[1])]
    _check_window(A, g, q, 0, q.size, 200, 3, 99)                 # N = 6 M > 2^22 - 1 -> uint64 keys
    _check_window(A, g, q, 100, 1600, 200, 3, 99, bucket=150)     # bucket overflow drops late nodes (subg_acc.c:814-828)
    _check_window(A, g, q, 0, 1200, 60, 2, 5)
    g.close()
