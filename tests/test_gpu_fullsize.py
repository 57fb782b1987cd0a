"""GPU: the hot path at BASELINE.json's full sizes, checked through size-independent properties
(the invariants of the reference's own test, subg_acc/test/test.py:34-45, the sortedness asserted by
sampler/random_walks.py:80, determinism, join symmetry) and against the oracle on windows of seeds the
oracle finishes in seconds.  Everything is evaluated on the device; nothing here reads /root/reference."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rows_of(indptr, nodes):
    """Positions of the concatenated rows SpG[nodes[0]], SpG[nodes[1]], ... (device tensors)."""
    beg = indptr[nodes]
    size = indptr[nodes + 1] - beg
    off = torch.cumsum(size, 0) - size
    pos = torch.arange(int(size.sum()), device=indptr.device) - torch.repeat_interleave(off, size) + torch.repeat_interleave(beg, size)
    return pos, size


@pytest.fixture(scope="module")
def ppa():
    from surel_plus_b200 import DeviceGraph
    from surel_plus_b200.graphs import named_graph
    A = named_graph("ppa")
    return A, DeviceGraph.from_scipy(A, "cuda:0")


def test_ppa_full_sampler_invariants(ppa):
    """configs[1]: 576 289 seeds, M=200, m=3 (CLI num_steps=4), Philox fast path."""
    from surel_plus_b200 import SpG
    A, g = ppa
    n, M, m = A.shape[0], 200, 3
    q = torch.arange(n, dtype=torch.int32, device="cuda:0")
    spg = SpG.sample(g, q, num_walks=M, num_steps=m, seed=111413, first_visit_ranks=False)
    sizes = spg.set_sizes()
    assert int(sizes.sum()) == spg.T                                              # test.py:34
    assert 1 <= int(sizes.min()) and int(sizes.max()) == spg.max_set <= M * m + 1
    v = spg.views()
    indptr, indices, data, enc = v["indptr"], v["indices"], v["data"], v["enc"]
    assert int(indptr[-1]) == spg.T and torch.equal((indptr[1:] - indptr[:-1]).to(torch.int32), sizes)
    assert int(indices.max()) == n - 1 and int(data.max()) == spg.c and int(data.min()) == 1   # test.py:36
    # every set is strictly ascending in node id (random_walks.py:80)
    asc = indices[1:] > indices[:-1]
    asc[indptr[1:-1] - 1] = True
    assert bool(asc.all())
    del asc
    # LP invariants: one root row per seed (col 0 == M), every LP column sums to M per seed (test.py:38-40)
    lp = enc.to(torch.int32)[(data - 1).long()]
    assert int((lp[:, 0] == M).sum()) == n
    ends = indptr[1:] - 1
    for col in range(1, m + 1):
        cs = torch.cumsum(lp[:, col].to(torch.int64), 0)
        per_seed = cs[ends] - torch.cat([cs.new_zeros(1), cs[ends[:-1]]])
        assert bool((per_seed == M).all()), f"LP column {col}"
    assert int(enc.max()) == M                                                      # test.py:45
    # the root of every set is the seed itself
    root_pos = torch.nonzero(lp[:, 0] == M).squeeze(1)
    assert torch.equal(indices[root_pos].long(), torch.arange(n, device="cuda:0"))
    del lp, cs, per_seed
    # same seed -> same SpG, bit for bit
    spg2 = SpG.sample(g, q, num_walks=M, num_steps=m, seed=111413, first_visit_ranks=False)
    v2 = spg2.views()
    assert torch.equal(v2["indptr"], indptr) and torch.equal(v2["indices"], indices) and torch.equal(v2["data"], data)
    assert torch.equal(v2["enc"], enc)
    spg2.close()
    spg.close()


def test_ppa_full_spjoin_properties(ppa):
    """Pair SpJoin on the full-size SpG: batch of 1024 x (1 + 20) queries (ogbl-ppa pattern)."""
    from surel_plus_b200 import SpG, gather
    A, g = ppa
    n, M, m, B = A.shape[0], 200, 3, 21504
    q = torch.arange(n, dtype=torch.int32, device="cuda:0")
    spg = SpG.sample(g, q, num_walks=M, num_steps=m, seed=7, first_visit_ranks=False)
    rng = np.random.default_rng(3)
    edge = torch.from_numpy(rng.integers(0, n, (2, B))).cuda()
    edge[1, :64] = edge[0, :64]                      # u == v queries
    xz, ptr = gather(edge, spg, "cuda:0", True, None)   # int pointers [N,2] as float32 (train.py:38-43)
    v = spg.views()
    indptr, data = v["indptr"], v["data"]
    pos_l, size_l = _rows_of(indptr, edge[0])
    pos_r, size_r = _rows_of(indptr, edge[1])
    N = int(size_l.sum() + size_r.sum())
    assert xz.shape[0] == N and int(ptr[-1]) == N and ptr.numel() == 2 * B + 1
    assert torch.equal(ptr[1:] - ptr[:-1], torch.cat([size_l, size_r]))
    own = torch.cat([data[pos_l], data[pos_r]]).float()
    assert torch.equal(xz[:, 0, 0], own)              # rows of S_u in ascending node order, all left then all right
    other = xz[:, 1, 0]
    nl = int(size_l.sum())
    assert int((other[:nl] > 0).sum()) == int((other[nl:] > 0).sum())   # |S_u & S_v| counted from both sides
    # swapping the endpoints swaps the blocks
    xz_s, ptr_s = gather(edge.flip(0), spg, "cuda:0", True, None)
    assert torch.equal(xz_s[:N - nl], xz[nl:]) and torch.equal(xz_s[N - nl:], xz[:nl])
    # u == v: both columns equal
    k0 = int(ptr[64])
    assert torch.equal(xz[:k0, 0, 0], xz[:k0, 1, 0])
    # fused LP lookup == encode[xz] (train.py:37)
    xpe = (torch.from_numpy(spg.enc_table()).float() / M).cuda()
    xf, ptr_f = gather(edge, spg, "cuda:0", True, xpe)
    assert torch.equal(ptr_f, ptr)
    assert torch.equal(xf, xpe[xz[:, :, 0].long()])
    spg.close()


def test_collab_window_vs_oracle_and_dblp_triplets():
    """configs[0]: collab shape, M=200, m=2 -- the first 3000 seeds (the hubs: deg > M, Fisher-Yates first hop)
    replayed bit-exactly against the oracle on the FULL graph; configs[3]: DBLP shape triplet queries, B=2048."""
    from oracle import pyoracle as po
    from surel_plus_b200 import DeviceGraph, SpG, _capi, gather, hgather
    from surel_plus_b200.graphs import named_graph
    A = named_graph("collab")
    g = DeviceGraph.from_scipy(A, "cuda:0")
    q = np.arange(3000, dtype=np.int32)
    spg = SpG.sample(g, q, num_walks=200, num_steps=2, seed=99, rng_mode=_capi.SUBG_RNG_RAND_R)
    got = spg.export_reference()
    exp = po.gset_sampler_replay(A.indptr.astype(np.int32), A.indices.astype(np.int32), q, 200, 2, -1, 99)
    for a, b, name in zip(got, exp, ("nsize", "remap", "enc")):
        assert np.array_equal(a, b), name
    spg.close()
    g.close()

    A = named_graph("dblp")
    n, M, m, B = A.shape[0], 100, 2, 2048
    g = DeviceGraph.from_scipy(A, "cuda:0")
    spg = SpG.sample(g, torch.arange(n, dtype=torch.int32, device="cuda:0"), num_walks=M, num_steps=m, seed=5,
                     first_visit_ranks=False)
    xpe = (torch.from_numpy(spg.enc_table()).float() / M).cuda()
    rng = np.random.default_rng(11)
    hedge = torch.from_numpy(rng.integers(0, n, (3, B))).cuda()
    xh, ind = hgather(hedge, spg, "cuda:0", xpe)
    # the 4 segments [u|w, w|u, v|w, w|v] (train.py:57-68) are two pair joins
    x1, p1 = gather(hedge[[0, 2]], spg, "cuda:0", True, xpe)
    x2, p2 = gather(hedge[[1, 2]], spg, "cuda:0", True, xpe)
    assert xh.shape[0] == x1.shape[0] + x2.shape[0]
    assert torch.equal(xh[:x1.shape[0]], x1) and torch.equal(xh[x1.shape[0]:], x2)
    sizes = torch.cat([p1[1:] - p1[:-1], p2[1:] - p2[:-1]])
    assert torch.equal(ind, torch.repeat_interleave(torch.arange(4 * B, device="cuda:0"), sizes))
    spg.close()


def test_citation2_ppr_full_size_window_vs_oracle_and_mrr_queries():
    """configs[2]: citation2 shape (2.93 M nodes, 61 M directed entries), PPR top-100 alpha 0.1 eps 1e-4 'sym'
    (main.py:44,111): a window of seeds (hubs, mid ids, the tail) bit-exact against the oracle on the FULL graph,
    size-independent properties of every row, and MRR-style 1-vs-1000 queries sharing their source joined on the
    value SpG (float mode of gather, train.py:38-43)."""
    from oracle import pyoracle as po
    from surel_plus_b200 import DeviceGraph, gather, topk_ppr_matrix
    from surel_plus_b200.graphs import named_graph
    A = named_graph("citation2").astype(np.int64)
    n, topk = A.shape[0], 100
    g = DeviceGraph.from_scipy(A, "cuda:0")
    win = np.concatenate([np.arange(0, 150), np.arange(n // 2, n // 2 + 150), np.arange(n - 150, n)])
    exp = po.topk_ppr_matrix(A, 0.1, 1e-4, win, topk, "sym").astype(np.float64).tocsr()
    exp.sort_indices()
    # the product call: graph resident in HBM, unweighted degree == adj.sum(1) for this 0/1 adjacency
    got = topk_ppr_matrix(g, 0.1, 1e-4, win, topk, normalization="sym").to_scipy()
    assert np.array_equal(got.indptr, exp.indptr) and np.array_equal(got.indices, exp.indices)
    assert np.allclose(got.data, exp.data, rtol=1e-5, atol=0)          # north star: scores within 1e-5 relative
    assert np.array_equal(got.data, exp.data)                           # and in fact identical
    # every node a seed
    x = topk_ppr_matrix(g, 0.1, 1e-4, np.arange(n), topk, normalization="sym")
    v = x.views()
    indptr, indices, data = v["indptr"], v["indices"], v["data"]
    sizes = indptr[1:] - indptr[:-1]
    assert x.n == n and int(sizes.max()) <= topk and int(sizes.min()) >= 1
    asc = indices[1:] > indices[:-1]
    asc[indptr[1:-1] - 1] = True
    assert bool(asc.all())
    assert bool(torch.isfinite(data).all()) and float(data.min()) > 0
    # the seed is always in its own top-k (p[s] >= alpha), found by searching its row
    rows = torch.repeat_interleave(torch.arange(n, device="cuda:0"), sizes)
    assert int((indices.long() == rows).sum()) == n
    del rows, asc
    # 1 positive + 1000 negatives per source (MRR evaluation pattern), 16 sources per call
    rng = np.random.default_rng(5)
    src = rng.integers(0, n, 16)
    edge = np.stack([np.repeat(src, 1001), rng.integers(0, n, 16 * 1001)])
    xz, ptr = gather(torch.from_numpy(edge), x, "cuda:0", True, None)
    assert xz.shape[1:] == (2, 1) and xz.dtype == torch.float32 and ptr.numel() == 2 * edge.shape[1] + 1
    B = edge.shape[1]
    ip = ptr.cpu().numpy()
    S = x.to_scipy()
    exz, sl, sr = po.spjoin_pair(S, edge[:, :3003])
    nl = int(sl.sum())
    assert np.array_equal(np.diff(ip)[:3003], sl) and np.array_equal(np.diff(ip)[B:B + 3003], sr)
    assert np.array_equal(xz[:nl, :, 0].cpu().numpy(), exz[:nl].astype(np.float32))
    x.close()
    g.close()
