"""GPU parity: SpJoin (pair, triplet, fused table lookup, value mode) vs the oracle (bit-exact)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po


@pytest.fixture(scope="module")
def lp_spg(mid_graph):
    from surel_plus_b200 import DeviceGraph, SpG, _capi
    A = mid_graph
    n = A.shape[0]
    g = DeviceGraph.from_scipy(A)
    M, m = 100, 3
    spg = SpG.sample(g, np.arange(n), M, m, seed=3, rng_mode=_capi.SUBG_RNG_PHILOX)
    z = spg.to_scipy()
    xpe = torch.from_numpy(spg.enc_table()).float().cuda() / M
    return spg, z, xpe


def _edges(n, B, arity, seed=0):
    rng = np.random.default_rng(seed)
    e = rng.integers(0, n, (arity, B))
    e[:, 0] = e[0, 0]               # u == v query
    if B > 1:
        e[:, 1] = n - 1 - np.arange(arity) % 3  # isolated nodes (sets of size 1)
    return e


@pytest.mark.parametrize("B", [1, 7, 1024])
def test_pair_join_index_and_fused(lp_spg, B):
    from surel_plus_b200 import gather, pgather, bgather
    spg, z, xpe = lp_spg
    edge = _edges(z.shape[0], B, 2, seed=B)
    exz, sl, sr = po.spjoin_pair(z, edge)
    out = np.empty(4, dtype=object)
    bgather(edge, spg, out)
    assert np.array_equal(np.vstack([out[0], out[1]]), exz)
    assert np.array_equal(out[2], sl) and np.array_equal(out[3], sr)
    for fn in (gather, lambda e, x, d, ptr, encode: pgather(e, x, d, encode, bgather, ptr=ptr)):
        xz, ptr = fn(torch.from_numpy(edge), spg, "cuda", ptr=True, encode=xpe)
        assert xz.dtype == torch.float32 and tuple(xz.shape) == (len(exz), 2, xpe.shape[1])
        assert torch.equal(xz.cpu(), xpe.cpu()[torch.from_numpy(exz).long()])
        assert np.array_equal(ptr.cpu().numpy(), po.pair_index(sl, sr, True))
        xz2, ind = fn(edge, spg, "cuda", ptr=False, encode=xpe)
        assert torch.equal(xz2, xz)
        assert np.array_equal(ind.cpu().numpy(), po.pair_index(sl, sr, False))


def test_pair_join_accepts_scipy_and_cuda_edges(lp_spg):
    from surel_plus_b200 import gather
    spg, z, xpe = lp_spg
    edge = _edges(z.shape[0], 300, 2, seed=11)
    exz, sl, sr = po.spjoin_pair(z, edge)
    a, _ = gather(torch.from_numpy(edge).cuda(), spg, "cuda", True, xpe)
    b, _ = gather(edge, z, "cuda", True, xpe)          # scipy CSR uploaded once
    c, _ = gather(edge, z, "cuda", True, xpe)          # cached
    exp = xpe.cpu()[torch.from_numpy(exz).long()]
    assert torch.equal(a.cpu(), exp) and torch.equal(b.cpu(), exp) and torch.equal(c.cpu(), exp)


@pytest.mark.parametrize("B", [1, 5, 2048])
def test_triplet_join(lp_spg, B):
    from surel_plus_b200 import hgather
    spg, z, xpe = lp_spg
    hedge = _edges(z.shape[0], B, 3, seed=100 + B)
    exz, sizes = po.spjoin_triplet(z, hedge)
    xz, ind = hgather(torch.from_numpy(hedge), spg, "cuda", encode=xpe)
    assert torch.equal(xz.cpu(), xpe.cpu()[torch.from_numpy(exz).long()])
    assert np.array_equal(ind.cpu().numpy(), np.repeat(np.arange(4 * B), sizes))
    with pytest.raises(NotImplementedError):
        hgather(hedge, spg, "cuda", encode=None)


def test_value_join_float64(mid_graph):
    """PPR/SPD mode: float64 store, (x+1)-1 rounding of train.py:33, float32 output [N,2,1]."""
    import scipy.sparse as sp
    from surel_plus_b200 import gather
    A = mid_graph
    rng = np.random.default_rng(1)
    n = A.shape[0]
    X = sp.random(n, n, density=30 / n, format="csr", random_state=3, dtype=np.float64)
    X.data = rng.random(X.nnz) * 0.9 + 1e-3
    X.sort_indices()
    edge = _edges(n, 500, 2, seed=4)
    exz, sl, sr = po.spjoin_pair(X, edge)
    xz, ptr = gather(edge, X, "cuda", ptr=True, encode=None)
    assert tuple(xz.shape) == (len(exz), 2, 1)
    assert torch.equal(xz.squeeze(-1).cpu(), torch.from_numpy(exz).float())
    assert np.array_equal(ptr.cpu().numpy(), po.pair_index(sl, sr, True))


def test_large_sets_use_global_path():
    """Sets far larger than shared memory take the generic kernel; same rows."""
    import scipy.sparse as sp
    from surel_plus_b200 import gather
    n = 60000
    rng = np.random.default_rng(2)
    rows = np.repeat(np.arange(4), 30000)
    cols = np.concatenate([np.sort(rng.choice(n, 30000, replace=False)) for _ in range(4)])
    X = sp.csr_matrix((rng.integers(1, 1000, rows.size).astype(np.int32), (rows, cols)), shape=(n, n))
    X.sort_indices()
    edge = np.array([[0, 1, 2, 3, 5], [1, 2, 3, 0, 0]])
    exz, sl, sr = po.spjoin_pair(X, edge)
    xz, ptr = gather(edge, X, "cuda", ptr=True, encode=None)
    assert torch.equal(xz.squeeze(-1).cpu(), torch.from_numpy(exz).float())


def test_bad_query_raises(lp_spg):
    from surel_plus_b200 import gather
    spg, z, xpe = lp_spg
    with pytest.raises(TypeError):
        gather(np.array([[0], [z.shape[0] + 5]]), spg, "cuda", True, xpe)


def test_fused_call_falls_back_when_the_estimate_is_too_small(mid_graph):
    """subg_spjoin sizes the output from the rows per segment seen so far; a batch that outgrows it must take the
    exact two-step route and give the same rows (first: isolated nodes, sets of size 1; then the hubs)."""
    from surel_plus_b200 import DeviceGraph, SpG, gather
    A = mid_graph
    n = A.shape[0]
    g = DeviceGraph.from_scipy(A)
    spg = SpG.sample(g, np.arange(n), 60, 2, seed=3, first_visit_ranks=False)
    z = spg.to_scipy()
    xpe = torch.from_numpy(spg.enc_table()).float().cuda() / 60
    tiny = np.full((2, 2048), n - 1)                       # isolated node: one row per segment
    gather(tiny, spg, "cuda", True, xpe)
    gather(tiny, spg, "cuda", True, xpe)                   # second call runs fused with a capacity of ~1.25 rows / segment
    assert spg._rows_per_seg == 1.0
    hubs = np.stack([np.arange(2048) % 50, (np.arange(2048) * 7) % 50])   # the largest sets
    exz, sl, sr = po.spjoin_pair(z, hubs)
    xz, ptr = gather(hubs, spg, "cuda", True, xpe)
    assert xz.shape[0] == len(exz) > 10 * 4096
    assert torch.equal(xz.cpu(), xpe.cpu()[torch.from_numpy(exz).long()])
    assert np.array_equal(ptr.cpu().numpy(), po.pair_index(sl, sr, True))
    xz2, ptr2 = gather(hubs, spg, "cuda", True, xpe)       # now the estimate fits: fused route, same result
    assert torch.equal(xz2, xz) and torch.equal(ptr2, ptr)


def test_empty_batch(lp_spg):
    from surel_plus_b200 import gather, hgather
    spg, z, xpe = lp_spg
    xz, ptr = gather(np.zeros((2, 0), np.int64), spg, "cuda", True, xpe)
    assert tuple(xz.shape) == (0, 2, xpe.shape[1]) and ptr.cpu().tolist() == [0]
    xz, ind = hgather(np.zeros((3, 0), np.int64), spg, "cuda", xpe)
    assert xz.shape[0] == 0 and ind.numel() == 0


@pytest.mark.parametrize("arity,B", [(2, 1024), (2, 37), (3, 2048), (3, 5), (2, 9000)])
def test_join_stream_graph_replay_equals_gather(mid_graph, arity, B):
    """JoinStream (CUDA-graph replay, no host sync) returns what gather / hgather return: rows [0, N), the segment
    pointers, the segment ids -- across ring slots, for host and device edge lists, with and without the fused LP
    lookup; a batch that outgrows the slot capacity falls back to the exact two-step join."""
    from surel_plus_b200 import DeviceGraph, JoinStream, SpG, gather, hgather
    A = mid_graph
    n, M = A.shape[0], 40
    g = DeviceGraph.from_scipy(A)
    spg = SpG.sample(g, np.arange(n), M, 2, seed=3, first_visit_ranks=False)
    xpe = torch.from_numpy(spg.enc_table()).float().cuda() / M
    rng = np.random.default_rng(B)
    js = JoinStream(spg, B, "cuda", encode=xpe, arity=arity, segid=True, depth=3)
    js_raw = JoinStream(spg, B, "cuda", encode=None, arity=2, depth=2) if arity == 2 else None
    outs = []
    for it in range(7):   # more submits than ring slots
        edge = rng.integers(0, n, (arity, B))
        e_in = torch.from_numpy(edge).cuda() if it % 3 == 2 else edge
        xz, indptr, nrows, segid = js.submit(e_in)
        if arity == 2:
            want_xz, want_ptr = gather(edge, spg, "cuda", True, xpe)
            _, want_seg = gather(edge, spg, "cuda", False, xpe)
        else:
            want_xz, want_seg = hgather(edge, spg, "cuda", xpe)
            want_ptr = None
        N = int(nrows[0].item())
        assert N == want_xz.shape[0] == js.rows() and int(nrows[1].item()) == 0
        assert torch.equal(xz[:N], want_xz) and torch.equal(segid[:N], want_seg)
        if want_ptr is not None:
            assert torch.equal(indptr, want_ptr)
            got_xz, got_ptr = js.gather(e_in)
            assert torch.equal(got_xz, want_xz) and torch.equal(got_ptr, want_ptr)
            raw_xz, raw_ptr = js_raw.gather(edge)
            ref_xz, ref_ptr = gather(edge, spg, "cuda", True, None)
            assert torch.equal(raw_xz, ref_xz) and torch.equal(raw_ptr, ref_ptr)
        else:
            got_xz, got_seg = js.gather(e_in, ptr=False)
            assert torch.equal(got_xz, want_xz) and torch.equal(got_seg, want_seg)
        outs.append(N)
    assert len(set(outs)) > 1
    # capacity overflow: nothing is written, rows() reports it, gather() falls back to the exact route
    tiny = JoinStream(spg, B, "cuda", encode=xpe, arity=arity, segid=True, capacity=8)
    edge = rng.integers(0, n, (arity, B))
    tiny.submit(edge)
    with pytest.raises(MemoryError):
        tiny.rows()
    if arity == 2:
        a, b = tiny.gather(edge)
        wa, wb = gather(edge, spg, "cuda", True, xpe)
        assert torch.equal(a, wa) and torch.equal(b, wb)
    # a node id outside the SpG is reported like gather reports it
    bad = edge.copy()
    bad[0, 0] = n + 5
    js.submit(bad)
    with pytest.raises(TypeError):
        js.rows()
    for x in (js, js_raw, tiny):
        if x is not None:
            x.close()


@pytest.mark.parametrize("B,arity", [(1024, 2), (2048, 3)])
def test_join_stream_prefetch_pattern_two_lanes(mid_graph, B, arity):
    """The training-loop pattern at the reference's batch sizes (train.py:121-127 B = 1024, main_horder.py:33 B = 2048):
    batch k+1 is submitted BEFORE batch k is consumed, nothing is synchronised in between, host-edge batches alternate
    between the joiner's two internal streams.  Every batch's rows, consumed on the caller's stream (a clone queued
    there), equal gather / hgather."""
    from surel_plus_b200 import DeviceGraph, JoinStream, SpG, gather, hgather
    A = mid_graph
    n, M = A.shape[0], 40
    g = DeviceGraph.from_scipy(A)
    spg = SpG.sample(g, np.arange(n), M, 2, seed=5, first_visit_ranks=False)
    xpe = torch.from_numpy(spg.enc_table()).float().cuda() / M
    rng = np.random.default_rng(100 + B)
    edges = [rng.integers(0, n, (arity, B)) for _ in range(12)]
    js = JoinStream(spg, B, "cuda", encode=xpe, arity=arity, segid=True, depth=3)
    got = []
    pending = js.submit(edges[0])
    for k in range(len(edges)):
        nxt = js.submit(edges[k + 1]) if k + 1 < len(edges) else None      # prefetch
        xz, indptr, nrows, segid = pending
        got.append((xz.clone(), indptr.clone(), nrows.clone(), segid.clone()))   # the consumer, on the current stream
        pending = nxt
    torch.cuda.synchronize()
    for edge, (xz, indptr, nrows, segid) in zip(edges, got):
        N = int(nrows[0].item())
        if arity == 2:
            want_xz, want_ptr = gather(edge, spg, "cuda", True, xpe)
            _, want_seg = gather(edge, spg, "cuda", False, xpe)
            assert torch.equal(indptr, want_ptr)
        else:
            want_xz, want_seg = hgather(edge, spg, "cuda", xpe)
        assert N == want_xz.shape[0] and int(nrows[1].item()) == 0
        assert torch.equal(xz[:N], want_xz) and torch.equal(segid[:N], want_seg)
    js.close()
