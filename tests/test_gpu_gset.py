"""GPU parity: CUDA set sampler / LP encoder / SpG build vs the oracle (bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po


def _sample(graph, A, query, M, m, bucket, seed, mode, walks=None):
    from surel_plus_b200 import SpG
    spg = SpG.sample(graph, query, num_walks=M, num_steps=m, bucket=bucket, seed=seed, rng_mode=mode, walks=walks)
    nsize, remap, enc, raw = spg.export_reference(want_raw=True)
    return spg, nsize, remap, enc, raw


def _assert_same(got, exp, tag):
    for g, e, name in zip(got, exp, ["nsize", "remap", "enc", "raw"]):
        assert g.shape == e.shape, f"{tag}: {name} shape {g.shape} vs {e.shape}"
        assert np.array_equal(g, e), f"{tag}: {name} differs at {np.argwhere(g != e)[:5].tolist()}"


@pytest.mark.parametrize("M,m,bucket", [(20, 3, -1), (50, 2, -1), (20, 3, 15), (200, 2, -1), (7, 4, -1),
                                        (200, 3, -1), (100, 2, -1), (33, 1, -1), (64, 2, -1), (32, 4, -1),
                                        (20, 6, -1), (100, 5, -1), (12, 9, 40), (250, 7, -1)])
def test_rand_r_replay_bit_exact(small_graph, M, m, bucket):
    """SUBG_RNG_RAND_R == reference gset_sampler(nthread=1) (via the pinned oracle), whole output."""
    from surel_plus_b200 import DeviceGraph, _capi
    A = small_graph
    q = np.arange(A.shape[0])
    g = DeviceGraph.from_scipy(A)
    _, nsize, remap, enc, raw = _sample(g, A, q, M, m, bucket, 99, _capi.SUBG_RNG_RAND_R)
    e_nsize, e_remap, e_enc, e_raw = po.gset_sampler_replay(A.indptr, A.indices, q, M, m, bucket, 99, debug=1)
    _assert_same((nsize, remap, enc, raw), (e_nsize, e_remap, e_enc, e_raw), f"M={M} m={m} b={bucket}")


@pytest.mark.parametrize("M,m", [(200, 2), (200, 3), (100, 2)])
def test_trace_mode_bit_exact(mid_graph, M, m):
    """Same walk traces in -> identical (nsize, remap, enc, raw_enc) out (north-star correctness #1)."""
    from surel_plus_b200 import DeviceGraph, _capi
    A = mid_graph
    rng = np.random.default_rng(5)
    q = rng.permutation(A.shape[0])[:3000].astype(np.int32)
    walks, _ = po.walks_rand_r(A.indptr, A.indices, q, M, m, seed=1234)
    g = DeviceGraph.from_scipy(A)
    _, nsize, remap, enc, raw = _sample(g, A, q, M, m, -1, 0, _capi.SUBG_RNG_TRACE, walks=walks)
    r = po.gset_from_walks(q, walks, M, m, -1, want_raw=True)
    _assert_same((nsize, remap, enc, raw), (r["nsize"], r["remap"], r["enc"], r["raw"]), f"trace M={M} m={m}")


def test_trace_random_walks_arbitrary_nodes():
    """Trace mode does not need a graph-consistent trace: random node ids, repeated nodes, bucket."""
    from surel_plus_b200 import DeviceGraph, _capi
    from surel_plus_b200.graphs import synthetic_graph
    A = synthetic_graph(5000, 20000, seed=3)
    rng = np.random.default_rng(0)
    n, M, m = 500, 64, 3
    q = rng.integers(0, 5000, n).astype(np.int32)
    walks = rng.integers(0, 40, (n, M, m)).astype(np.int32)  # tiny id range -> heavy duplication
    walks[::3] = rng.integers(0, 5000, (len(walks[::3]), M, m))
    g = DeviceGraph.from_scipy(A)
    for bucket in (-1, 10):
        _, nsize, remap, enc, raw = _sample(g, A, q, M, m, bucket, 0, _capi.SUBG_RNG_TRACE, walks=walks)
        r = po.gset_from_walks(q, walks, M, m, bucket, want_raw=True)
        _assert_same((nsize, remap, enc, raw), (r["nsize"], r["remap"], r["enc"], r["raw"]), f"bucket={bucket}")


def test_multi_chunk_and_table_growth(mid_graph, monkeypatch):
    """Chunked staging and LP-table regrowth give the same bits as the single-pass run."""
    from surel_plus_b200 import DeviceGraph, _capi
    A = mid_graph
    q = np.arange(A.shape[0])
    g = DeviceGraph.from_scipy(A)
    _, *ref = _sample(g, A, q, 50, 3, -1, 7, _capi.SUBG_RNG_RAND_R)
    monkeypatch.setenv("SUBG_STAGING_BYTES", str(3000 * 152 * 10))
    monkeypatch.setenv("SUBG_LP_TABLE_LOG2", "6")
    _, *got = _sample(g, A, q, 50, 3, -1, 7, _capi.SUBG_RNG_RAND_R)
    _assert_same(got, ref, "chunked")
    exp = po.gset_sampler_replay(A.indptr, A.indices, q, 50, 3, -1, 7, debug=1)
    _assert_same(got, exp, "chunked vs oracle")


def test_philox_invariants(mid_graph):
    """The reference's own test invariants (subg_acc/test/test.py:34-45) on the Philox fast path,
    plus the exact first-hop law (subg_acc.c:790-800)."""
    from surel_plus_b200 import DeviceGraph, _capi
    A = mid_graph
    n = A.shape[0]
    q = np.arange(n)
    M, m = 100, 3
    g = DeviceGraph.from_scipy(A)
    spg, nsize, remap, enc, raw = _sample(g, A, q, M, m, -1, 111413, _capi.SUBG_RNG_PHILOX)
    assert nsize.sum() == remap.shape[1]                                         # test.py:34
    assert (remap.max(axis=1) - [n - 1, enc.shape[0] - 1]).sum() == 0            # test.py:36
    assert (enc[remap[1]][:, 0] == M).sum() == n                                 # test.py:38
    assert np.abs((enc[remap[1]].sum(axis=0) / n - M).sum()) < 1e-10             # test.py:39-40
    assert (raw[:, 0] == M).sum() == n                                           # test.py:43
    assert (enc[remap[1]] - raw).sum() == 0                                      # test.py:44
    assert (raw.max(axis=0) - M).sum() == 0                                      # test.py:45
    # per seed every LP column sums to M; first-hop column = round-robin counts when deg <= M
    off = np.concatenate([[0], np.cumsum(nsize)])
    deg = np.diff(A.indptr)
    seg = np.repeat(np.arange(n), nsize)
    colsum = np.zeros((n, m + 1), np.int64)
    np.add.at(colsum, seg, raw.astype(np.int64))
    assert (colsum[:, 1:] == M).all()
    for u in np.random.default_rng(0).choice(n, 200, replace=False):
        ids, rows = remap[0][off[u]:off[u + 1]], raw[off[u]:off[u + 1]]
        assert ids[0] == u
        d = deg[u]
        if d == 0:
            assert len(ids) == 1 and (rows[0] == M).all()
            continue
        first = dict(zip(ids.tolist(), rows[:, 1].tolist()))
        nb = A.indices[A.indptr[u]:A.indptr[u + 1]]
        if d <= M:
            exp = np.bincount(np.arange(M) % d, minlength=d)
            assert all(first.get(int(v), 0) == int(c) for v, c in zip(nb, exp))
        else:
            hit = [first.get(int(v), 0) for v in nb]
            assert set(hit) <= {0, 1} and sum(hit) == M
    # SpG (sorted CSR-of-sets) agrees with the export
    z = spg.to_scipy()
    assert z.has_sorted_indices
    z2, _ = po.subg_matrix_from(nsize, remap, enc, q, n, m + 1)
    assert (z != z2).nnz == 0
    # same seed -> same bits; different seed -> different sample
    _, n2, r2, e2, _ = _sample(g, A, q, M, m, -1, 111413, _capi.SUBG_RNG_PHILOX)
    assert np.array_equal(remap, r2) and np.array_equal(enc, e2)
    _, n3, r3, _, _ = _sample(g, A, q, M, m, -1, 5, _capi.SUBG_RNG_PHILOX)
    assert r3.shape != remap.shape or not np.array_equal(r3, remap)


def test_empty_and_errors(small_graph):
    from surel_plus_b200 import DeviceGraph, SpG, _capi
    g = DeviceGraph.from_scipy(small_graph)
    spg = SpG.sample(g, np.zeros(0, np.int32), 10, 2)
    assert spg.n == 0 and spg.T == 0
    nsize, remap, enc = spg.export_reference()
    assert nsize.shape == (0,) and remap.shape == (2, 0)
    with pytest.raises(TypeError):
        SpG.sample(g, np.array([10 ** 6], np.int32), 10, 2)        # node id out of range
    with pytest.raises(AssertionError):
        SpG.sample(g, np.arange(4), 30000, 5)                       # 5*15+1 > 64 bits (subg_acc.c:913)
    with pytest.raises(TypeError):
        SpG.sample(g, np.arange(4), 40000, 2)                       # int16 landing counts


def test_subg_acc_module_signature(small_graph):
    """Boundary A: the drop-in module returns the reference's list of numpy arrays."""
    from surel_plus_b200 import subg_acc
    A = small_graph
    q = np.arange(A.shape[0])
    out = subg_acc.gset_sampler(A.indptr, A.indices, q, num_walks=20, num_steps=3, nthread=1, seed=99)
    exp = po.gset_sampler_replay(A.indptr, A.indices, q, 20, 3, -1, 99)
    assert isinstance(out, list) and len(out) == 3
    for a, b in zip(out, exp):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    out4 = subg_acc.gset_sampler(A.indptr, A.indices, q, num_walks=20, num_steps=3, nthread=1, seed=99, debug=1)
    assert len(out4) == 4 and np.array_equal(out4[2][out4[1][1]], out4[3])
    # callers then do remap[1]+1, np.repeat(idx, nsize), np.insert(enc, 0, ...) (random_walks.py:79-81)
    z, enc0 = po.subg_matrix_from(out[0], out[1], out[2], q, A.shape[0], 4)
    assert z.nnz == out[0].sum() and enc0.shape[0] == out[2].shape[0] + 1


def test_subg_matrix_on_a_permuted_subset_of_the_nodes(small_graph):
    """subg_matrix(G, idx) with idx != arange(N): the reference builds an (N, N) matrix with row idx[i] = set i and
    empty rows elsewhere (random_walks.py:79); here the device SpG's row table is expanded (subg_spg_expand_rows).
    Checked against the oracle's replay of the same stream, through the top-level `subg_acc` name as well."""
    import subg_acc as top
    from surel_plus_b200 import _capi, gather, subg_matrix
    A = small_graph
    N, M, K = A.shape[0], 30, 4
    idx = np.random.default_rng(3).permutation(N)[: N // 3].astype(np.int64)
    z, enc0 = subg_matrix(A, idx, num_walks=M, num_steps=K, seed=77, rng_mode=_capi.SUBG_RNG_RAND_R)
    nsize, remap, enc = po.gset_sampler_replay(A.indptr, A.indices, idx, M, K - 1, -1, 77)
    keep = np.repeat(np.diff(A.indptr)[idx] > 0, nsize)
    remap = remap.copy()
    remap[0][~keep] = np.repeat(idx, nsize)[~keep]        # isolated seeds: the reference leaves the id unwritten
    z_ref, enc_ref = po.subg_matrix_from(nsize, remap, enc, idx, N, K)
    assert z.shape == (N, N) and z.n == N
    got = z.to_scipy()
    assert np.array_equal(got.indptr, z_ref.indptr) and np.array_equal(got.indices, z_ref.indices)
    assert np.array_equal(got.data, z_ref.data) and np.array_equal(enc0, enc_ref)
    # joins index rows by node id; nodes outside idx have empty sets
    edge = np.random.default_rng(4).integers(0, N, (2, 300))
    xz, ptr = gather(edge, z, "cuda", True, None)
    want, sl, sr = po.spjoin_pair(z_ref, edge)
    assert np.array_equal(xz.cpu().numpy()[..., 0].astype(np.int64), want)
    assert np.array_equal(ptr.cpu().numpy(), po.pair_index(sl, sr, True))
    # the reference's own call through the top-level module name, then its 3-line CSR build (random_walks.py:77-81)
    os_env = __import__("os").environ
    os_env["SUBG_RNG"] = "rand_r"
    try:
        ns2, rm2, enc2 = top.gset_sampler(A.indptr.astype(np.int32), A.indices.astype(np.int32), idx, num_walks=M, num_steps=K - 1, seed=77)
    finally:
        os_env.pop("SUBG_RNG", None)
    z2, _ = po.subg_matrix_from(ns2, rm2, enc2, idx, N, K)
    assert (z2 != z_ref).nnz == 0
