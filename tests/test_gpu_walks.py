"""GPU parity: SUREL-v1 walk_sampler (walks + relative-position encoder, subg_acc.c:144-389) vs the
oracle (bit-exact in the rand_r replay mode), the committed reference fixtures, and law-level
properties of the Philox path."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _flat(obj):
    n = obj.shape[0]
    sizes = np.array([len(obj[i, 0]) for i in range(n)], np.int64)
    ids = np.concatenate([obj[i, 0] for i in range(n)]) if n else np.zeros(0, np.int32)
    rpe = np.vstack([obj[i, 1] for i in range(n)]) if n else np.zeros((0, 1), np.int32)
    return np.concatenate([[0], np.cumsum(sizes)]), ids, rpe


@pytest.mark.parametrize("M,m,rep", [(20, 3, -1), (20, 3, True), (50, 2, True), (7, 4, -1), (200, 3, True),
                                     (200, 3, -1), (33, 1, True), (33, 1, -1), (100, 5, True), (1, 1, -1)])
def test_walk_sampler_replay_bit_exact(small_graph, M, m, rep):
    from surel_plus_b200 import subg_acc
    A = small_graph
    q = np.arange(A.shape[0], dtype=np.int32)
    walks, obj = subg_acc.walk_sampler(A.indptr, A.indices, q, num_walks=M, num_steps=m, nthread=1, seed=99, replacement=rep)
    e_walks, e_obj = po.walk_sampler(A.indptr, A.indices, q, M, m, 99, rep)
    assert walks.dtype == np.int32 and walks.shape == e_walks.shape
    assert np.array_equal(walks, e_walks), np.argwhere(walks != e_walks)[:5].tolist()
    assert obj.shape == (len(q), 2) and obj.dtype == object
    for a, b in zip(_flat(obj), _flat(e_obj)):
        assert a.dtype == b.dtype and np.array_equal(a, b)


def test_walk_sampler_matches_reference_fixtures(small_graph):
    """Against tests/golden/walks.npz, written by the unmodified compiled reference."""
    from surel_plus_b200 import subg_acc
    gold = np.load(os.path.join(GOLD, "walks.npz"))
    A = small_graph
    q_all = np.arange(A.shape[0], dtype=np.int32)
    for ci, (M, m, without, seed) in enumerate(gold["walk_cases"].tolist()):
        q = gold["walk_subset_query"] if ci == 2 else q_all
        walks, obj = subg_acc.walk_sampler(A.indptr, A.indices, q, num_walks=M, num_steps=m, nthread=1, seed=seed,
                                           replacement=True if without else -1)
        off, ids, rpe = _flat(obj)
        assert np.array_equal(walks, gold[f"walk{ci}_walks"]), ci
        assert np.array_equal(off, gold[f"walk{ci}_off"]) and np.array_equal(ids, gold[f"walk{ci}_ids"]), ci
        assert np.array_equal(rpe, gold[f"walk{ci}_rpe"]), ci


def test_walk_sampler_larger_graph_subset_query(mid_graph):
    from surel_plus_b200 import subg_acc
    A = mid_graph
    rng = np.random.default_rng(3)
    q = rng.permutation(A.shape[0])[:2500].astype(np.int32)
    q[:3] = [A.shape[0] - 1, A.shape[0] - 2, 0]
    for rep in (True, -1):
        walks, obj = subg_acc.walk_sampler(A.indptr, A.indices, q, num_walks=200, num_steps=3, nthread=1, seed=5, replacement=rep)
        e_walks, e_obj = po.walk_sampler(A.indptr, A.indices, q, 200, 3, 5, rep)
        assert np.array_equal(walks, e_walks)
        for a, b in zip(_flat(obj), _flat(e_obj)):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("rep", [True, -1])
def test_walk_sampler_philox_properties(mid_graph, rep):
    """Philox path: every hop follows an edge (or stays on a node without neighbours), the first hop
    without replacement is the reference's round-robin / distinct-neighbour law (subg_acc.c:220-233),
    and the encoder applied to these walks equals the oracle's encoder on the same walks."""
    from surel_plus_b200 import subg_acc
    A = mid_graph
    M, m = 100, 3
    q = np.arange(0, A.shape[0], 7, dtype=np.int32)
    walks, obj = subg_acc.walk_sampler(A.indptr, A.indices, q, num_walks=M, num_steps=m, nthread=-1, seed=1, replacement=rep)
    W = walks.reshape(len(q), M, m + 1)
    assert np.array_equal(W[:, :, 0], np.broadcast_to(q[:, None], (len(q), M)))
    deg = np.diff(A.indptr)
    src, dst = W[:, :, :-1].reshape(-1), W[:, :, 1:].reshape(-1)
    has = np.asarray(A[src, dst]).reshape(-1) > 0
    assert np.all(has | ((deg[src] == 0) & (src == dst)))
    if rep is True:
        for i in np.where(deg[q] > 0)[0][:400]:
            u, d = q[i], deg[q[i]]
            first = W[i, :, 1]
            if d <= M:
                assert np.array_equal(first, A.indices[A.indptr[u] + np.arange(M) % d])
            else:
                assert len(np.unique(first)) == M
    off, ids, rpe = po.rpe_encode(walks, M, m)
    g_off, g_ids, g_rpe = _flat(obj)
    assert np.array_equal(off, g_off) and np.array_equal(ids, g_ids) and np.array_equal(rpe, g_rpe)
    # two different seeds give different walks, the same seed the same walks
    w2, _ = subg_acc.walk_sampler(A.indptr, A.indices, q, num_walks=M, num_steps=m, nthread=-1, seed=1, replacement=rep)
    w3, _ = subg_acc.walk_sampler(A.indptr, A.indices, q, num_walks=M, num_steps=m, nthread=-1, seed=2, replacement=rep)
    assert np.array_equal(walks, w2) and not np.array_equal(walks, w3)


def test_walk_sampler_errors_and_empty(small_graph):
    from surel_plus_b200 import subg_acc
    from surel_plus_b200._capi import SubgUnsupported
    A = small_graph
    walks, obj = subg_acc.walk_sampler(A.indptr, A.indices, np.zeros(0, np.int32), num_walks=10, num_steps=2)
    assert walks.shape == (0, 30) and obj.shape == (0, 2)
    with pytest.raises(TypeError):
        subg_acc.walk_sampler(A.indptr, A.indices, np.array([A.shape[0]], np.int32), num_walks=10, num_steps=2)
    with pytest.raises(SubgUnsupported):
        subg_acc.walk_sampler(A.indptr, A.indices, np.array([0], np.int32), num_walks=5000, num_steps=4)


def test_rw_matrix_matches_reference_formulation(small_graph):
    """rw_matrix (sampler/random_walks.py:58-73) over the device walk_sampler == the same formulation
    over the oracle's walk_sampler (single batch so that the rand_r stream is one call)."""
    import scipy.sparse as sp
    from surel_plus_b200.sampler import rw_matrix
    A = small_graph
    n, M, K = A.shape[0], 50, 3
    idx = np.arange(n)
    z, freqs = rw_matrix(A, idx, num_walks=M, num_steps=K, batch_size=n, nthread=1, seed=7)
    _, obj = po.walk_sampler(A.indptr, A.indices, idx, M, K - 1, 7, True)
    fr = np.vstack([obj[i, 1] for i in range(n)])
    proj = np.array([(M + 1) ** i for i in reversed(range(K))], dtype=np.int64)
    val, first, inv = np.unique(fr.astype(np.int64) @ proj, return_index=True, return_inverse=True)
    rows = np.repeat(idx, [len(obj[i, 0]) for i in range(n)])
    ez = sp.csr_matrix((inv + 1, (rows, np.concatenate([obj[i, 0] for i in range(n)]))), (n, n))
    assert (z != ez).nnz == 0
    assert np.array_equal(freqs, np.insert(fr[first], 0, np.zeros((1, K)), axis=0))


def test_walk_join_bit_exact(mid_graph):
    """SUREL-v1 walk_join (subg_acc.c:509-647) on the device == oracle, incl. u == v, 3-D walk input and a query
    node that is not a root (-1 row)."""
    from surel_plus_b200 import subg_acc
    A = mid_graph
    rng = np.random.default_rng(2)
    q = rng.permutation(A.shape[0])[:900].astype(np.int32)
    M, m = 40, 3
    walks, obj = subg_acc.walk_sampler(A.indptr, A.indices, q, num_walks=M, num_steps=m, nthread=1, seed=4, replacement=True)
    qq = q[rng.integers(0, len(q), (500, 2))].astype(np.int32)
    qq[0] = [q[0], q[0]]
    out, xq = subg_acc.walk_join(walks, list(obj[:, 0]), qq, return_idx=True)
    e_out, e_xq = po.walk_join(walks, list(obj[:, 0]), qq, return_idx=True)
    assert out.dtype == np.int32 and out.shape == (2, 500 * 2 * M * (m + 1))
    assert np.array_equal(out, e_out) and np.array_equal(xq, e_xq)
    out3 = subg_acc.walk_join(walks.reshape(len(q), M, m + 1), list(obj[:, 0]), qq)
    assert isinstance(out3, np.ndarray) and np.array_equal(out3, e_out)
    # a node that is no root: documented -1 entries, identical in the oracle
    missing = np.setdiff1d(np.arange(A.shape[0]), q)[:1].astype(np.int32)
    qm = np.array([[q[1], missing[0]], [missing[0], q[2]]], np.int32)
    om, xm = subg_acc.walk_join(walks, list(obj[:, 0]), qm, return_idx=True)
    eo, ex = po.walk_join(walks, list(obj[:, 0]), qm, return_idx=True)
    assert np.array_equal(om, eo) and np.array_equal(xm, ex) and xm[0, 1] == -1 and xm[1, 0] == -1
    with pytest.raises(AssertionError):
        subg_acc.walk_join(walks, list(obj[:-1, 0]), qq)


def test_walk_join_matches_reference_fixture(small_graph):
    from surel_plus_b200 import subg_acc
    gold = np.load(os.path.join(GOLD, "walks.npz"))
    for ci in (1, 2):
        walks, off, ids = gold[f"walk{ci}_walks"], gold[f"walk{ci}_off"], gold[f"walk{ci}_ids"]
        keys = [ids[off[i]:off[i + 1]] for i in range(len(off) - 1)]
        out, xq = subg_acc.walk_join(walks, keys, gold[f"walk{ci}_join_query"], return_idx=True)
        assert np.array_equal(out, gold[f"walk{ci}_join_out"]) and np.array_equal(xq, gold[f"walk{ci}_join_xq"])
