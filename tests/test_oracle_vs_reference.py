"""CPU, this container only (`reference` marker: needs /root/reference and oracle/_ref): the oracle
against the UNMODIFIED reference run live on larger seeded inputs than the committed fixtures."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import pyoracle as po
from oracle import reference as ref

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def subg():
    m = ref.subg_acc()
    if m is None:
        pytest.skip("oracle/_ref/subg_acc*.so not built (make -C oracle ref)")
    return m


def _mask_isolated(remap, nsize, deg_of_seed, seeds):
    remap = remap.copy()
    off = np.concatenate([[0], np.cumsum(nsize)])
    iso = np.where(deg_of_seed == 0)[0]
    remap[0, off[iso]] = seeds[iso]  # the reference leaves this entry unwritten (subg_acc.c:753-761)
    return remap


@pytest.mark.parametrize("M,m,bucket,seed", [(200, 2, -1, 111413), (200, 3, -1, 5), (100, 2, -1, 1), (64, 4, 40, 9),
                                             (1, 1, -1, 2), (300, 2, -1, 3)])
def test_gset_replay_vs_compiled_reference(subg, mid_graph, M, m, bucket, seed, capfd):
    A = mid_graph
    rng = np.random.default_rng(seed)
    q = rng.permutation(A.shape[0])[:1500].astype(np.int32)
    q[:3] = [A.shape[0] - 1, A.shape[0] - 2, 0]  # isolated seeds + the hub
    indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    nsize, remap, enc, raw = subg.gset_sampler(indptr, indices, q, num_walks=M, num_steps=m, bucket=bucket, nthread=1,
                                               seed=seed, debug=1)
    capfd.readouterr()
    remap = _mask_isolated(remap, nsize, np.diff(indptr)[q], q)
    o = po.gset_sampler_replay(indptr, indices, q, M, m, bucket, seed, debug=1)
    assert np.array_equal(o[0], nsize) and np.array_equal(o[1], remap)
    assert np.array_equal(o[2], enc) and np.array_equal(o[3], raw)


def test_key_width_assertion_matches(subg, small_graph, capfd):
    A = small_graph
    indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    q = np.arange(8, dtype=np.int32)
    with pytest.raises(AssertionError):
        subg.gset_sampler(indptr, indices, q, num_walks=30000, num_steps=5, nthread=1)
    capfd.readouterr()
    with pytest.raises(AssertionError):
        po.gset_sampler_replay(indptr, indices, q, 30000, 5)


@pytest.mark.parametrize("B", [1, 33, 1024])
def test_spjoin_vs_reference_train(subg, mid_graph, B, capfd):
    import torch
    train = ref.train()
    A = mid_graph
    n = A.shape[0]
    q = np.arange(n, dtype=np.int32)
    M, m = 60, 3
    indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    nsize, remap, enc = subg.gset_sampler(indptr, indices, q, num_walks=M, num_steps=m, nthread=1, seed=4)
    capfd.readouterr()
    remap = _mask_isolated(remap, nsize, np.diff(indptr), q)
    z, enc0 = po.subg_matrix_from(nsize, remap, enc, q, n, m + 1)
    xpe = torch.from_numpy(enc0).float() / M
    rng = np.random.default_rng(B)
    edge = rng.integers(0, n, (2, B))
    xz, sl, sr = po.spjoin_pair(z, edge)
    rxz, rptr = train.gather(edge, z, "cpu", True, xpe)
    assert torch.equal(xpe[torch.from_numpy(xz).long()], rxz)
    assert np.array_equal(po.pair_index(sl, sr, True), rptr.numpy())
    _, rind = train.gather(edge, z, "cpu", False, xpe)
    assert np.array_equal(po.pair_index(sl, sr, False), rind.numpy())
    if B >= 4:
        pxz, pptr = train.pgather(edge, z, "cpu", xpe, train.bgather, True, 4)
        assert torch.equal(pxz, rxz) and torch.equal(pptr, rptr)
    hedge = rng.integers(0, n, (3, B))
    hxz, sizes = po.spjoin_triplet(z, hedge)
    rhxz, rhind = train.hgather(hedge, z, "cpu", xpe)
    assert torch.equal(xpe[torch.from_numpy(hxz).long()], rhxz)
    assert np.array_equal(np.repeat(np.arange(4 * B), sizes), rhind.numpy())


def test_ppr_push_vs_numba(mid_graph):
    import numba
    pprgo = ref.pprgo()
    A = mid_graph.astype(np.float32)
    deg = np.sum(A > 0, axis=1).A1
    rng = np.random.default_rng(0)
    seeds = np.concatenate([[0, A.shape[0] - 1], rng.integers(0, A.shape[0], 60)])
    for alpha, eps in ((0.1, 1e-4), (0.15, 1e-3), (0.5, 1e-5)):
        for s in seeds:
            k, v = pprgo._calc_ppr_node(int(s), A.indptr, A.indices, deg, numba.float32(alpha), numba.float32(eps))
            ok, ov, _ = po.ppr_push(A.indptr, A.indices, deg.astype(np.int64), int(s), alpha, eps)
            assert np.array_equal(ok, np.array(k, np.int64))
            assert np.array_equal(ov.view(np.uint32), np.array(v, np.float32).view(np.uint32))


def test_topk_ppr_matrix_vs_reference(mid_graph):
    pprgo = ref.pprgo()
    A = mid_graph.astype(np.float32)
    idx = np.arange(0, A.shape[0], 7)
    for norm in ("sym", "row"):
        exp = pprgo.topk_ppr_matrix(A, 0.1, 1e-4, idx, 50, normalization=norm).tocsr()
        got = po.topk_ppr_matrix(A, 0.1, 1e-4, idx, 50, norm)
        exp.sort_indices(); got.sort_indices()
        assert np.array_equal(np.diff(exp.indptr), np.diff(got.indptr))
        d = (exp != got)
        # rows may only differ where the k-th score is tied
        assert d.nnz <= 0.02 * exp.nnz
        common = exp.multiply(got > 0)
        common2 = got.multiply(exp > 0)
        assert np.array_equal(common.tocsr().data, common2.tocsr().data)


@pytest.mark.parametrize("M,m,rep,seed", [(200, 3, True, 111413), (200, 3, -1, 5), (100, 2, True, 1), (1, 1, -1, 2),
                                          (300, 2, True, 3), (17, 6, -1, 4)])
def test_walk_sampler_vs_compiled_reference(subg, mid_graph, M, m, rep, seed):
    """SUREL-v1 walk_sampler (subg_acc.c:144-389): oracle == compiled reference with nthread=1."""
    A = mid_graph
    rng = np.random.default_rng(seed)
    q = rng.permutation(A.shape[0])[:800].astype(np.int32)
    q[:3] = [A.shape[0] - 1, A.shape[0] - 2, 0]
    indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    kw = {} if rep == -1 else {"replacement": rep}
    walks, obj = subg.walk_sampler(indptr, indices, q, num_walks=M, num_steps=m, nthread=1, seed=seed, **kw)
    o_walks, o_obj = po.walk_sampler(indptr, indices, q, M, m, seed, rep)
    assert np.array_equal(walks, o_walks)
    for i in range(len(q)):
        assert np.array_equal(obj[i, 0], o_obj[i, 0]) and np.array_equal(obj[i, 1], o_obj[i, 1]), i


@pytest.mark.parametrize("M,m,rep", [(50, 3, True), (20, 2, -1)])
def test_walk_join_vs_compiled_reference(subg, mid_graph, M, m, rep):
    A = mid_graph
    rng = np.random.default_rng(M)
    q = rng.permutation(A.shape[0])[:700].astype(np.int32)
    indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    kw = {} if rep == -1 else {"replacement": rep}
    walks, obj = subg.walk_sampler(indptr, indices, q, num_walks=M, num_steps=m, nthread=1, seed=1, **kw)
    qq = q[rng.integers(0, len(q), (400, 2))].astype(np.int32)
    out, xq = subg.walk_join(walks, list(obj[:, 0]), qq, return_idx=True)
    o_out, o_xq = po.walk_join(walks, list(obj[:, 0]), qq, return_idx=True)
    assert np.array_equal(out, o_out) and np.array_equal(xq, o_xq)


@pytest.mark.parametrize("M,m,thld,nq", [(200, 8, 1000, 50), (20, 3, 100, 30), (5, 4, 400, 200), (300, 2, 50, 10),
                                         (10, 1, 600, 605), (3, 5, 10 ** 6, 40)])
def test_batch_sampler_oracle_equals_compiled_reference(small_graph, M, m, thld, nq):
    """batch_sampler (subg_acc.c:391-507) seeds its stream with seed + getpid(): called in THIS process with the same
    seed, the unmodified compiled reference and the oracle restatement return the same nodes in the same order
    (hubs with more than num_walks neighbours, isolated nodes, thresholds that never / always cut the walks short)."""
    import os
    R = ref.subg_acc()
    A = small_graph
    ptr, nb = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    q = np.random.default_rng(M + nq).permutation(A.shape[0])[:nq].astype(np.int32)
    want = R.batch_sampler(ptr, nb, q, num_walks=M, num_steps=m, thld=thld, seed=7)
    got = po.batch_sampler(ptr, nb, q, M, m, thld, seed=7, pid=os.getpid())
    assert np.array_equal(got, want)
