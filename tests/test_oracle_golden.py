"""CPU: the oracle (oracle/subg_oracle.c + oracle/pyoracle.py) against the committed fixtures that
tests/golden/make_golden.py generated from the UNMODIFIED reference.  Pins the oracle on the GPU
box too, where /root/reference does not exist."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gset():
    return np.load(os.path.join(GOLD, "gset.npz"))


@pytest.fixture(scope="module")
def spjoin():
    return np.load(os.path.join(GOLD, "spjoin.npz"))


@pytest.fixture(scope="module")
def ppr():
    return np.load(os.path.join(GOLD, "ppr.npz"))


def test_fixture_graph_is_the_conftest_graph(gset, small_graph):
    assert np.array_equal(gset["graph_indptr"], small_graph.indptr)
    assert np.array_equal(gset["graph_indices"], small_graph.indices)


@pytest.mark.parametrize("seed", [111413, 99])
def test_rand_r_matches_glibc(gset, seed):
    assert np.array_equal(po.rand_r_stream(seed, 64), gset[f"rand_r_{seed}"])


def test_gset_replay_equals_reference_nthread1(gset):
    """orc_walks_rand_r + orc_gset_from_walks == reference gset_sampler(nthread=1), every array."""
    indptr, indices = gset["graph_indptr"], gset["graph_indices"]
    q = np.arange(len(indptr) - 1)
    for ci, (M, m, bucket, seed) in enumerate(gset["gset_cases"].tolist()):
        nsize, remap, enc, raw = po.gset_sampler_replay(indptr, indices, q, M, m, bucket, seed, debug=1)
        for got, name in ((nsize, "nsize"), (remap, "remap"), (enc, "enc"), (raw, "raw")):
            exp = gset[f"gset{ci}_{name}"]
            assert got.dtype == exp.dtype and got.shape == exp.shape, (ci, name)
            assert np.array_equal(got, exp), (ci, name)


def test_gset_replay_permuted_subset_query(gset):
    indptr, indices = gset["graph_indptr"], gset["graph_indices"]
    nsize, remap, enc = po.gset_sampler_replay(indptr, indices, gset["gsetq_query"], 30, 2, -1, 11)
    assert np.array_equal(nsize, gset["gsetq_nsize"])
    assert np.array_equal(remap, gset["gsetq_remap"])
    assert np.array_equal(enc, gset["gsetq_enc"])


def test_reference_test_invariants_hold_on_fixture(gset):
    """subg_acc/test/test.py:34-45 on the stored reference output."""
    n = len(gset["graph_indptr"]) - 1
    for ci, (M, m, bucket, seed) in enumerate(gset["gset_cases"].tolist()):
        nsize, remap, enc, raw = (gset[f"gset{ci}_{k}"] for k in ("nsize", "remap", "enc", "raw"))
        assert nsize.sum() == remap.shape[1]
        assert remap[1].max() == enc.shape[0] - 1
        assert (enc[remap[1]][:, 0] == M).sum() == n
        assert np.array_equal(enc[remap[1]], raw)
        if bucket < 0:
            assert remap[0].max() == n - 1
            assert np.allclose(raw.sum(axis=0) / n, M)


def _z(spjoin):
    n = len(spjoin["spg_indptr"]) - 1
    return sp.csr_matrix((spjoin["spg_data"], spjoin["spg_indices"], spjoin["spg_indptr"]), shape=(n, n))


def test_spg_build_c_equals_scipy(gset):
    """orc_spg_build == scipy's COO->CSR of random_walks.py:79."""
    n = len(gset["graph_indptr"]) - 1
    q = np.arange(n)
    nsize, remap, enc = gset["gset1_nsize"], gset["gset1_remap"], gset["gset1_enc"]
    z, enc0 = po.subg_matrix_from(nsize, remap, enc, q, n, enc.shape[1])
    indptr, indices, data = po.spg_build(n, q, nsize, remap[0], remap[1])
    assert np.array_equal(indptr, z.indptr) and np.array_equal(indices, z.indices) and np.array_equal(data, z.data)
    assert enc0.shape[0] == enc.shape[0] + 1 and not enc0[0].any()


def test_pair_join_equals_reference_gather(spjoin):
    z = _z(spjoin)
    xpe = spjoin["spg_enc0"].astype(np.float32) / 50
    edge = spjoin["pair_edge"]
    xz, sl, sr = po.spjoin_pair(z, edge)
    assert np.array_equal(xpe[xz], spjoin["pair_xz"])                                   # train.py:37
    assert np.array_equal(po.pair_index(sl, sr, True), spjoin["pair_ptr"])              # train.py:20-22
    assert np.array_equal(po.pair_index(sl, sr, False), spjoin["pair_ind"])             # train.py:24-30
    assert np.array_equal(xz.astype(np.float32)[..., None], spjoin["pair_xz_noenc"])    # train.py:39-43
    nl = int(sl.sum())
    assert np.array_equal(xz[:nl], spjoin["pair_bg_xl"]) and np.array_equal(xz[nl:], spjoin["pair_bg_xr"])
    assert np.array_equal(sl, spjoin["pair_bg_sl"]) and np.array_equal(sr, spjoin["pair_bg_sr"])
    # the scipy restatement used as the SpJoin CPU baseline computes the same rows
    left, right, nl2, nr2 = po.scipy_pair_join(edge, z)
    assert np.array_equal(np.vstack([left, right]), xz)
    # ... and so does the threaded pgather restatement (train.py:88-111, njobs = 4), incl. its segment pointer
    xz4, ptr4 = po.scipy_pgather(edge, z, njobs=4)
    assert np.array_equal(xz4, xz) and np.array_equal(ptr4, spjoin["pair_ptr"])


def test_triplet_join_equals_reference_hgather(spjoin):
    z = _z(spjoin)
    xpe = spjoin["spg_enc0"].astype(np.float32) / 50
    hedge = spjoin["trip_edge"]
    xz, sizes = po.spjoin_triplet(z, hedge)
    assert np.array_equal(xpe[xz], spjoin["trip_xz"])
    assert np.array_equal(np.repeat(np.arange(4 * hedge.shape[1]), sizes), spjoin["trip_ind"])
    xz2, ind2 = po.scipy_triplet_join(hedge, z)     # the scipy restatement timed as the triplet CPU baseline
    assert np.array_equal(xz2, xz) and np.array_equal(ind2, spjoin["trip_ind"])


def _adj(gset):
    n = len(gset["graph_indptr"]) - 1
    return sp.csr_matrix((np.ones(len(gset["graph_indices"]), np.int64), gset["graph_indices"], gset["graph_indptr"]),
                         shape=(n, n))


def test_ppr_push_bit_identical(gset, ppr):
    """orc_ppr_push == numba _calc_ppr_node: same support, same insertion order, same float32 bits."""
    A = _adj(gset)
    deg = np.diff(A.indptr).astype(np.int64)
    alpha, eps, _ = ppr["ppr_params"]
    for si in range(5):
        s = int(ppr[f"push{si}_seed"])
        keys, vals, _ = po.ppr_push(A.indptr, A.indices, deg, s, alpha, eps)
        assert np.array_equal(keys, ppr[f"push{si}_keys"])
        assert np.array_equal(vals.view(np.uint32), ppr[f"push{si}_vals"].view(np.uint32))


@pytest.mark.parametrize("norm", ["row", "sym", "col"])
def test_topk_ppr_matrix(gset, ppr, norm):
    """Top-k sets equal except for entries tied with the k-th score (unstable argsort, pprgo.py:59);
    values bit-identical on the common support."""
    A = _adj(gset)
    alpha, eps, topk = ppr["ppr_params"]
    mat = po.topk_ppr_matrix(A, alpha, eps, np.arange(A.shape[0]), int(topk), norm)
    mat.sort_indices()
    n = A.shape[0]
    exp = sp.csr_matrix((ppr[f"ppr_{norm}_data"], ppr[f"ppr_{norm}_indices"], ppr[f"ppr_{norm}_indptr"]), shape=(n, n))
    assert np.array_equal(np.diff(mat.indptr), np.diff(exp.indptr))
    same_rows = 0
    for u in range(n):
        a = dict(zip(mat.indices[mat.indptr[u]:mat.indptr[u + 1]].tolist(), mat.data[mat.indptr[u]:mat.indptr[u + 1]].tolist()))
        b = dict(zip(exp.indices[exp.indptr[u]:exp.indptr[u + 1]].tolist(), exp.data[exp.indptr[u]:exp.indptr[u + 1]].tolist()))
        common = set(a) & set(b)
        assert all(a[w] == b[w] for w in common), (u, norm)
        if set(a) == set(b):
            same_rows += 1
        else:  # only k-th-score ties may differ: the odd ones out share one raw score per row
            assert len(set(a) - set(b)) == len(set(b) - set(a))
    assert same_rows >= 0.95 * n


def test_encoders(gset, ppr):
    A = _adj(gset)
    n = A.shape[0]
    x = sp.csr_matrix((ppr["ppr_sym_data"], ppr["ppr_sym_indices"], ppr["ppr_sym_indptr"]), shape=(n, n))
    xp = po.encoding_ppr(x)
    assert np.array_equal(xp.indices, ppr["enc_ppr_indices"]) and np.array_equal(xp.data, ppr["enc_ppr_data"])
    xs = po.encoding_spd(x, A)
    assert np.array_equal(xs.indptr, ppr["enc_spd_indptr"]) and np.array_equal(xs.indices, ppr["enc_spd_indices"])
    assert np.array_equal(xs.data, ppr["enc_spd_data"])


def test_value_join_equals_reference_gather(ppr, gset):
    n = len(gset["graph_indptr"]) - 1
    xs = sp.csr_matrix((ppr["enc_spd_data"], ppr["enc_spd_indices"], ppr["enc_spd_indptr"]), shape=(n, n))
    xz, sl, sr = po.spjoin_pair(xs, ppr["spd_edge"])
    assert np.array_equal(xz.astype(np.float32)[..., None], ppr["spd_xz"])
    assert np.array_equal(po.pair_index(sl, sr, True), ppr["spd_ptr"])


# ------------------------------------------------------------------ SUREL-v1 walk_sampler
@pytest.fixture(scope="module")
def walks_gold():
    return np.load(os.path.join(GOLD, "walks.npz"))


def test_walk_sampler_equals_reference_nthread1(gset, walks_gold):
    """orc_walk_sampler_walks + orc_rpe_encode == reference walk_sampler(nthread=1) (subg_acc.c:144-389)."""
    indptr, indices = gset["graph_indptr"], gset["graph_indices"]
    q_all = np.arange(len(indptr) - 1)
    for ci, (M, m, without, seed) in enumerate(walks_gold["walk_cases"].tolist()):
        q = walks_gold["walk_subset_query"] if ci == 2 else q_all
        walks, obj = po.walk_sampler(indptr, indices, q, M, m, seed, True if without else -1)
        assert walks.dtype == np.int32 and np.array_equal(walks, walks_gold[f"walk{ci}_walks"]), ci
        off = walks_gold[f"walk{ci}_off"]
        assert np.array_equal(np.concatenate([obj[i, 0] for i in range(len(q))]), walks_gold[f"walk{ci}_ids"]), ci
        assert np.array_equal(np.vstack([obj[i, 1] for i in range(len(q))]), walks_gold[f"walk{ci}_rpe"]), ci
        assert np.array_equal(np.cumsum([len(obj[i, 0]) for i in range(len(q))]), off[1:]), ci


def test_rpe_invariants_on_golden(walks_gold):
    """Column sums: every step column of a seed's rpe sums to M; entry [0][0] = M (subg_acc.c:294-303)."""
    for ci, (M, m, without, seed) in enumerate(walks_gold["walk_cases"].tolist()):
        off, rpe = walks_gold[f"walk{ci}_off"], walks_gold[f"walk{ci}_rpe"]
        seg = np.add.reduceat(rpe, off[:-1], axis=0)
        assert np.all(seg == M), ci
        assert np.all(rpe[off[:-1], 0] == M), ci


def test_walk_join_equals_reference(walks_gold):
    """orc_walk_join == reference walk_join (subg_acc.c:509-647) on the reference's own walks and key sets."""
    for ci in (1, 2):
        walks, off, ids = walks_gold[f"walk{ci}_walks"], walks_gold[f"walk{ci}_off"], walks_gold[f"walk{ci}_ids"]
        keys = [ids[off[i]:off[i + 1]] for i in range(len(off) - 1)]
        out, xq = po.walk_join(walks, keys, walks_gold[f"walk{ci}_join_query"], return_idx=True)
        assert out.dtype == np.int32 and np.array_equal(out, walks_gold[f"walk{ci}_join_out"]), ci
        assert np.array_equal(xq, walks_gold[f"walk{ci}_join_xq"]), ci


def test_rand_r_low_bits_cycle():
    """Documents why the Philox path is checked against the sampling LAW and not against the reference stream
    alone ('Parity of the fast path'): glibc rand_r is a 32-bit LCG whose output modulo a power of two is
    state bits 16..18 of every third step -- short cycles, so consecutive draws are far MORE evenly spread than
    independent draws would be (chi-square of the pair histogram ~ 0 instead of ~ its degrees of freedom)."""
    from scipy.stats import chi2
    x = po.rand_r_stream(111413, 300_000)
    for a in (4, 8, 16):
        cnt = np.bincount((x[:-1] % a) * a + (x[1:] % a), minlength=a * a).astype(float)
        exp = (len(x) - 1) / (a * a)
        stat = ((cnt - exp) ** 2 / exp).sum()
        assert chi2.cdf(stat, a * a - 1) < 2e-3, (a, stat)      # under-dispersed: an independent stream fails this
    cnt = np.bincount((x[:-1] % 11) * 13 + (x[1:] % 13), minlength=143).astype(float)   # other moduli look fine
    stat = ((cnt - (len(x) - 1) / 143) ** 2 / ((len(x) - 1) / 143)).sum()
    assert 0.001 < chi2.cdf(stat, 142) < 0.999
