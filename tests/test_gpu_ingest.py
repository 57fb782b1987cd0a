"""GPU: data formats either side of the path -- device edge-list -> CSR ingestion
against scipy's csr_matrix (what edge2csr does, subg_acc/test/test.py:15-19), text edge lists, and the
scipy .npz layout of the SpG (main.py:184-202)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

pytestmark = pytest.mark.gpu


def _scipy_csr(row, col, N, symmetrize, drop_self):
    if drop_self:
        keep = row != col
        row, col = row[keep], col[keep]
    if symmetrize:
        row, col = np.concatenate([row, col]), np.concatenate([col, row])
    A = sp.csr_matrix((np.ones(len(row), dtype=bool), (row, col)), shape=(N, N))   # test.py:17-19
    A.sum_duplicates()
    A.sort_indices()
    return A


@pytest.mark.parametrize("N,E,sym,drop", [(50, 400, False, False), (50, 400, True, True), (1000, 20000, True, False),
                                          (100000, 500000, False, True), (7, 3, False, False), (300, 0, True, True)])
@pytest.mark.parametrize("where", ["host", "device"])
def test_from_edges_equals_scipy(N, E, sym, drop, where):
    from surel_plus_b200 import DeviceGraph
    rng = np.random.default_rng(N + E)
    row = rng.integers(0, max(N - 3, 1), E)          # the last ids stay isolated (trailing empty rows)
    col = rng.integers(0, max(N - 3, 1), E)
    if E > 10:
        row[:5], col[:5] = row[5:10], col[5:10]      # duplicates
        col[10] = row[10]                            # a self loop
    r, c = (torch.from_numpy(row).cuda(), torch.from_numpy(col).cuda()) if where == "device" else (row, col)
    g = DeviceGraph.from_edges(r, c, num_nodes=N, symmetrize=sym, drop_self_loops=drop, device="cuda:0")
    A = _scipy_csr(row, col, N, sym, drop)
    indptr, indices = g.csr()
    assert g.N == N and g.E == A.nnz
    assert np.array_equal(indptr, A.indptr) and np.array_equal(indices, A.indices)
    G = g.to_scipy()
    assert (G != A).nnz == 0
    g.close()


def test_from_edges_infers_size_and_rejects_bad_ids():
    from surel_plus_b200 import DeviceGraph
    g = DeviceGraph.from_edges(np.array([0, 5, 2]), np.array([9, 1, 2]), device="cuda:0")
    assert g.N == 10 and g.E == 3                     # nmax + 1 (test.py:18)
    g.close()
    with pytest.raises(TypeError):
        DeviceGraph.from_edges(np.array([0, -1]), np.array([1, 1]), device="cuda:0")
    with pytest.raises(TypeError):
        DeviceGraph.from_edges(np.array([0, 7]), np.array([1, 1]), num_nodes=5, device="cuda:0")


def test_ingested_graph_samples_like_uploaded_graph(mid_graph):
    """A graph built on the device from its edge list gives the same SpG as the uploaded CSR (rand_r replay)."""
    from surel_plus_b200 import DeviceGraph, SpG, _capi
    A = mid_graph.tocoo()
    g1 = DeviceGraph.from_edges(A.row, A.col, num_nodes=A.shape[0], device="cuda:0")
    g2 = DeviceGraph.from_scipy(mid_graph, "cuda:0")
    q = np.arange(0, A.shape[0], 5, dtype=np.int32)
    outs = []
    for g in (g1, g2):
        s = SpG.sample(g, q, num_walks=50, num_steps=3, seed=3, rng_mode=_capi.SUBG_RNG_RAND_R)
        outs.append(s.export_reference())
        s.close()
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_edge2csr_text_file(tmp_path):
    """test.py:15-19 on a whitespace-separated edge list with duplicate lines."""
    from surel_plus_b200.io import edge2csr
    rng = np.random.default_rng(0)
    e = rng.integers(0, 500, (4000, 2))
    e[:50] = e[50:100]
    path = tmp_path / "test.edgelist"
    np.savetxt(path, e, fmt="%d")
    g = edge2csr(str(path), device="cuda:0")
    row, col = np.loadtxt(path, dtype=int).T
    nmax = max(row.max(), col.max())
    A = sp.csr_matrix((np.ones(len(row), dtype=bool), (row, col)), shape=(nmax + 1, nmax + 1))
    A.sum_duplicates(); A.sort_indices()
    indptr, indices = g.csr()
    assert np.array_equal(indptr, A.indptr) and np.array_equal(indices, A.indices)


def test_spg_npz_round_trip(tmp_path, small_graph):
    """SpG -> .npz -> scipy.sparse.load_npz and back (main.py:187,202), LP pointers and PPR values."""
    from surel_plus_b200 import DeviceGraph, SpG, topk_ppr_matrix
    from surel_plus_b200.io import load_npz, save_npz
    A = small_graph
    g = DeviceGraph.from_scipy(A, "cuda:0")
    s = SpG.sample(g, np.arange(A.shape[0]), num_walks=30, num_steps=2, seed=1)
    ref = s.to_scipy()
    save_npz(tmp_path / "lp", s)
    back = sp.load_npz(tmp_path / "lp.npz")
    assert back.dtype == ref.dtype and (back != ref).nnz == 0 and np.array_equal(back.indices, ref.indices)
    s2 = load_npz(tmp_path / "lp.npz", "cuda:0")
    assert (s2.to_scipy() != ref).nnz == 0
    z = topk_ppr_matrix(A.astype(np.int64), 0.1, 1e-4, np.arange(A.shape[0]), 16, normalization="sym")
    zs = z.to_scipy() if isinstance(z, SpG) else z
    save_npz(tmp_path / "ppr", z)
    zb = sp.load_npz(tmp_path / "ppr.npz")
    assert zb.dtype == np.float64 and np.array_equal(zb.data, zs.data) and np.array_equal(zb.indices, zs.indices)
    z2 = load_npz(tmp_path / "ppr.npz", "cuda:0")
    assert z2.value_kind == 1 and np.array_equal(z2.to_scipy().data, zs.data)
    sp.save_npz(tmp_path / "ref_written.npz", zs)                  # a file the reference wrote
    z3 = load_npz(tmp_path / "ref_written.npz", "cuda:0")
    assert np.array_equal(z3.to_scipy().data, zs.data)
