"""GPU parity: subg_batch_sample (one warp replaying the reference's serial rand_r stream) vs the oracle restatement of
batch_sampler (subg_acc.c:391-507), which tests/test_oracle_vs_reference.py pins against the compiled reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po


@pytest.mark.parametrize("M,m,thld,nq", [(200, 8, 1000, 50), (20, 3, 100, 30), (5, 4, 400, 200), (300, 2, 50, 10),
                                         (10, 1, 600, 605), (3, 5, 10 ** 6, 40)])
def test_batch_sampler_bit_exact(small_graph, M, m, thld, nq):
    from surel_plus_b200 import subg_acc
    A = small_graph
    ptr, nb = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    q = np.random.default_rng(M + nq).permutation(A.shape[0])[:nq].astype(np.int32)
    got = subg_acc.batch_sampler(ptr, nb, q, num_walks=M, num_steps=m, thld=thld, seed=7, pid=4242)
    want = po.batch_sampler(ptr, nb, q, M, m, thld, seed=7, pid=4242)
    assert got.dtype == np.int32 and np.array_equal(got, want)
    assert len(set(got.tolist())) == len(got)                      # distinct nodes, insertion order
    assert set(q.tolist()) <= set(got.tolist())                    # every query node joins the batch (subg_acc.c:443)


def test_batch_sampler_mid_graph_and_errors(mid_graph):
    from surel_plus_b200 import DeviceGraph, subg_acc
    A = mid_graph
    g = DeviceGraph.from_scipy(A)
    q = np.random.default_rng(1).permutation(A.shape[0])[:512].astype(np.int32)
    got = subg_acc.batch_sampler(g, None, q, num_walks=50, num_steps=4, thld=5000, seed=3, pid=1)
    want = po.batch_sampler(A.indptr, A.indices, q, 50, 4, 5000, seed=3, pid=1)
    assert np.array_equal(got, want)
    assert len(subg_acc.batch_sampler(g, None, np.zeros(0, np.int32), pid=1)) == 0
    with pytest.raises(TypeError):
        subg_acc.batch_sampler(g, None, np.array([A.shape[0] + 3], np.int32), pid=1)
    g.close()
