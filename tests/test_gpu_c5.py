"""GPU: BASELINE.json configs[4] at full size on one B200 -- synthetic twitter-2010 shape (41.65 M nodes, about
1.47 G directed entries), LP M=100, walk length 2 (CLI num_steps=3), every node a seed: more than 2^31 SpG
entries, so every 64-bit offset path is exercised; plus a graph with more than 2^31 directed entries (64-bit row
pointer) replayed against the oracle on a window of seeds.  Graphs are generated and ingested on the device
(subg_graph_from_edges); checks are size-independent properties evaluated on the device in seed chunks."""
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _device_edges(N, E, seed, gamma=2.0):
    gen = torch.Generator(device="cuda:0")
    gen.manual_seed(seed)
    src = (torch.rand(E, dtype=torch.float64, device="cuda:0", generator=gen).pow_(gamma) * N).to(torch.int64).clamp_(max=N - 1)
    dst = torch.randint(0, N, (E,), dtype=torch.int64, device="cuda:0", generator=gen)
    return src, dst


def _need_big_gpu():
    if torch.cuda.get_device_properties(0).total_memory < 150 * 2 ** 30:
        pytest.skip("needs a 180 GB B200")


def _check_chunk(rows, enc, lo, hi, M, m, seeds):
    rb, ns = rows["rowbeg"][lo:hi], rows["nsize"][lo:hi].long()
    off = torch.cumsum(ns, 0) - ns
    tot = int(ns.sum())
    pos = torch.arange(tot, device="cuda:0") - torch.repeat_interleave(off, ns) + torch.repeat_interleave(rb, ns)
    idx = rows["indices"][pos]
    dat = rows["data"][pos].long()
    del pos
    asc = idx[1:] > idx[:-1]
    asc[(off[1:] - 1)] = True
    assert bool(asc.all()), "set not strictly ascending (random_walks.py:80)"
    assert int(dat.min()) >= 1 and int(dat.max()) <= enc.shape[0]
    lp = enc[dat - 1]
    root = lp[:, 0] == M
    assert int(root.sum()) == hi - lo                                   # test.py:38
    assert torch.equal(idx[root].long(), seeds[lo:hi].long())          # the root of a set is its seed
    ends = off + ns - 1
    for col in range(1, m + 1):
        cs = torch.cumsum(lp[:, col].long(), 0)
        per = cs[ends] - torch.cat([cs.new_zeros(1), cs[ends[:-1]]])
        assert bool((per == M).all()), f"LP column {col} does not sum to M"   # test.py:39-40
    return tot


def test_c5_twitter_shape_full_size_one_gpu():
    _need_big_gpu()
    from surel_plus_b200 import DeviceGraph, SpG, gather
    from surel_plus_b200.graphs import SHAPES
    N, E_und, gseed = SHAPES["twitter"]
    M, m = 100, 2
    t0 = time.perf_counter()
    src, dst = _device_edges(N, E_und, gseed)
    g = DeviceGraph.from_edges(src, dst, num_nodes=N, symmetrize=True, drop_self_loops=True, device="cuda:0")
    del src, dst
    torch.cuda.synchronize()
    t_graph = time.perf_counter() - t0
    assert g.N == N and 1.3e9 < g.E <= 2 * E_und
    q = torch.arange(N, dtype=torch.int32, device="cuda:0")
    SpG.sample(g, q[:1_000_000], num_walks=M, num_steps=m, seed=1, first_visit_ranks=False).close()   # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    spg = SpG.sample(g, q, num_walks=M, num_steps=m, seed=111413, first_visit_ranks=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"\n[c5] graph N={g.N} E={g.E} built in {t_graph:.1f}s; sampled {N} seeds in {ms:.1f} ms = {N / ms * 1e3:.3e} seeds/s; "
          f"T={spg.T} c={spg.c} max_set={spg.max_set}")
    assert spg.n == N and spg.T > 2 ** 31 and spg.max_set <= M * m + 1
    rows = spg.rows()
    assert int(rows["nsize"].long().sum()) == spg.T                      # test.py:34
    enc = torch.from_numpy(spg.enc_table()[1:].astype(np.int32)).cuda()
    assert int(enc.max()) == M                                           # test.py:45
    tot, chunk = 0, 4_000_000
    for lo in range(0, N, chunk):
        tot += _check_chunk(rows, enc, lo, min(lo + chunk, N), M, m, q)
    assert tot == spg.T
    # pair SpJoin straight on the scattered > 2^31-entry layout
    B = 21504
    rng = np.random.default_rng(3)
    edge = torch.from_numpy(rng.integers(0, N, (2, B))).cuda()
    edge[0, :8] = N - 1 - torch.arange(8, device="cuda:0")               # rows that live beyond entry 2^31
    xz, ptr = gather(edge, spg, "cuda:0", True, None)
    ns = rows["nsize"].long()
    sizes = torch.cat([ns[edge[0]], ns[edge[1]]])
    assert torch.equal(ptr[1:] - ptr[:-1], sizes) and xz.shape[0] == int(sizes.sum())
    own = []
    for u in edge[0, :64].tolist():
        b = int(rows["rowbeg"][u])
        own.append(rows["data"][b:b + int(ns[u])])
    own = torch.cat(own).float()
    assert torch.equal(xz[:own.numel(), 0, 0], own)
    nl = int(ns[edge[0]].sum())
    assert int((xz[:nl, 1, 0] > 0).sum()) == int((xz[nl:, 1, 0] > 0).sum())   # |S_u & S_v| from both sides
    spg.close()
    g.close()


def test_int64_row_pointer_graph_replays_against_oracle():
    """More than 2^31 directed entries: 64-bit row pointer in HBM; seeds whose rows lie beyond entry 2^31 are
    replayed bit-exactly against the oracle on the exported CSR."""
    _need_big_gpu()
    from oracle import pyoracle as po
    from surel_plus_b200 import DeviceGraph, SpG, _capi
    N, E_und = 12_000_000, 1_120_000_000
    src, dst = _device_edges(N, E_und, 77, gamma=1.5)
    g = DeviceGraph.from_edges(src, dst, num_nodes=N, symmetrize=True, drop_self_loops=True, device="cuda:0")
    del src, dst
    assert g.E >= 2 ** 31, g.E
    indptr, indices = g.csr()
    assert indptr.dtype == np.int64 and indptr[-1] == g.E and np.all(np.diff(indptr) >= 0)
    q = np.concatenate([np.arange(N - 600, N), np.arange(0, 40), np.arange(N // 2, N // 2 + 360)]).astype(np.int32)
    assert indptr[q[0]] > 2 ** 31
    spg = SpG.sample(g, q, num_walks=100, num_steps=2, seed=9, rng_mode=_capi.SUBG_RNG_RAND_R)
    got = spg.export_reference()
    exp = po.gset_sampler_replay(indptr, indices, q, 100, 2, -1, 9)
    for a, b, name in zip(got, exp, ("nsize", "remap", "enc")):
        assert np.array_equal(a, b), name
    spg.close()
    g.close()
