"""GPU: the multi-GPU building blocks.  (1) shards of a range-partitioned query sampled with
subg_gset_sample_shard + LP-table merge + relabel concatenate to exactly the single-call SpG (one GPU
is enough for that); (2) the full exchange over NCCL with one process per GPU, when >= 2 GPUs exist."""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard(graph, q, lo, hi, M, m, seed, mode):
    from surel_plus_b200 import SpG, _capi
    from surel_plus_b200.spg import _ptr, _stream
    h = C.c_void_p()
    _capi.check(_capi.load().subg_gset_sample_shard(graph._h, _ptr(q), q.size, lo, hi, M, m, -1, seed, mode, None,
                                                     0, _stream(graph.device), C.byref(h)))
    return SpG(h, graph.device, n_nodes=graph.N, num_walks=M)


@pytest.mark.parametrize("mode_name", ["RAND_R", "PHILOX"])
@pytest.mark.parametrize("cuts", [(0.5,), (0.1, 0.35, 0.9)])
def test_shards_concatenate_to_single_call(mid_graph, mode_name, cuts):
    from surel_plus_b200 import DeviceGraph, SpG, _capi
    from surel_plus_b200 import parallel as par
    from surel_plus_b200.spg import _ptr, _stream
    A = mid_graph
    n = A.shape[0]
    M, m, seed = 50, 3, 7
    mode = getattr(_capi, f"SUBG_RNG_{mode_name}")
    q = np.random.default_rng(1).permutation(n).astype(np.int32)
    g = DeviceGraph.from_scipy(A)
    full = SpG.sample(g, q, M, m, seed=seed, rng_mode=mode)
    fv = full.views()
    bounds = [0] + [int(c * n) for c in cuts] + [n]
    shards = [_shard(g, q, bounds[i], bounds[i + 1], M, m, seed, mode) for i in range(len(bounds) - 1)]
    tables = [s.views()["enc"].cpu().numpy() for s in shards]
    merged, maps = par.merge_lp_tables(tables)
    assert np.array_equal(merged, fv["enc"].cpu().numpy())
    lib = _capi.load()
    for s, mp_ in zip(shards, maps):
        _capi.check(lib.subg_spg_set_lp_table(s._h, _ptr(mp_), _ptr(np.ascontiguousarray(merged)), merged.shape[0], -1,
                                              _stream(g.device)))
    for key in ("indices", "data", "nsize"):
        cat = torch.cat([s.views()[key] for s in shards])
        assert torch.equal(cat, fv[key]), key
    if mode_name == "RAND_R":  # and that equals the reference's single stream (oracle replay)
        nsize, remap, enc = po.gset_sampler_replay(A.indptr, A.indices, q, M, m, -1, seed)
        assert np.array_equal(fv["nsize"].cpu().numpy(), nsize) and np.array_equal(merged, enc)


def _exchange_in_one_process(g, shards, M, m):
    """world = len(shards) exchange contexts in ONE process on one GPU: every shard is packed into its own slab and
    every 'rank' assembles the full SpG reading the slabs through `srcs` (what the NCCL-staged mode does; the peer
    mode differs only in where the same kernels read from).  Returns the full SpG of every rank."""
    from surel_plus_b200 import SpG, _capi
    from surel_plus_b200 import parallel as par
    from surel_plus_b200.spg import _stream
    lib = _capi.load()
    W = len(shards)
    need = max(par.slab_bytes_needed(s.n, M * m + 1, lp_rows=max(s.c, 1)) for s in shards)
    ctx = []
    for r in range(W):
        h = C.c_void_p()
        _capi.check(lib.subg_xchg_create(g.device, r, W, need, C.byref(h)))
        ctx.append(h)
    headers = np.zeros((W, 8), np.int64)
    slabs = []
    for r in range(W):
        _capi.check(lib.subg_xchg_pack(ctx[r], shards[r]._h, g.N, headers[r].ctypes.data, _stream(g.device)))
        p = C.c_void_p()
        _capi.check(lib.subg_xchg_slab(ctx[r], C.byref(p), None))
        slabs.append(p.value)
    assert (headers[:, par.H_FMT] >= 0).all()
    srcs = (C.c_void_p * W)(*slabs)
    fulls = []
    for r in range(W):
        fh = C.c_void_p()
        _capi.check(lib.subg_xchg_assemble(ctx[r], headers.ctypes.data, srcs, M, m + 1, _stream(g.device), C.byref(fh)))
        fulls.append(SpG(fh, g.device, n_nodes=g.N, num_walks=M))
    torch.cuda.synchronize()
    for h in ctx:
        lib.subg_xchg_free(h)
    return fulls, headers


@pytest.mark.parametrize("extra", [0, 1, 2, 4])
@pytest.mark.parametrize("mode_name,cuts", [("RAND_R", (0.5,)), ("PHILOX", (0.1, 0.35, 0.9)), ("PHILOX", (0.0, 0.5))])
def test_exchange_pack_merge_pull(mid_graph, mode_name, cuts, extra, monkeypatch):
    """The device-side exchange (csrc/xchg.cu): pack -> LP-table merge -> pull rebuilds exactly the single-call SpG
    (rows, LP ids, LP table: global first-occurrence order), in every wire format, with an empty shard, and the
    scattered result joins like the single-GPU one."""
    from surel_plus_b200 import DeviceGraph, SpG, _capi, gather
    from surel_plus_b200 import parallel as par
    A = mid_graph
    n = A.shape[0]
    M, m, seed = 50, 3, 7
    if extra:
        monkeypatch.setenv("SUBG_XCHG_EXTRA", str(extra))
    mode = getattr(_capi, f"SUBG_RNG_{mode_name}")
    q = np.random.default_rng(1).permutation(n).astype(np.int32)
    g = DeviceGraph.from_scipy(A)
    full = SpG.sample(g, q, M, m, seed=seed, rng_mode=mode, first_visit_ranks=False)
    bounds = [0] + [int(c * n) for c in cuts] + [n]
    shards = []
    for i in range(len(bounds) - 1):
        h = C.c_void_p()
        from surel_plus_b200.spg import _ptr, _stream
        _capi.check(_capi.load().subg_gset_sample_shard(g._h, _ptr(q), q.size, bounds[i], bounds[i + 1], M, m, -1, seed, mode,
                                                         None, _capi.SAMPLE_NO_RANKS, _stream(g.device), C.byref(h)))
        shards.append(SpG(h, g.device, n_nodes=g.N, num_walks=M))
    reps, headers = _exchange_in_one_process(g, shards, M, m)
    assert set((headers[:, par.H_FMT] & 0xff).tolist()) <= {extra} or extra == 0
    xpe = torch.from_numpy(full.enc_table()).float().cuda() / M
    edge = np.random.default_rng(0).integers(0, n, (2, 700))
    want_xz, want_ptr = gather(edge, full, "cuda", True, xpe)
    for rep in reps:
        assert (rep.n, rep.T, rep.c, rep.max_set) == (full.n, full.T, full.c, full.max_set)
        assert np.array_equal(rep.enc_table(), full.enc_table())
        got_xz, got_ptr = gather(edge, rep, "cuda", True, xpe)      # joins the scattered layout in place
        assert torch.equal(got_xz, want_xz) and torch.equal(got_ptr, want_ptr)
    fv = full.views()
    for rep in reps:
        rv = rep.views()
        for key in ("indptr", "indices", "data"):
            assert torch.equal(rv[key], fv[key]), key
    if mode_name == "RAND_R":  # and that is the reference's single stream (oracle replay)
        nsize, remap, enc = po.gset_sampler_replay(A.indptr, A.indices, q, M, m, -1, seed)
        assert np.array_equal(reps[0].enc_table()[1:], enc)
        assert np.array_equal(reps[0].views()["nsize"].cpu().numpy(), nsize)


@pytest.mark.parametrize("mode_name,cuts", [("RAND_R", (0.5,)), ("PHILOX", (0.1, 0.35, 0.9)), ("PHILOX", (0.0, 0.5))])
def test_linked_shards_join_like_the_single_call_spg(mid_graph, mode_name, cuts):
    """The no-replication alternative (subg_xchg_stage / subg_xchg_link): the shards stay in their slabs, the LP tables are
    merged, every 'rank' relabels its own id plane, and the linked SpG -- rows addressed by 64-bit offsets that reach into
    the other slabs -- joins (pairs, triplets, JoinStream) exactly like the SpG of a single call; views() pulls the rows
    into a local compact copy equal to the single call's CSR.  One process, one GPU, world = number of shards (on a
    multi-GPU box the slabs are the peers' IPC mappings: tests/multi_gpu_worker.py)."""
    from surel_plus_b200 import DeviceGraph, JoinStream, SpG, _capi, gather, hgather
    from surel_plus_b200 import parallel as par
    from surel_plus_b200.spg import _ptr, _stream
    A = mid_graph
    n = A.shape[0]
    M, m, seed = 50, 3, 7
    mode = getattr(_capi, f"SUBG_RNG_{mode_name}")
    q = np.random.default_rng(1).permutation(n).astype(np.int32)
    g = DeviceGraph.from_scipy(A)
    lib = _capi.load()
    full = SpG.sample(g, q, M, m, seed=seed, rng_mode=mode, first_visit_ranks=False)
    bounds = [0] + [int(c * n) for c in cuts] + [n]
    W = len(bounds) - 1
    shards = []
    for i in range(W):
        h = C.c_void_p()
        _capi.check(lib.subg_gset_sample_shard(g._h, _ptr(q), q.size, bounds[i], bounds[i + 1], M, m, -1, seed, mode,
                                               None, _capi.SAMPLE_NO_RANKS | _capi.SAMPLE_NO_COMPACT, _stream(g.device), C.byref(h)))
        shards.append(SpG(h, g.device, n_nodes=g.N, num_walks=M))
    row_cap = (M * m + 1 + 3) & ~3
    n_max = max(s.n for s in shards)
    need = par.slab_bytes_needed(n_max, row_cap, lp_rows=max(max(s.c for s in shards), 1)) + (1 << 16)
    ctx, slabs = [], []
    headers = np.zeros((W, 8), np.int64)
    for r in range(W):
        h = C.c_void_p()
        _capi.check(lib.subg_xchg_create(g.device, r, W, need, C.byref(h)))
        ctx.append(h)
        _capi.check(lib.subg_xchg_stage(h, shards[r]._h, g.N, n_max * row_cap, headers[r].ctypes.data, _stream(g.device)))
        p = C.c_void_p()
        _capi.check(lib.subg_xchg_slab(h, C.byref(p), None))
        slabs.append(p.value)
    assert (headers[:, par.H_FMT] >= 0).all() and set((headers[:, par.H_FMT] & 0xff).tolist()) == {8}
    srcs = (C.c_void_p * W)(*slabs)
    linked = []
    for r in range(W):
        fh = C.c_void_p()
        _capi.check(lib.subg_xchg_link(ctx[r], headers.ctypes.data, srcs, M, m + 1, _stream(g.device), C.byref(fh)))
        linked.append(SpG(fh, g.device, n_nodes=g.N, num_walks=M))
    torch.cuda.synchronize()          # "barrier": every id plane is relabelled
    for s in shards:
        s.close()
    xpe = torch.from_numpy(full.enc_table()).float().cuda() / M
    rng = np.random.default_rng(0)
    edge = rng.integers(0, n, (2, 700))
    hedge = rng.integers(0, n, (3, 300))
    want_xz, want_ptr = gather(edge, full, "cuda", True, xpe)
    want_raw, _ = gather(edge, full, "cuda", True, None)
    want_h, want_hseg = hgather(hedge, full, "cuda", xpe)
    for rep in linked:
        assert (rep.n, rep.T, rep.c, rep.max_set) == (full.n, full.T, full.c, full.max_set)
        assert np.array_equal(rep.enc_table(), full.enc_table())
        got_xz, got_ptr = gather(edge, rep, "cuda", True, xpe)
        assert torch.equal(got_xz, want_xz) and torch.equal(got_ptr, want_ptr)
        got_raw, _ = gather(edge, rep, "cuda", True, None)
        assert torch.equal(got_raw, want_raw)                       # the LP ids themselves, not only their table rows
        got_h, got_hseg = hgather(hedge, rep, "cuda", xpe)
        assert torch.equal(got_h, want_h) and torch.equal(got_hseg, want_hseg)
    js = JoinStream(linked[0], 700, "cuda", encode=xpe)
    sxz, sptr = js.gather(edge)
    assert torch.equal(sxz, want_xz) and torch.equal(sptr, want_ptr)
    js.close()
    fv = full.views()
    rv = linked[-1].views()            # pulls the rows into an owned compact copy
    for key in ("indptr", "indices", "data"):
        assert torch.equal(rv[key], fv[key]), key
    got_xz, _ = gather(edge, linked[-1], "cuda", True, xpe)
    assert torch.equal(got_xz, want_xz)
    for rep in linked:
        rep.close()
    torch.cuda.synchronize()
    for h in ctx:
        lib.subg_xchg_free(h)
    full.close()
    g.close()


def test_from_device_csr_roundtrip(mid_graph):
    from surel_plus_b200 import DeviceGraph, SpG, gather
    A = mid_graph
    g = DeviceGraph.from_scipy(A)
    spg = SpG.sample(g, np.arange(A.shape[0]), 40, 2, seed=3)
    v = spg.views()
    twin = SpG.from_device_csr(v["indptr"], v["indices"], v["data"], n_nodes=A.shape[0], enc=v["enc"], num_walks=40)
    assert twin.T == spg.T and twin.c == spg.c and twin.max_set == spg.max_set
    assert np.array_equal(twin.enc_table(), spg.enc_table())
    xpe = torch.from_numpy(spg.enc_table()).float().cuda() / 40
    edge = np.random.default_rng(0).integers(0, A.shape[0], (2, 500))
    a, pa = gather(edge, spg, "cuda", True, xpe)
    b, pb = gather(edge, twin, "cuda", True, xpe)
    assert torch.equal(a, b) and torch.equal(pa, pb)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_sample_nccl_two_gpus():
    """One process per GPU over NCCL: replicated SpG equals the single-GPU one on every rank."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "multi_gpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("SHARDED_OK") == world
