"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (seed partitioning, merge of
the per-shard LP tables into global first-occurrence order, variable-size all-gather, CSR re-basing).
The per-rank shard is produced by the oracle here (the CUDA sampler needs a GPU); the assembled SpG
must equal the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle as po
from surel_plus_b200 import parallel as par


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _local_shard(nsize, remap, enc, lo, hi):
    """Cut seeds [lo, hi) out of a whole-query result and re-number its LP rows locally (first occurrence
    inside the shard), sorted per set as the device SpG stores them."""
    off = np.concatenate([[0], np.cumsum(nsize)])
    a, b = off[lo], off[hi]
    nodes, ids = remap[0, a:b], remap[1, a:b]
    uniq, first = np.unique(ids, return_index=True)
    order = uniq[np.argsort(first)]                  # global ids in local first-occurrence order
    local_of = {g: i for i, g in enumerate(order.tolist())}
    lid = np.array([local_of[g] for g in ids.tolist()], np.int32)
    ns = nsize[lo:hi]
    o2 = np.concatenate([[0], np.cumsum(ns)])
    idx_sorted = np.concatenate([np.argsort(nodes[o2[i]:o2[i + 1]], kind="stable") + o2[i] for i in range(len(ns))]) \
        if len(ns) else np.zeros(0, np.int64)
    return ns.astype(np.int32), nodes[idx_sorted].astype(np.int32), (lid[idx_sorted] + 1).astype(np.int32), enc[order]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from surel_plus_b200.graphs import synthetic_graph
        A = synthetic_graph(605, 3000, seed=1, gamma=3.0, isolated=5)
        n = A.shape[0]
        q = np.arange(n)
        M, m = 30, 3
        nsize, remap, enc = po.gset_sampler_replay(A.indptr, A.indices, q, M, m, -1, 99)
        for bounds in (None, par.partition_by_work(np.minimum(np.diff(A.indptr), M) + M * (m - 1.0), world)):
            lo, hi = par.partition(n, world, rank) if bounds is None else (int(bounds[rank]), int(bounds[rank + 1]))
            ns, ind, dat, tab = _local_shard(nsize, remap, enc, lo, hi)
            parts = par.assemble_shards(torch.from_numpy(ns), torch.from_numpy(ind), torch.from_numpy(dat),
                                        torch.from_numpy(np.ascontiguousarray(tab)))
            indptr, indices, data = po.spg_build(n, q, nsize, remap[0], remap[1])
            assert np.array_equal(parts["indptr"].numpy(), indptr)
            assert np.array_equal(parts["indices"].numpy(), indices)
            assert np.array_equal(parts["data"].numpy(), data)
            assert np.array_equal(parts["enc"], enc)
            assert parts["bytes_gathered"] >= 8 * len(indices)
        # ragged / empty shard: rank 1 contributes nothing
        e = np.zeros(0, np.int32)
        if rank == 0:
            ns, ind, dat, tab = _local_shard(nsize, remap, enc, 0, n)
        else:
            ns, ind, dat, tab = e, e, e, np.zeros((0, m + 1), np.int16)
        parts = par.assemble_shards(torch.from_numpy(ns), torch.from_numpy(ind), torch.from_numpy(dat),
                                    torch.from_numpy(np.ascontiguousarray(tab)))
        assert parts["indices"].numel() == remap.shape[1] and np.array_equal(parts["enc"], enc)
        # host-side decisions of the CUDA exchange (parallel.sharded_sample / ShardExchange) that every rank has to
        # take identically: the peer-mapping vote, and the slab size derived from the largest seed range
        assert par.agree(True, "cpu") is True
        assert par.agree(rank == 0, "cpu") is False          # one rank could not map its peers -> all fall back
        assert par.agree(False, "cpu") is False
        bounds = par.partition_by_work(np.minimum(np.diff(A.indptr), M) + M * (m - 1.0), world)
        need = par.slab_bytes_needed(int(np.max(np.diff(bounds))), M * m + 1)
        t = torch.tensor([need], dtype=torch.int64)
        both = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(both, t)
        assert all(int(x) == need for x in both)               # same slab size everywhere without communicating
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        ns, ind, dat, tab = _local_shard(nsize, remap, enc, lo, hi)
        assert need >= 8 * len(ind) + 12 * len(ns) + 16 * len(tab)   # the widest wire format of this rank's shard fits
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


def test_assemble_shards_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: 1, 1: 1}


def test_partition_covers_everything():
    for n in (0, 1, 7, 576289):
        for world in (1, 2, 3, 8):
            edges = [par.partition(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            assert max(hi - lo for lo, hi in edges) - min(hi - lo for lo, hi in edges) <= 1
    b = par.partition_by_work(np.array([5.0, 1, 1, 1, 1, 1]), 2)
    assert b.tolist() == [0, 1, 6]


def test_merge_lp_tables_first_occurrence():
    t0 = np.array([[5, 0, 0], [0, 1, 2], [0, 3, 0]], np.int16)
    t1 = np.array([[0, 3, 0], [5, 0, 0], [0, 9, 9]], np.int16)
    merged, maps = par.merge_lp_tables([t0, t1, np.zeros((0, 3), np.int16)])
    assert merged.tolist() == [[5, 0, 0], [0, 1, 2], [0, 3, 0], [0, 9, 9]]
    assert maps[0].tolist() == [0, 1, 2] and maps[1].tolist() == [2, 0, 3] and maps[2].tolist() == []
