"""Multi-GPU SubGAcc: one process per GPU (torch.distributed), graph replicated in every GPU's
HBM, seeds range-partitioned, SpG shards all-gathered once (NCCL over NVLink/NVSwitch), queries
then joined locally on the replicated SpG.  The reference has no
distributed code at all; this is the B200 scaling path of the same operators.

Exchange step (the only collectives on the path; `sharded_sample`):
  1. one small equal-size all-gather with every rank's sizes (n_r, T_r, c_r, status) and its unique LP table
     (c_r x ncol int16, KBs) -> every rank merges the tables in rank order, which reproduces the
     first-occurrence order of the single-process scan (subg_acc.c:957-978) because shards are contiguous seed
     ranges; local LP ids are re-labelled on the device (subg_spg_set_lp_table);
  2. the full SpG is allocated once (subg_spg_alloc) and nsize / indices / data of every peer (8 B per set entry)
     are received straight into their slices in ONE NCCL group of sends and receives; subg_spg_seal derives the
     row pointer on the device.
`assemble_shards` is the same exchange on plain tensors with staged all-gathers (any backend).
The host-side logic below is device-agnostic (it is exercised with gloo on CPU tensors in
tests/test_parallel_gloo.py); the sampling itself is CUDA only.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _capi


# ------------------------------------------------------------------------------ partitioning
def partition(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous seed range [lo, hi) of `rank`: the first n % world ranks get one extra seed."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def partition_by_work(weights: np.ndarray, world: int) -> np.ndarray:
    """Contiguous ranges balanced by a per-seed work estimate (e.g. min(deg, M) + M*(m-1)):
    returns the world+1 range boundaries."""
    cum = np.concatenate([[0], np.cumsum(weights, dtype=np.float64)])
    targets = cum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(cum, targets, side="left")
    return np.concatenate([[0], cuts, [len(weights)]]).astype(np.int64)


# ------------------------------------------------------------------------------ LP table merge
def merge_lp_tables(tables: List[np.ndarray]) -> Tuple[np.ndarray, List[np.ndarray]]:
    """tables[r]: int16 [c_r, ncol] unique LP rows of shard r in local first-occurrence order.
    Returns (merged int16 [c, ncol] in global first-occurrence order, [id_map_r int32 [c_r]])."""
    ncol = tables[0].shape[1] if tables else 0
    cat = np.ascontiguousarray(np.concatenate(tables, axis=0), dtype=np.int16) if tables else np.zeros((0, 0), np.int16)
    if cat.shape[0] == 0:
        return cat.reshape(0, ncol), [np.zeros(0, np.int32) for _ in tables]
    keys = cat.view(np.dtype((np.void, cat.dtype.itemsize * ncol))).ravel()
    _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # unique rows by first occurrence in rank order
    rank_of_unique = np.empty(len(order), np.int32)
    rank_of_unique[order] = np.arange(len(order), dtype=np.int32)
    gid = rank_of_unique[inv.ravel()]
    merged = cat[np.sort(first)]
    maps, off = [], 0
    for t in tables:
        maps.append(np.ascontiguousarray(gid[off:off + t.shape[0]], dtype=np.int32))
        off += t.shape[0]
    return merged, maps


# ------------------------------------------------------------------------------ collectives
def _group_info(group):
    return dist.get_world_size(group), dist.get_rank(group), dist.get_backend(group)


def all_gather_sizes(vals: List[int], device, group=None) -> np.ndarray:
    """[world, len(vals)] int64 matrix of every rank's values."""
    world, _, _ = _group_info(group)
    t = torch.tensor(vals, dtype=torch.int64, device=device)
    out = torch.empty(world * len(vals), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.view(world, len(vals)).cpu().numpy()


def all_gather_varlen(local: torch.Tensor, counts: np.ndarray, out: torch.Tensor, group=None, async_op: bool = False):
    """Concatenate the ranks' `local` (rank r contributes counts[r] leading-dim rows) into `out`.
    NCCL: the output slices are handed to all_gather directly (uneven sizes become grouped
    ncclBroadcasts that write in place, no staging copy).  Other backends (gloo in the CPU tests):
    padded all-gather, then compaction."""
    world, rank, backend = _group_info(group)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    assert local.shape[0] == counts[rank] and out.shape[0] == offs[-1]
    if backend == "nccl":
        # bytes on the wire: NCCL has no int16 (the LP table), and a byte view costs nothing.  Every rank sends its
        # shard to every peer and receives the peers' shards straight into their slices of `out` (grouped
        # ncclSend/ncclRecv over NVLink: 614 GB/s per GPU at world 2 where the uneven list all-gather, which torch
        # turns into one broadcast per rank, reaches 356 GB/s -- scripts/nccl_probe.py); the own slice is a local copy.
        ob = out.reshape(out.shape[0], -1).view(torch.uint8) if out.shape[0] else out.reshape(0, 1).view(torch.uint8)
        lb = local.contiguous().reshape(local.shape[0], -1).view(torch.uint8) if local.shape[0] else ob[:0]
        if counts[rank]:
            ob[offs[rank]:offs[rank + 1]].copy_(lb)
        ops = []
        for r in range(world):
            if r == rank:
                continue
            peer = r if group is None else dist.get_global_rank(group, r)
            if counts[rank]:
                ops.append(dist.P2POp(dist.isend, lb, peer, group))
            if counts[r]:
                ops.append(dist.P2POp(dist.irecv, ob[offs[r]:offs[r + 1]], peer, group))
        works = dist.batch_isend_irecv(ops) if ops else []
        if async_op:
            return works
        for w in works:
            w.wait()
        return out
    # bytes on the wire (gloo has no int16): rows -> uint8 [rows, bytes_per_row]
    per_row = int(np.prod(local.shape[1:], dtype=np.int64))
    width = per_row * local.element_size()
    rows = (local.contiguous().reshape(local.shape[0], per_row).view(torch.uint8) if local.shape[0]
            else torch.zeros((0, width), dtype=torch.uint8, device=local.device))
    mx = int(max(counts.max(), 1))
    pad = torch.zeros((mx, width), dtype=torch.uint8, device=local.device)
    pad[:rows.shape[0]] = rows
    buf = torch.empty((world * mx, width), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    out_rows = out.reshape(out.shape[0], per_row).view(torch.uint8) if out.shape[0] else None
    for r in range(world):
        if counts[r]:
            out_rows[offs[r]:offs[r + 1]] = buf[r * mx:r * mx + int(counts[r])]
    return out


_LP_INLINE_ROWS = 4096  # unique LP rows of a shard that travel with the sizes in the first (small) all-gather


def _p2p_ops(local: torch.Tensor, counts: np.ndarray, out: torch.Tensor, group=None) -> list:
    """Send / recv operations that put every rank's `local` rows into its slice of `out` (byte views; the own slice
    is copied locally).  The caller batches the operations of several arrays into ONE NCCL group."""
    world, rank, _ = _group_info(group)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    ob = out.reshape(out.shape[0], -1).view(torch.uint8) if out.shape[0] else out.reshape(0, 1).view(torch.uint8)
    lb = local.contiguous().reshape(local.shape[0], -1).view(torch.uint8) if local.shape[0] else ob[:0]
    if counts[rank]:
        ob[offs[rank]:offs[rank + 1]].copy_(lb)
    ops = []
    for r in range(world):
        if r == rank:
            continue
        peer = r if group is None else dist.get_global_rank(group, r)
        if counts[rank]:
            ops.append(dist.P2POp(dist.isend, lb, peer, group))
        if counts[r]:
            ops.append(dist.P2POp(dist.irecv, ob[offs[r]:offs[r + 1]], peer, group))
    return ops


def assemble_shards(nsize: torch.Tensor, indices: torch.Tensor, data: torch.Tensor, enc: torch.Tensor,
                    relabel: Optional[Callable[[np.ndarray, np.ndarray], torch.Tensor]] = None, group=None) -> dict:
    """The exchange step on plain tensors.  Inputs are this rank's shard: nsize int32 [n_r],
    indices int32 [T_r] (ascending per set), data int32 [T_r] (local LP id + 1), enc int16 [c_r, ncol].
    `relabel(id_map, merged_enc)` must return the shard's data with global ids (+1); default: torch indexing.
    Returns dict(indptr int64 [n+1], indices, data, enc int16 [c, ncol], nsize, counts=[world,3], bytes)."""
    world, rank, _ = _group_info(group)
    dev = indices.device
    ncol = enc.shape[1]
    counts = all_gather_sizes([nsize.shape[0], indices.shape[0], enc.shape[0]], dev, group)
    n_r, T_r, c_r = counts[:, 0], counts[:, 1], counts[:, 2]
    # (2) unique LP tables
    enc_all = torch.empty((int(c_r.sum()), ncol), dtype=torch.int16, device=dev)
    all_gather_varlen(enc.contiguous(), c_r, enc_all, group)
    enc_np = enc_all.cpu().numpy()
    offs = np.concatenate([[0], np.cumsum(c_r)])
    merged, maps = merge_lp_tables([enc_np[offs[r]:offs[r + 1]] for r in range(world)])
    if relabel is not None:
        data = relabel(maps[rank], merged)
    elif data.numel():
        m = torch.from_numpy(maps[rank]).to(dev)
        data = (m[(data - 1).long()] + 1).to(torch.int32)
    # (3) the shards themselves
    n, T = int(n_r.sum()), int(T_r.sum())
    g_nsize = torch.empty(n, dtype=torch.int32, device=dev)
    g_indices = torch.empty(T, dtype=torch.int32, device=dev)
    g_data = torch.empty(T, dtype=torch.int32, device=dev)
    all_gather_varlen(nsize, n_r, g_nsize, group)
    all_gather_varlen(indices, T_r, g_indices, group)
    all_gather_varlen(data, T_r, g_data, group)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(g_nsize, 0, out=indptr[1:])
    wire = 8 * T + 4 * n + 2 * ncol * int(c_r.sum())
    return {"indptr": indptr, "indices": g_indices, "data": g_data, "enc": merged, "nsize": g_nsize,
            "counts": counts, "bytes_gathered": wire}


# ------------------------------------------------------------------------------ the CUDA path
def sharded_sample(graph, query, num_walks=100, num_steps=3, bucket=-1, seed=111413, rng_mode=None, group=None,
                   bounds: Optional[np.ndarray] = None):
    """Sample this rank's seed range on its GPU and exchange shards: returns the full SpG (replicated
    on every rank), identical to SpG.sample(graph, query, ...) of a single process (bit for bit in
    RAND_R / TRACE-free modes: seed indices, rand_r offsets and LP ids are global)."""
    from .spg import SpG, _ptr, _stream, _view
    lib = _capi.load()
    world, rank, _ = _group_info(group)
    q = np.ascontiguousarray(np.asarray(query).astype(np.int32, copy=False))
    n = q.size
    lo, hi = partition(n, world, rank) if bounds is None else (int(bounds[rank]), int(bounds[rank + 1]))
    h = C.c_void_p()
    mode = _capi.SUBG_RNG_PHILOX if rng_mode is None else rng_mode
    _capi.check(lib.subg_gset_sample_shard(graph._h, _ptr(q), n, lo, hi, int(num_walks), int(num_steps), int(bucket),
                                           int(seed) & 0xFFFFFFFFFFFFFFFF, int(mode), None, _capi.SAMPLE_NO_RANKS,
                                           _stream(graph.device), C.byref(h)))
    shard = SpG(h, graph.device, n_nodes=graph.N, num_walks=num_walks)
    v = shard.views()
    dev = v["indices"].device
    st = _stream(graph.device)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    prof = os.environ.get("SUBG_PROFILE_HOST") is not None   # per-phase host times (adds a device sync per phase)
    marks, t_last = [], time.perf_counter()

    def mark(name):
        nonlocal t_last
        if prof:
            torch.cuda.synchronize()
            t = time.perf_counter()
            marks.append(f"{name}={1e3 * (t - t_last):.2f}")
            t_last = t
    ncol = num_steps + 1
    enc_local = v["enc"] if "enc" in v else torch.zeros((0, ncol), dtype=torch.int16, device=dev)
    nsize_local = v["nsize"] if "nsize" in v else torch.zeros(0, dtype=torch.int32, device=dev)
    # (1) ONE small equal-size all-gather carries every rank's sizes AND its unique LP table (a few hundred to a few
    # thousand int16 rows; _LP_INLINE_ROWS of them ride along, a larger table falls back to a second exchange)
    row_b = 2 * ncol
    blob = torch.zeros(32 + _LP_INLINE_ROWS * row_b, dtype=torch.uint8, device=dev)
    blob[:32] = torch.tensor([shard.n, shard.T, shard.c, shard.status], dtype=torch.int64).view(torch.uint8).to(dev)
    c_in = min(shard.c, _LP_INLINE_ROWS)
    if c_in:
        blob[32:32 + c_in * row_b] = enc_local[:c_in].contiguous().view(torch.uint8).reshape(-1)
    blobs = torch.empty((world, blob.numel()), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(blobs, blob, group=group)
    hb = blobs.cpu().numpy()
    counts = hb[:, :32].copy().view(np.int64).reshape(world, 4)
    n_r, T_r, c_r = counts[:, 0], counts[:, 1], counts[:, 2]
    n_tot, T_tot = int(n_r.sum()), int(T_r.sum())
    mark("sizes + LP tables")
    if int(c_r.max()) <= _LP_INLINE_ROWS:
        tables = [hb[r, 32:32 + int(c_r[r]) * row_b].copy().view(np.int16).reshape(int(c_r[r]), ncol) for r in range(world)]
    else:
        enc_all = torch.empty((int(c_r.sum()), ncol), dtype=torch.int16, device=dev)
        all_gather_varlen(enc_local.contiguous(), c_r, enc_all, group)
        enc_np = enc_all.cpu().numpy()
        offs = np.concatenate([[0], np.cumsum(c_r)])
        tables = [enc_np[offs[r]:offs[r + 1]] for r in range(world)]
    merged, maps = merge_lp_tables(tables)
    merged = np.ascontiguousarray(merged, dtype=np.int16)
    mark("lp merge")
    # (2) this rank's ids become global ids, the full SpG is allocated once and every shard lands in place
    _capi.check(lib.subg_spg_set_lp_table(shard._h, _ptr(maps[rank]), _ptr(merged), merged.shape[0], -1, st))
    mark("relabel")
    fh = C.c_void_p()
    _capi.check(lib.subg_spg_alloc(n_tot, T_tot, graph.device, st, C.byref(fh)))
    try:
        p = [C.c_void_p() for _ in range(6)]
        _capi.check(lib.subg_spg_views(fh, st, *[C.byref(x) for x in p]))
        g_indices = _view(p[1].value, (T_tot,), "<i4", graph.device, None)
        g_data = _view(p[2].value, (T_tot,), "<i4", graph.device, None)
        g_nsize = _view(p[5].value, (n_tot,), "<i4", graph.device, None)
        # (3) one NCCL group: set sizes, node ids and LP ids of every peer (grouped send / recv over NVLink)
        ops = []
        for local, cnt, out in ((nsize_local, n_r, g_nsize), (v["indices"], T_r, g_indices), (v["data"], T_r, g_data)):
            ops += _p2p_ops(local, cnt, out, group)
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        mark("shards")
        _capi.check(lib.subg_spg_seal(fh, st))
        _capi.check(lib.subg_spg_set_lp_table(fh, None, _ptr(merged), merged.shape[0], ncol, st))
        mark("seal")
        if prof and rank == 0:
            print("[subg host ms] exchange: " + " ".join(marks), file=sys.stderr, flush=True)
    except Exception:
        lib.subg_spg_free(fh)
        raise
    status = int(np.bitwise_or.reduce(counts[:, 3]))
    shard.close()
    full = SpG(fh, graph.device, n_nodes=graph.N, num_walks=num_walks)
    full.status = status
    ev[1].record()
    full.exchange_bytes = 8 * T_tot + 4 * n_tot + 2 * ncol * int(c_r.sum())
    full.exchange_events = ev   # elapsed = LP-table merge + all-gathers + CSR assembly (device time)
    return full


def sharded_subg_matrix(G, train_idx, num_walks=200, num_steps=4, device=None, seed=111413, rng_mode=None, group=None):
    """Multi-GPU subg_matrix (sampler/random_walks.py:74-82): (z, enc) with z the replicated device SpG."""
    from .spg import DeviceGraph
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    graph = DeviceGraph.from_scipy(G, device)
    z = sharded_sample(graph, np.asarray(train_idx), num_walks=num_walks, num_steps=num_steps - 1, seed=seed,
                       rng_mode=rng_mode, group=group)
    graph.close()
    return z, z.enc_table()
