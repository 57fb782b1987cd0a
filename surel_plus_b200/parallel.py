"""Multi-GPU SubGAcc: one process per GPU (torch.distributed), graph replicated in every GPU's
HBM, seeds range-partitioned, SpG shards all-gathered once (NCCL over NVLink/NVSwitch), queries
then joined locally on the replicated SpG.  The reference has no
distributed code at all; this is the B200 scaling path of the same operators.

Exchange step (`sharded_sample`; kernels in csrc/xchg.cu):
  1. the shard is packed (4-8 bytes per set entry) into the rank's IPC-exported slab, together with its set sizes,
     row offsets and unique LP keys + first stream positions;
  2. one 64-byte header all-gather (sizes, wire format): the only host-visible collective, and the barrier after
     which every slab is complete;
  3. on every GPU: the unique LP tables are merged ON THE DEVICE in first-occurrence order (positions are global, so
     this is the order of the single-process scan, subg_acc.c:957-978), and ONE kernel pulls all peers' packed
     entries over NVLink (plain loads from the mapped slabs), widens and relabels them in flight and writes the full
     SpG in place.  Mode 'nccl' moves the slabs with an NCCL all-gather into a staging buffer instead.
`assemble_shards` is the same exchange on plain tensors with staged all-gathers (any backend; CPU tests).
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
def slab_bytes_needed(n_seeds: int, row_cap: int, lp_rows: int = 1 << 20) -> int:
    """Upper bound of a shard's packed size (csrc/xchg.cu slab_layout): 8 bytes per entry in the widest wire
    format, rows padded to 4 entries, 12 bytes per seed, 16 per unique LP row, 128-byte aligned sections."""
    row_cap = (int(row_cap) + 3) & ~3
    return 8 * n_seeds * row_cap + 12 * n_seeds + 16 * lp_rows + 8 * 128 + 64


def agree(flag: bool, device=None, group=None) -> bool:
    """True on every rank iff `flag` is true on every rank (a decision all ranks have to take together,
    e.g. whether the peers' slabs could be mapped)."""
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


class ShardExchange:
    """This rank's side of the multi-GPU exchange (include/subg_b200.h, subg_xchg_*): the IPC-exported slab the
    rank's shard is packed into, and the mappings of the peers' slabs.  mode 'peer': the assemble kernel pulls the
    peers' slabs over NVLink itself; mode 'nccl': the slabs are all-gathered into a staging buffer with NCCL and
    unpacked from there (fallback when cudaIpcOpenMemHandle is not possible, and the comparison point)."""

    def __init__(self, device: int, slab_bytes: int, group=None, mode: Optional[str] = None):
        self._lib = _capi.load()
        self.device, self.group = device, group
        self.world, self.rank, self.backend = _group_info(group)
        self._h = C.c_void_p()
        _capi.check(self._lib.subg_xchg_create(device, self.rank, self.world, int(slab_bytes), C.byref(self._h)))
        p, b = C.c_void_p(), C.c_int64()
        _capi.check(self._lib.subg_xchg_slab(self._h, C.byref(p), C.byref(b)))
        self.slab_ptr, self.slab_bytes = p.value, b.value
        tdev = torch.device("cuda", device)
        want = (mode or os.environ.get("SUBG_EXCHANGE", "peer")).lower()
        self.mode = "nccl"
        self.why = "requested"
        if want == "peer" and self.world > 1:
            # the 64-byte IPC handles travel through the process group; every rank maps every peer
            mine = np.zeros(64, np.uint8)
            _capi.check(self._lib.subg_xchg_export(self._h, mine.ctypes.data))
            allh = torch.empty((self.world, 64), dtype=torch.uint8, device=tdev)
            dist.all_gather_into_tensor(allh, torch.from_numpy(mine).to(tdev), group=group)
            handles = np.ascontiguousarray(allh.cpu().numpy())
            rc = self._lib.subg_xchg_open(self._h, handles.ctypes.data)
            err = self._lib.subg_last_error().decode("utf-8", "replace") if rc else ""
            if agree(rc == 0, tdev, group):
                self.mode = "peer"
            else:
                self.why = f"peer mapping unavailable ({err or 'on another rank'})"
                if self.rank == 0:
                    print(f"[surel_plus_b200] exchange falls back to NCCL staging: {self.why}", file=sys.stderr, flush=True)
        elif self.world == 1:
            self.mode = "peer"

    def slab_view(self, nbytes: int) -> torch.Tensor:
        from .spg import _view
        return _view(self.slab_ptr, (int(nbytes),), "|u1", self.device, self)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.subg_xchg_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_exchanges: dict = {}


def exchange_for(device: int, need_bytes: int, group=None, mode: Optional[str] = None) -> ShardExchange:
    """Cached ShardExchange of (device, group).  `need_bytes` must be the same on every rank (the slab is
    re-created collectively when it is too small)."""
    key = (device, id(group) if group is not None else None, mode)
    x = _exchanges.get(key)
    if x is not None and x.slab_bytes >= need_bytes:
        return x
    if x is not None:
        torch.cuda.synchronize(device)
        dist.barrier(group=group)  # nobody is still reading the old slab
        x.close()
    x = ShardExchange(device, int(need_bytes * 1.05) + (1 << 20), group, mode)
    _exchanges[key] = x
    return x


def close_exchanges():
    for x in list(_exchanges.values()):
        x.close()
    _exchanges.clear()


H_N, H_T, H_EXTENT, H_C, H_FMT, H_MAXSET, H_STATUS, H_BYTES = range(8)


def sharded_sample(graph, query, num_walks=100, num_steps=3, bucket=-1, seed=111413, rng_mode=None, group=None,
                   bounds: Optional[np.ndarray] = None, mode: Optional[str] = None):
    """Sample this rank's seed range on its GPU and exchange shards: returns the full SpG (replicated
    on every rank), identical to SpG.sample(graph, query, ...) of a single process (bit for bit: seed
    indices, Philox counters, rand_r offsets and LP first-occurrence positions are global).

    Pass = sample (one kernel) -> pack into the rank's slab -> header all-gather (barrier) -> device-side LP
    table merge + ONE pull kernel over the peers' slabs (NVLink) -> end barrier.  See csrc/xchg.cu."""
    from .spg import SpG, _ptr, _stream
    lib = _capi.load()
    world, rank, _ = _group_info(group)
    q = query if isinstance(query, torch.Tensor) else np.ascontiguousarray(np.asarray(query).astype(np.int32, copy=False))
    n = q.numel() if isinstance(q, torch.Tensor) else q.size
    if bounds is None:
        bounds = np.array([partition(n, world, r)[0] for r in range(world)] + [n], dtype=np.int64)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    Kt = num_walks * num_steps + 1
    row_cap = min(Kt, bucket) if bucket > 0 else Kt
    need = slab_bytes_needed(int(np.max(np.diff(bounds))), row_cap)
    xc = exchange_for(graph.device, need, group, mode)
    st = _stream(graph.device)
    tdev = torch.device("cuda", graph.device)
    h = C.c_void_p()
    rmode = _capi.SUBG_RNG_PHILOX if rng_mode is None else rng_mode
    _capi.check(lib.subg_gset_sample_shard(graph._h, _ptr(q), n, lo, hi, int(num_walks), int(num_steps), int(bucket),
                                           int(seed) & 0xFFFFFFFFFFFFFFFF, int(rmode), None,
                                           _capi.SAMPLE_NO_RANKS | _capi.SAMPLE_NO_COMPACT, st, C.byref(h)))
    shard = SpG(h, graph.device, n_nodes=graph.N, num_walks=num_walks)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    prof = os.environ.get("SUBG_PROFILE_HOST") is not None   # per-phase wall times (adds a device sync per phase)
    marks, t_last = [], time.perf_counter()

    def mark(name):
        nonlocal t_last
        if prof:
            torch.cuda.synchronize()
            t = time.perf_counter()
            marks.append(f"{name}={1e3 * (t - t_last):.2f}")
            t_last = t
    mark("sync-after-sample")
    hdr = np.zeros(8, np.int64)
    _capi.check(lib.subg_xchg_pack(xc._h, shard._h, graph.N, hdr.ctypes.data, st))
    mark("pack")
    # the header all-gather is the barrier: when it completes here, every peer's pack kernel has completed
    allh = torch.empty((world, 8), dtype=torch.int64, device=tdev)
    dist.all_gather_into_tensor(allh, torch.from_numpy(hdr).to(tdev), group=group)
    headers = np.ascontiguousarray(allh.cpu().numpy())
    mark("headers")
    if (headers[:, H_FMT] < 0).any():
        shard.close()
        raise MemoryError("a shard did not fit its exchange slab")
    srcs = None
    staging = None
    if xc.mode == "nccl" and world > 1:
        stride = (int(headers[:, H_BYTES].max()) + 4095) & ~4095
        staging = torch.empty((world, stride), dtype=torch.uint8, device=tdev)
        dist.all_gather_into_tensor(staging, xc.slab_view(stride), group=group)
        srcs = (C.c_void_p * world)(*[staging.data_ptr() + r * stride for r in range(world)])
    fh = C.c_void_p()
    mark("staging" if staging is not None else "-")
    _capi.check(lib.subg_xchg_assemble(xc._h, headers.ctypes.data, srcs, int(num_walks), int(num_steps) + 1, st, C.byref(fh)))
    mark("assemble")
    if xc.mode == "peer" and world > 1:
        # nobody may re-pack its slab before every peer has pulled it
        dist.all_reduce(torch.zeros(1, dtype=torch.int32, device=tdev), group=group)
    mark("end-barrier")
    shard.close()
    full = SpG(fh, graph.device, n_nodes=graph.N, num_walks=num_walks)
    mark("close-shard")
    if prof and rank == 0:
        print("[subg host ms] sharded_sample: " + " ".join(marks), file=sys.stderr, flush=True)
    ev[1].record()
    received = int(headers[:, H_BYTES].sum() - headers[rank, H_BYTES])
    full.exchange_mode = xc.mode
    full.exchange_bytes = int(headers[:, H_BYTES].sum())         # packed bytes of all shards
    full.exchange_received = received                            # what this GPU read from its peers
    full.exchange_entry_bytes = 4 + int(headers[rank, H_FMT] & 0xff)
    full.exchange_events = ev   # elapsed = pack + header all-gather + LP merge + pull (+ end barrier), device time
    return full


def linked_exchange(graph, n_seeds_max: int, num_walks: int, num_steps: int, bucket: int = -1, group=None) -> ShardExchange:
    """A slab (with the peers' mappings) large enough for linked_sample of seed ranges of up to n_seeds_max seeds; pass it
    as `exchange=` to reuse it over several passes (each pass re-stages it: close the previous linked SpG on every rank
    first)."""
    Kt = num_walks * num_steps + 1
    row_cap = (min(Kt, bucket) if bucket > 0 else Kt) + 3 & ~3
    need = slab_bytes_needed(int(n_seeds_max), row_cap)
    xc = ShardExchange(graph.device, int(need * 1.05) + (1 << 20), group, "peer")
    if xc.mode != "peer":
        why = xc.why
        xc.close()
        raise RuntimeError(f"linked shards need the peers' slabs mapped: {why}")
    return xc


def linked_sample(graph, query, num_walks=100, num_steps=3, bucket=-1, seed=111413, rng_mode=None, group=None,
                  bounds: Optional[np.ndarray] = None, exchange: Optional[ShardExchange] = None):
    """The no-replication alternative to sharded_sample: every rank samples its seed range and keeps its rows; the shards
    are LINKED (csrc/xchg.cu: staged planes in the IPC-mapped slabs, LP tables merged, ids relabelled in place, 12 bytes of
    row metadata per seed fetched).  Returns an SpG with the same rows, ids and LP table as SpG.sample of one process,
    whose remote rows are read over NVLink when a query joins them: the pass costs no bulk transfer, the joins do.
    The SpG keeps its exchange context alive (and with it the slab the peers read): close it on every rank together."""
    from .spg import SpG, _ptr, _stream
    lib = _capi.load()
    world, rank, _ = _group_info(group)
    q = query if isinstance(query, torch.Tensor) else np.ascontiguousarray(np.asarray(query).astype(np.int32, copy=False))
    n = q.numel() if isinstance(q, torch.Tensor) else q.size
    if bounds is None:
        bounds = np.array([partition(n, world, r)[0] for r in range(world)] + [n], dtype=np.int64)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    Kt = num_walks * num_steps + 1
    row_cap = (min(Kt, bucket) if bucket > 0 else Kt) + 3 & ~3
    plane = int(np.max(np.diff(bounds))) * row_cap
    # the slab is owned by the result (or by the caller who passed `exchange`), never the cache sharded_sample re-packs
    xc = exchange if exchange is not None else linked_exchange(graph, int(np.max(np.diff(bounds))), num_walks, num_steps, bucket, group)
    st = _stream(graph.device)
    tdev = torch.device("cuda", graph.device)
    h = C.c_void_p()
    rmode = _capi.SUBG_RNG_PHILOX if rng_mode is None else rng_mode
    _capi.check(lib.subg_gset_sample_shard(graph._h, _ptr(q), n, lo, hi, int(num_walks), int(num_steps), int(bucket),
                                           int(seed) & 0xFFFFFFFFFFFFFFFF, int(rmode), None,
                                           _capi.SAMPLE_NO_RANKS | _capi.SAMPLE_NO_COMPACT, st, C.byref(h)))
    shard = SpG(h, graph.device, n_nodes=graph.N, num_walks=num_walks)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    hdr = np.zeros(8, np.int64)
    _capi.check(lib.subg_xchg_stage(xc._h, shard._h, graph.N, plane, hdr.ctypes.data, st))
    allh = torch.empty((world, 8), dtype=torch.int64, device=tdev)
    dist.all_gather_into_tensor(allh, torch.from_numpy(hdr).to(tdev), group=group)    # barrier: every shard is staged
    headers = np.ascontiguousarray(allh.cpu().numpy())
    if (headers[:, H_FMT] < 0).any():
        shard.close()
        if exchange is None:
            xc.close()
        raise MemoryError("a shard did not fit its exchange slab")
    fh = C.c_void_p()
    _capi.check(lib.subg_xchg_link(xc._h, headers.ctypes.data, None, int(num_walks), int(num_steps) + 1, st, C.byref(fh)))
    dist.all_reduce(torch.zeros(1, dtype=torch.int32, device=tdev), group=group)      # barrier: every id plane is relabelled
    shard.close()
    ev[1].record()
    full = SpG(fh, graph.device, n_nodes=graph.N, num_walks=num_walks)
    full._exchange = xc            # the slabs live as long as the SpG
    full.exchange_mode = "linked"
    full.exchange_events = ev
    return full


def sharded_subg_matrix(G, train_idx, num_walks=200, num_steps=4, device=None, seed=111413, rng_mode=None, group=None):
    """Multi-GPU subg_matrix (sampler/random_walks.py:74-82): (z, enc) with z the replicated device SpG."""
    from .spg import DeviceGraph
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    world, _, backend = _group_info(group)
    graph = DeviceGraph.from_scipy_sharded(G, device, group) if (world > 1 and backend == "nccl") else DeviceGraph.from_scipy(G, device)
    deg = np.diff(G.indptr)
    idx = np.asarray(train_idx)
    w = 0.5 * num_walks * (num_steps - 1) + 1.5 * np.minimum(deg[idx], num_walks)   # set-size estimate per seed
    z = sharded_sample(graph, idx, num_walks=num_walks, num_steps=num_steps - 1, seed=seed,
                       rng_mode=rng_mode, group=group, bounds=partition_by_work(w, world))
    graph.close()
    return z, z.enc_table()
