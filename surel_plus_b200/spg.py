"""Device-resident graph and SpG (CSR-of-sets) handles over the C ABI.

`SpG` stands where the reference keeps a scipy CSR (`z` of
sampler/random_walks.py:79): row u = sorted node set S_u, data = LP-row id + 1
(or a float64 structural value in PPR/SPD mode).  It stays in HBM; SpJoin reads it
in place (surel_plus_b200/train.py).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np
import torch

from . import _capi


def _dev_index(device) -> int:
    d = torch.device(device) if not isinstance(device, torch.device) else device
    if d.type != "cuda":
        raise RuntimeError("surel_plus_b200 runs on CUDA devices only (no CPU fallback)")
    return d.index if d.index is not None else torch.cuda.current_device()


def _stream(dev: int) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(a) -> int:
    """Raw address of a numpy array / torch tensor (host or device)."""
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    return a.ctypes.data


class _DevView:
    """Zero-copy view of library-owned device memory for torch.as_tensor()."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}
        self._owner = owner


def _view(ptr, shape, typestr, dev, owner):
    if ptr is None or any(s == 0 for s in shape):
        dt = {"<i8": torch.int64, "<i4": torch.int32, "<i2": torch.int16, "<u2": torch.uint16, "<f8": torch.float64,
              "|u1": torch.uint8, "<f4": torch.float32}[typestr]
        return torch.empty(shape, dtype=dt, device=f"cuda:{dev}")
    return torch.as_tensor(_DevView(ptr, shape, typestr, owner), device=f"cuda:{dev}")


class _PinnedPool:
    """Page-locked host buffers for the arrays handed back to the caller (subg_host_alloc).  Locking
    pages costs about a second per GB, so blocks are recycled: a block returns to the pool when the
    numpy array built over it (and every view of it) is garbage collected.  Blocks are 12.5 % larger than
    asked so that the next call, whose sizes differ slightly, still fits."""

    def __init__(self, keep_bytes: int | None = None):
        if keep_bytes is None:  # default 6 GiB per node, shared by the ranks of a torchrun job (one process per GPU)
            local_world = max(int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1), 1)
            keep_bytes = int(os.environ.get("SUBG_PINNED_POOL_BYTES", (6 << 30) // local_world))
        self._free: list[tuple[int, int]] = []  # (capacity, address)
        self._keep = keep_bytes
        self._lib = None
        self._lock = threading.RLock()  # re-entrant: blocks come back from __del__, i.e. from whichever thread drops the last view

    def take(self, nbytes: int) -> tuple[int, int]:
        with self._lock:
            fit = [b for b in self._free if b[0] >= nbytes]
            if fit:
                b = min(fit)
                self._free.remove(b)
                return b
        with self._lock:
            self._lib = self._lib or _capi.load()
        cap = max(4096, ((nbytes + (nbytes >> 3) + (1 << 21) - 1) >> 21) << 21) if nbytes > (1 << 20) else max(nbytes, 64)
        p = C.c_void_p()
        _capi.check(self._lib.subg_host_alloc(C.byref(p), cap))
        return cap, p.value

    def trim(self, keep_bytes: int = 0) -> None:
        """Return pooled blocks to the driver until at most keep_bytes stay page-locked."""
        with self._lock:
            while self._free and sum(b[0] for b in self._free) > keep_bytes:
                c, a = self._free.pop(0)
                self._lib.subg_host_free(C.c_void_p(a))

    def give(self, cap: int, addr: int) -> None:
        try:
            with self._lock:
                self._free.append((cap, addr))
                while sum(b[0] for b in self._free) > self._keep:
                    c, a = self._free.pop(0)
                    self._lib.subg_host_free(C.c_void_p(a))
        except Exception:  # interpreter shutdown
            pass


_pinned = _PinnedPool()


class _PinnedBlock:
    """Owner of one pooled block; numpy arrays reference it through __array_interface__."""

    def __init__(self, shape, dtype):
        self.cap = self.addr = None   # __del__ runs even if take() raises
        self.shape, self.dtype = tuple(int(x) for x in shape), np.dtype(dtype)
        nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self.cap, self.addr = _pinned.take(nbytes)
        self.__array_interface__ = {"shape": self.shape, "typestr": self.dtype.str, "data": (self.addr, False), "version": 3}

    def __del__(self):
        if self.addr is not None:
            _pinned.give(self.cap, self.addr)


def pinned_empty(shape, dtype) -> np.ndarray:
    """Uninitialised numpy array over pooled page-locked memory (asynchronous D2H target)."""
    return np.asarray(_PinnedBlock(shape, dtype))


_copy_pool = None


def staged_h2d(arr: np.ndarray, device) -> torch.Tensor:
    """Pageable host array -> device tensor through pooled page-locked staging: worker threads copy 8 MB chunks into the
    pinned block (numpy releases the GIL for plain copies) while the chunks that are ready are already on their way
    over PCIe.  A plain cudaMemcpy from pageable memory runs at roughly a quarter of the link rate (one driver thread
    bouncing through its own staging)."""
    global _copy_pool
    from concurrent.futures import ThreadPoolExecutor
    a = np.ascontiguousarray(arr)
    dev = torch.device("cuda", _dev_index(device))
    out = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, device=dev)
    flat = a.reshape(-1).view(np.uint8)
    if flat.size < (32 << 20):
        out.copy_(torch.from_numpy(a))
        return out
    if _copy_pool is None:
        nthr = int(os.environ.get("SUBG_COPY_THREADS", "0")) or min(6, max(2, (os.cpu_count() or 4) // 2))
        _copy_pool = ThreadPoolExecutor(max_workers=nthr)
    stage = pinned_empty((flat.size,), np.uint8)
    stage_t = torch.from_numpy(stage)
    dst = out.reshape(-1).view(torch.uint8)
    CH = int(os.environ.get("SUBG_COPY_CHUNK_MB", "8")) << 20
    futs = [(lo, min(lo + CH, flat.size), _copy_pool.submit(np.copyto, stage[lo:min(lo + CH, flat.size)], flat[lo:min(lo + CH, flat.size)]))
            for lo in range(0, flat.size, CH)]
    for lo, hi, f in futs:
        f.result()
        dst[lo:hi].copy_(stage_t[lo:hi], non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()   # the staging block goes back to the pool when `stage` dies
    return out


class DeviceGraph:
    """CSR adjacency resident in HBM (replaces the `indptr, indices` numpy arguments of
    gset_sampler, subg_acc/subg_acc.c:655-671)."""

    def __init__(self, indptr, indices, device="cuda"):
        self._lib = _capi.load()
        self.device = _dev_index(device)
        self._h = C.c_void_p()
        if isinstance(indices, np.ndarray) and indices.dtype == np.int32 and indices.nbytes >= (32 << 20) \
                and isinstance(indptr, np.ndarray) and indptr.dtype in (np.int32, np.int64):
            # large pageable arrays: pipelined upload through pinned staging (staged_h2d)
            indptr, indices = staged_h2d(indptr, self.device), staged_h2d(indices, self.device)
        if isinstance(indptr, torch.Tensor):
            if indptr.dtype not in (torch.int32, torch.int64) or indices.dtype != torch.int32:
                raise TypeError("indptr must be int32/int64 and indices int32")
            indptr, indices = indptr.contiguous(), indices.contiguous()
            is64 = indptr.dtype == torch.int64
            n1, E = indptr.numel(), indices.numel()
        else:
            indptr = np.asarray(indptr)
            if indptr.dtype not in (np.int32, np.int64):
                raise TypeError(f"Cannot cast indptr from {indptr.dtype} to int32/int64")
            indptr = np.ascontiguousarray(indptr)
            indices = np.asarray(indices)
            if not np.can_cast(indices.dtype, np.int32, "safe") and indices.size and indices.max() >= 2 ** 31:
                raise TypeError("indices must fit int32")
            indices = np.ascontiguousarray(indices, dtype=np.int32)
            is64 = indptr.dtype == np.int64
            n1, E = indptr.size, indices.size
        self.N, self.E = n1 - 1, E
        self._keep = (indptr, indices)
        _capi.check(self._lib.subg_graph_create(_ptr(indptr), int(is64), _ptr(indices), self.N, self.E, self.device,
                                                _stream(self.device), C.byref(self._h)))
        self._keep = None

    @classmethod
    def from_scipy(cls, G, device="cuda"):
        return cls(G.indptr, G.indices, device)

    @classmethod
    def from_scipy_sharded(cls, G, device, group=None):
        """Same graph on every rank of `group`, uploaded cooperatively: every rank copies only its 1/world slice of the
        column array from host memory and the slices are all-gathered over NVLink (NCCL), instead of `world` processes
        pushing the whole CSR through the host's memory system at once.  Every rank must hold the same G."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = torch.device("cuda", _dev_index(device))
        indices = np.ascontiguousarray(G.indices, dtype=np.int32)
        E = indices.size
        chunk = max((E + world - 1) // world, 1)
        lo, hi = min(rank * chunk, E), min((rank + 1) * chunk, E)
        mine = torch.zeros(chunk, dtype=torch.int32, device=dev)
        if hi > lo:
            mine[:hi - lo].copy_(torch.from_numpy(indices[lo:hi]))
        full = torch.empty(world * chunk, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(full, mine, group=group)
        indptr = torch.from_numpy(np.ascontiguousarray(G.indptr)).to(dev)
        if indptr.dtype not in (torch.int32, torch.int64):
            indptr = indptr.to(torch.int64)
        return cls(indptr, full[:E], dev)

    @classmethod
    def from_edges(cls, row, col, num_nodes=None, symmetrize=False, drop_self_loops=False, device="cuda"):
        """Edge list -> CSR on the device: duplicates coalesced, columns ascending, as scipy's
        csr_matrix((ones, (row, col))) in edge2csr (subg_acc/test/test.py:15-19); symmetrize=True also
        stores every edge reversed (dataloader.py:119-129).  row / col: int64 numpy arrays or torch tensors
        (host or device)."""
        self = cls.__new__(cls)
        self._lib = _capi.load()
        if isinstance(row, torch.Tensor):
            row, col = row.to(torch.int64).contiguous(), col.to(torch.int64).contiguous()
            E = row.numel()
            if row.is_cuda:
                device = row.device
        else:
            row = np.ascontiguousarray(np.asarray(row), dtype=np.int64)
            col = np.ascontiguousarray(np.asarray(col), dtype=np.int64)
            E = row.size
        if (col.numel() if isinstance(col, torch.Tensor) else col.size) != E:
            raise TypeError("row and col must have the same length")
        self.device = _dev_index(device)
        self._h = C.c_void_p()
        self._keep = None
        _capi.check(self._lib.subg_graph_from_edges(_ptr(row), _ptr(col), E, -1 if num_nodes is None else int(num_nodes),
                                                    int(bool(symmetrize)), int(bool(drop_self_loops)), self.device,
                                                    _stream(self.device), C.byref(self._h)))
        N, Ecount = C.c_int64(), C.c_int64()
        _capi.check(self._lib.subg_graph_info(self._h, C.byref(N), C.byref(Ecount), None))
        self.N, self.E = N.value, Ecount.value
        return self

    def csr(self, device=None):
        """(indptr int64[N+1], indices int32[E]) of the resident graph: numpy arrays, or torch tensors
        on `device` when given."""
        if device is None:
            indptr, indices = np.empty(self.N + 1, np.int64), np.empty(self.E, np.int32)
        else:
            indptr = torch.empty(self.N + 1, dtype=torch.int64, device=device)
            indices = torch.empty(self.E, dtype=torch.int32, device=device)
        _capi.check(self._lib.subg_graph_export(self._h, _ptr(indptr), _ptr(indices), _stream(self.device)))
        return indptr, indices

    def to_scipy(self):
        """Boolean scipy CSR of the resident graph (what edge2csr returns, test.py:17-19)."""
        import scipy.sparse as sp
        indptr, indices = self.csr()
        return sp.csr_matrix((np.ones(self.E, dtype=bool), indices, indptr), shape=(self.N, self.N))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.subg_graph_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SpG:
    """Handle of a device SpG.  Attributes: n rows, T entries, c unique LP rows, ncol,
    max_set, status, value_kind (0 int pointers, 1 float64 values), shape (scipy-like)."""

    def __init__(self, handle: C.c_void_p, device: int, n_nodes: int | None = None, num_walks: int | None = None):
        self._lib = _capi.load()
        self._h = handle
        self.device = device
        n, T = C.c_int64(), C.c_int64()
        c, ncol, mx, vk = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        st = C.c_uint32()
        _capi.check(self._lib.subg_spg_info(handle, C.byref(n), C.byref(T), C.byref(c), C.byref(ncol), C.byref(mx),
                                            C.byref(st), C.byref(vk)))
        self.n, self.T, self.c, self.ncol = n.value, T.value, c.value, ncol.value
        self.max_set, self.status, self.value_kind = mx.value, st.value, vk.value
        self.num_walks = num_walks
        self.shape = (self.n, n_nodes if n_nodes is not None else self.n)
        self.nnz = self.T

    # ---- construction -------------------------------------------------------
    @classmethod
    def sample(cls, graph: DeviceGraph, query, num_walks=100, num_steps=3, bucket=-1, seed=111413,
               rng_mode=_capi.SUBG_RNG_PHILOX, walks=None, first_visit_ranks=True, dump_walks=False) -> "SpG":
        """Walk-based set sampling + LP encoding + SpG build on the device
        (subg_acc.c:649-1034 and random_walks.py:79).  `num_steps` is the walk length m.
        first_visit_ranks=False skips the per-entry first-visit rank that only export_reference()
        (the reference's `remap` order) needs; the SpG itself is identical.  dump_walks=True keeps the
        walks the kernel drew (walks()): the parity hook of the Philox path."""
        lib = _capi.load()
        if isinstance(query, torch.Tensor):
            q = query.to(torch.int32).contiguous()
            n = q.numel()
        else:
            q = np.ascontiguousarray(np.asarray(query).astype(np.int32, copy=False))
            n = q.size
        w = None
        if rng_mode == _capi.SUBG_RNG_TRACE:
            if walks is None:
                raise TypeError("trace mode needs walks[n, num_walks, num_steps]")
            w = walks.to(torch.int32).contiguous() if isinstance(walks, torch.Tensor) else np.ascontiguousarray(walks, dtype=np.int32)
            if (w.numel() if isinstance(w, torch.Tensor) else w.size) != n * num_walks * num_steps:
                raise TypeError("walks must have n*num_walks*num_steps entries")
        h = C.c_void_p()
        _capi.check(lib.subg_gset_sample(graph._h, _ptr(q), n, int(num_walks), int(num_steps), int(bucket),
                                         int(seed) & 0xFFFFFFFFFFFFFFFF, int(rng_mode), _ptr(w) if w is not None else None,
                                         (0 if first_visit_ranks else _capi.SAMPLE_NO_RANKS)
                                         | (_capi.SAMPLE_DUMP_WALKS if dump_walks else 0), _stream(graph.device),
                                         C.byref(h)))
        out = cls(h, graph.device, n_nodes=graph.N, num_walks=num_walks)
        out.num_steps = int(num_steps)
        return out

    @classmethod
    def from_scipy(cls, x, device="cuda") -> "SpG":
        """Upload a scipy CSR produced by the reference's subg_matrix / topk_ppr_matrix+encoding."""
        lib = _capi.load()
        dev = _dev_index(device)
        x = x.tocsr()
        if not x.has_sorted_indices:
            x = x.sorted_indices()
        indptr = np.ascontiguousarray(x.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(x.indices, dtype=np.int32)
        if np.issubdtype(x.data.dtype, np.floating):
            data, kind = np.ascontiguousarray(x.data, dtype=np.float64), 1
        else:
            data, kind = np.ascontiguousarray(x.data, dtype=np.int32), 0
        h = C.c_void_p()
        _capi.check(lib.subg_spg_from_csr(_ptr(indptr), _ptr(indices), _ptr(data), kind, x.shape[0], indices.size, dev,
                                          _stream(dev), C.byref(h)))
        return cls(h, dev, n_nodes=x.shape[1])

    @classmethod
    def from_device_csr(cls, indptr, indices, data, n_nodes=None, enc=None, num_walks=None, status=0) -> "SpG":
        """Wrap device tensors (indptr int64 [n+1], indices int32 [T] ascending per row, data int32 [T] =
        LP id + 1 or float64 [T]) as an SpG; the arrays are copied into library-owned HBM.  `enc`: int16
        [c, ncol] LP table (numpy or tensor) attached to the handle."""
        lib = _capi.load()
        dev = indices.device.index if indices.device.index is not None else torch.cuda.current_device()
        indptr = indptr.to(torch.int64).contiguous()
        indices = indices.to(torch.int32).contiguous()
        kind = 1 if data.dtype.is_floating_point else 0
        data = data.to(torch.float64 if kind else torch.int32).contiguous()
        h = C.c_void_p()
        _capi.check(lib.subg_spg_from_csr(indptr.data_ptr(), indices.data_ptr(), data.data_ptr(), kind,
                                          indptr.numel() - 1, indices.numel(), dev, _stream(dev), C.byref(h)))
        if enc is not None and not kind:
            e = enc.cpu().numpy() if isinstance(enc, torch.Tensor) else np.asarray(enc)
            e = np.ascontiguousarray(e, dtype=np.int16)
            _capi.check(lib.subg_spg_set_lp_table(h, None, _ptr(e), e.shape[0], e.shape[1], _stream(dev)))
        out = cls(h, dev, n_nodes=n_nodes, num_walks=num_walks)
        out.status = status
        return out

    # ---- views / export -----------------------------------------------------
    def views(self) -> dict:
        """Zero-copy torch views of the device arrays in the CSR layout (valid while this object
        lives).  A sampler-built SpG is compacted into that layout on the first call."""
        p = [C.c_void_p() for _ in range(6)]
        _capi.check(self._lib.subg_spg_views(self._h, _stream(self.device), *[C.byref(x) for x in p]))
        d = self.device
        out = {
            "indptr": _view(p[0].value, (self.n + 1,), "<i8", d, self),
            "indices": _view(p[1].value, (self.T,), "<i4", d, self),
            "data": _view(p[2].value, (self.T,), "<f8" if self.value_kind else "<i4", d, self),
        }
        if not self.value_kind:
            if p[3].value:
                out["slot"] = _view(p[3].value, (self.T,), "<u2", d, self)
            if p[4].value and self.ncol > 0:
                out["enc"] = _view(p[4].value, (self.c, self.ncol), "<i2", d, self)
            if p[5].value:
                out["nsize"] = _view(p[5].value, (self.n,), "<i4", d, self)
        return out

    def expand_rows(self, num_nodes: int) -> "SpG":
        """Rows in seed order -> one (possibly empty) row per graph node (subg_spg_expand_rows): what
        subg_matrix's (N, N) csr_matrix is for a query that is not arange(N)."""
        _capi.check(self._lib.subg_spg_expand_rows(self._h, int(num_nodes), _stream(self.device)))
        self.n = int(num_nodes)
        self.shape = (self.n, self.shape[1])
        return self

    def walks(self) -> torch.Tensor:
        """int32 [n, num_walks, num_steps] view of the walks kept by sample(..., dump_walks=True)."""
        p = C.c_void_p()
        _capi.check(self._lib.subg_spg_walks(self._h, _stream(self.device), C.byref(p)))
        if not p.value:
            raise RuntimeError("this SpG was sampled without dump_walks=True")
        return _view(p.value, (self.n, self.num_walks, self.ncol - 1), "<i4", self.device, self)

    def set_sizes(self) -> torch.Tensor:
        """int32 [n] set sizes without forcing the CSR layout."""
        p = [C.c_void_p() for _ in range(4)]
        ext = C.c_int64()
        _capi.check(self._lib.subg_spg_rows(self._h, *[C.byref(x) for x in p], C.byref(ext)))
        if p[1].value:
            return _view(p[1].value, (self.n,), "<i4", self.device, self).clone()
        rb = _view(p[0].value, (self.n + 1,), "<i8", self.device, self)
        return (rb[1:] - rb[:-1]).to(torch.int32)

    def rows(self) -> dict:
        """Zero-copy torch views of the layout SpJoin reads, without compaction: row u = entries
        [rowbeg[u], rowbeg[u] + nsize[u]) of indices / data (sampler-built SpGs are "scattered": rows are
        16-byte aligned and in completion order).  Invalidated by views() / to_scipy(), which compact."""
        p = [C.c_void_p() for _ in range(4)]
        ext = C.c_int64()
        _capi.check(self._lib.subg_spg_rows(self._h, *[C.byref(x) for x in p], C.byref(ext)))
        d = self.device
        compact = not p[1].value
        rb = _view(p[0].value, (self.n + 1 if compact else self.n,), "<i8", d, self)
        nsize = (rb[1:] - rb[:-1]).to(torch.int32) if compact else _view(p[1].value, (self.n,), "<i4", d, self)
        return {"rowbeg": rb[:self.n], "nsize": nsize,
                "indices": _view(p[2].value, (ext.value,), "<i4", d, self),
                "data": _view(p[3].value, (ext.value,), "<f8" if self.value_kind else "<i4", d, self)}

    def export_reference(self, want_raw: bool = False):
        """[nsize, remap, enc(, raw_enc)] exactly as gset_sampler returns them
        (subg_acc.c:1017-1024).  Host numpy arrays (filled through pinned memory)."""
        nsize = pinned_empty((self.n,), np.int32)
        remap = pinned_empty((2, self.T), np.int32)
        enc = pinned_empty((self.c, self.ncol), np.int16)
        raw = pinned_empty((self.T, self.ncol), np.int16) if want_raw else None
        _capi.check(self._lib.subg_spg_export(self._h, nsize.ctypes.data, remap.ctypes.data, enc.ctypes.data,
                                              raw.ctypes.data if want_raw else None, _stream(self.device)))
        out = [nsize, remap, enc]
        if want_raw:
            out.append(raw)
        return out

    def enc_table(self) -> np.ndarray:
        """int16 [c+1, ncol] LP table with the all-zero row 0 (random_walks.py:81)."""
        p = C.c_void_p()
        _capi.check(self._lib.subg_spg_enc(self._h, _stream(self.device), C.byref(p)))
        if self.value_kind or not p.value or self.c == 0 or self.ncol < 1:
            return np.zeros((1, max(self.ncol, 0)), np.int16)
        enc = _view(p.value, (self.c, self.ncol), "<i2", self.device, self).cpu().numpy()  # no row compaction (views() would)
        return np.concatenate([np.zeros((1, self.ncol), np.int16), enc], axis=0)

    @property
    def pushes(self) -> int:
        """Forward pushes performed while sampling a PPR SpG (0 otherwise)."""
        v = C.c_int64(0)
        _capi.check(self._lib.subg_spg_pushes(self._h, C.byref(v)))
        return v.value

    def to_scipy(self):
        """The equivalent scipy CSR (what the reference's subg_matrix returns)."""
        import scipy.sparse as sp
        v = self.views()
        return sp.csr_matrix((v["data"].cpu().numpy(), v["indices"].cpu().numpy(), v["indptr"].cpu().numpy()),
                             shape=self.shape)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.subg_spg_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
