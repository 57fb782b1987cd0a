"""Host mirror of the SpJoin operators of the reference's train.py (gather :13-45,
hgather :48-72, bgather :75-85, pgather :88-111), running on the device SpG.

Same names, argument meaning and return values: `x` may be a surel_plus_b200.SpG (stays in
HBM) or the scipy CSR the reference builds (uploaded once and cached); `edge` is the [2,B] /
[3,B] int64 index tensor (torch CPU/CUDA tensor or numpy); the outputs are device tensors
consumed by Net.forward(x, ptr) (model.py:76-90).
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
import torch

from . import _capi
from .spg import SpG, _dev_index, _stream

_uploaded: "weakref.WeakValueDictionary[int, SpG]" = weakref.WeakValueDictionary()
_upload_keep: dict = {}


def _fingerprint(x, device):
    """What identifies an uploaded scipy matrix besides the object itself: its three buffers and the target device.
    `utils.encoding(x, ..., 'PPR')` assigns a NEW `x.data` array (utils.py:36), which changes the fingerprint and
    forces a fresh upload; a write INTO the same buffer cannot be seen from here (invalidate_uploads())."""
    try:
        return (x.data.ctypes.data, x.indices.ctypes.data, x.indptr.ctypes.data, int(x.nnz), str(device))
    except AttributeError:
        return (None, None, None, -1, str(device))


def invalidate_uploads() -> None:
    """Forget the device copies of scipy matrices passed to gather / pgather / hgather (call after modifying a
    matrix's buffers in place)."""
    _upload_keep.clear()


def _as_spg(x, device) -> SpG:
    if isinstance(x, SpG):
        return x
    key = id(x)
    fp = _fingerprint(x, device)
    hit = _upload_keep.get(key)
    if hit is not None and hit[0]() is x and hit[2] == fp:
        return hit[1]
    spg = SpG.from_scipy(x, device)
    try:
        _upload_keep[key] = (weakref.ref(x, lambda _r, k=key: _upload_keep.pop(k, None)), spg, fp)
    except TypeError:
        pass
    return spg


def _edge_arg(edge, arity: int):
    """-> (object keeping the memory alive, raw pointer, B, is_device)."""
    if isinstance(edge, torch.Tensor):
        e = edge.to(torch.int64).contiguous()
        if e.dim() != 2 or e.shape[0] != arity:
            raise TypeError(f"edge must have shape [{arity}, B]")
        return e, e.data_ptr(), e.shape[1], e.is_cuda
    e = np.ascontiguousarray(np.asarray(edge), dtype=np.int64)
    if e.ndim != 2 or e.shape[0] != arity:
        raise TypeError(f"edge must have shape [{arity}, B]")
    return e, e.ctypes.data, e.shape[1], False


def _join(edge, x, device, arity, encode=None, want_segid=False, want_ptr=True):
    """One SpJoin batch.  The output is sized from the rows-per-segment seen on this SpG so far (+25 %), so
    that sizes, scan and join run back to back with a single host synchronisation (subg_spjoin); the first
    batch, and a batch that outgrows the estimate, take the two-step plan / run route with an exact buffer."""
    spg = _as_spg(x, device)
    lib = _capi.load()
    dev = spg.device
    keep, eptr, B, on_dev = _edge_arg(edge, arity)
    tdev = torch.device("cuda", dev)
    nseg = (2 if arity == 2 else 4) * B
    edge_dev = keep if on_dev else torch.empty((arity, B), dtype=torch.int64, device=tdev)
    indptr = torch.empty(nseg + 1, dtype=torch.int64, device=tdev)
    st = _stream(dev)
    table = None
    if spg.value_kind == 1:
        if encode is not None:
            raise TypeError("a value SpG (PPR/SPD) is joined without an LP table")
        shape_tail, dtype, k = (2,), torch.float32, 0
    elif encode is not None:
        table = encode if (isinstance(encode, torch.Tensor) and encode.is_cuda and encode.dtype == torch.float32
                           and encode.is_contiguous()) else torch.as_tensor(encode, dtype=torch.float32, device=tdev).contiguous()
        k = table.shape[1]
        shape_tail, dtype = (2, k), torch.float32
    else:
        shape_tail, dtype, k = (2,), torch.int32, 0
    tptr = table.data_ptr() if table is not None else None
    if B == 0:  # nothing to join: empty rows, a single segment pointer
        indptr.zero_()
        return (torch.empty((0,) + shape_tail, dtype=dtype, device=tdev), indptr,
                torch.empty(0, dtype=torch.int64, device=tdev) if want_segid else None)
    N = C.c_int64(0)
    rate = getattr(spg, "_rows_per_seg", None)
    out = segid = None
    if rate is not None and nseg > 0:
        cap = int(rate * nseg * 1.25) + 1024
        out = torch.empty((cap,) + shape_tail, dtype=dtype, device=tdev)
        segid = torch.empty(cap, dtype=torch.int64, device=tdev) if want_segid else None
        ran = C.c_int(0)
        _capi.check(lib.subg_spjoin(spg._h, eptr, B, arity, edge_dev.data_ptr(), indptr.data_ptr(), tptr, k, out.data_ptr(),
                                    cap, segid.data_ptr() if segid is not None else None, C.byref(N), C.byref(ran), st))
        if ran.value:
            out = out[:N.value]
            segid = segid[:N.value] if segid is not None else None
        else:
            out = segid = None
    else:
        _capi.check(lib.subg_spjoin_plan(spg._h, eptr, B, arity, edge_dev.data_ptr(), indptr.data_ptr(), C.byref(N), st))
    N = N.value
    if nseg > 0:
        spg._rows_per_seg = max(rate or 0.0, N / nseg)
    if out is None:
        out = torch.empty((N,) + shape_tail, dtype=dtype, device=tdev)
        segid = torch.empty(N, dtype=torch.int64, device=tdev) if want_segid else None
        _capi.check(lib.subg_spjoin_run(spg._h, edge_dev.data_ptr(), B, arity, indptr.data_ptr(), tptr, k, out.data_ptr(),
                                        segid.data_ptr() if segid is not None else None, st))
    return out, indptr, segid


def gather(edge, x, device, ptr=True, encode=None):
    """train.py:13-45.  Returns (xz, indptr) with xz float32 [N,2,k] = encode[pointer pairs]
    (or [N,2,1] raw values when encode is None) and indptr int64 [2B+1] (ptr=True) or the
    per-row segment id int64 [N] (ptr=False)."""
    out, indptr, segid = _join(edge, x, device, 2, encode=encode, want_segid=not ptr)
    if encode is None:
        out = out.float().unsqueeze(dim=-1)
    return out, (indptr if ptr else segid)


def hgather(hedge, x, device, encode=None):
    """train.py:48-72.  Returns (xz float32 [N,2,k], ind int64 [N]) with segments
    [u|w, w|u, v|w, w|v] numbered 0..4B-1."""
    if encode is None:
        raise NotImplementedError  # train.py:69-70
    out, _, segid = _join(hedge, x, device, 3, encode=encode, want_segid=True)
    assert out.size(0) == segid.size(0)
    return out, segid


def bgather(edge, x, out):
    """train.py:75-85: fills out[0..3] = (xl [Su,2], xr [Sv,2], sizes_l [B], sizes_r [B]) as host
    numpy arrays (pointer pairs, no table lookup)."""
    spg = _as_spg(x, "cuda")
    xz, indptr, _ = _join(edge, spg, None, 2)
    B = (indptr.numel() - 1) // 2
    ip = indptr.cpu().numpy()
    xz = xz.cpu().numpy()
    nl = int(ip[B])
    out[0], out[1] = xz[:nl], xz[nl:]
    sizes = np.diff(ip)
    out[2], out[3] = sizes[:B], sizes[B:]


def pgather(edge, M, device, encode, gather_func=None, ptr=True, njobs=4):
    """train.py:88-111.  The reference splits the batch over `njobs` Python threads and
    re-concatenates [all left blocks, all right blocks]; that order equals gather()'s, so the
    device join is one launch and `gather_func` / `njobs` are accepted for signature
    compatibility only."""
    out, indptr, segid = _join(edge, M, device, 2, encode=encode, want_segid=not ptr)
    if encode is None:
        out = out.float().unsqueeze(dim=-1)
    return out, (indptr if ptr else segid)


class JoinStream:
    """The per-batch SpJoin of a training / evaluation loop without allocation or host synchronisation
    (subg_joiner_* in include/subg_b200.h): the reference calls gather / pgather / hgather once per mini-batch
    (train.py:121-127 with batch 1024, main_horder.py:33 with 2048 triplets), where the join kernel itself runs for
    10-20 us.  A JoinStream fixes (SpG, batch size, arity, LP table, output capacity), captures the batch as a CUDA
    graph per ring slot and `submit` is one pinned memcpy + one graph launch.

        js = JoinStream(z, batch_size=1024, device=dev, encode=xpe)            # pairs; arity=3 for triplets
        xz, indptr, nrows = js.submit(edge)          # device tensors, nothing synchronised
        xz, indptr = js.gather(edge)                 # same result with the reference's exact shapes (one sync)

    `xz` of submit has `capacity` rows; rows [0, nrows[0]) are the join, in the order and layout of gather; the segment
    pointer `indptr` (int64 [2B+1], or [4B+1] for triplets) addresses only those.  Host-edge batches alternate between two
    internal streams (the plan of batch k+1 runs beside the join of batch k; SUBG_JOIN_LANES=1 turns that off) and the
    caller's stream waits for each batch, so consumers just use the tensors on that stream.  The buffers of a submit are
    reused `depth` submits later: queue the work that reads them before the (depth-1)-th following submit (with the default
    depth of 3: submit(k), consume(k), or one batch of prefetch).  segid=True also returns the per-row segment id
    (gather's ptr=False / hgather's `ind`).  Batches are queued on the CUDA stream that was current when the JoinStream was
    built (or `stream=`)."""

    def __init__(self, x, batch_size: int, device="cuda", encode=None, arity: int = 2, capacity: int | None = None,
                 segid: bool = False, depth: int = 3, stream=None):
        from .spg import _view
        self._view = _view
        self.spg = _as_spg(x, device)
        self._lib = _capi.load()
        self.B, self.arity, self.depth, self.want_segid = int(batch_size), int(arity), int(depth), bool(segid)
        self.nseg = (2 if arity == 2 else 4) * self.B
        dev = self.spg.device
        tdev = torch.device("cuda", dev)
        self.table = None
        if self.spg.value_kind == 1:
            if encode is not None:
                raise TypeError("a value SpG (PPR/SPD) is joined without an LP table")
            self.tail, self.k, self.typestr = (2,), 0, "<f4"
        elif encode is not None:
            self.table = encode if (isinstance(encode, torch.Tensor) and encode.is_cuda and encode.dtype == torch.float32
                                    and encode.is_contiguous()) else torch.as_tensor(encode, dtype=torch.float32, device=tdev).contiguous()
            self.k = int(self.table.shape[1])
            self.tail, self.typestr = (2, self.k), "<f4"
        else:
            self.tail, self.k, self.typestr = (2,), 0, "<i4"
        if capacity is None:  # mean rows per segment x 1.5 + room for a few of the largest sets
            avg = self.spg.T / max(self.spg.n, 1)
            capacity = int(self.nseg * avg * 1.5) + 16 * max(self.spg.max_set, 1) + 1024
        self.capacity = int(capacity)
        self._h = C.c_void_p()
        _capi.check(self._lib.subg_joiner_create(self.spg._h, self.B, self.arity, self.table.data_ptr() if self.table is not None else None,
                                                 self.k, self.capacity, int(self.want_segid), self.depth, C.byref(self._h)))
        self._last_slot = -1
        # the stream the batches are queued on: the one current at construction (or `stream`); looked up once, not per batch
        self._st = (stream if stream is not None else torch.cuda.current_stream(dev)).cuda_stream
        self._views: dict = {}
        self._outv = [C.c_void_p() for _ in range(4)]
        self._out = tuple(C.byref(x) for x in self._outv)
        self._slot = C.c_int(0)
        self._slot_ref = C.byref(self._slot)
        self._submit = self._lib.subg_joiner_submit

    def submit(self, edge):
        # the hot call of a training loop: keep the host side to one ctypes call (the slot's tensors are built once)
        if isinstance(edge, np.ndarray) and edge.dtype == np.int64 and edge.flags.c_contiguous and edge.shape == (self.arity, self.B):
            keep, eptr, on_dev = edge, edge.ctypes.data, 0
        else:
            keep, eptr, B, on_dev = _edge_arg(edge, self.arity)
            if B != self.B:
                raise TypeError(f"this JoinStream joins batches of {self.B} queries")
        out, indptr, segid, nrows = self._out
        dev = self.spg.device
        rc = self._submit(self._h, eptr, int(on_dev), self._st, out, indptr, segid, nrows, self._slot_ref)
        if rc:
            _capi.check(rc)
        q = self._slot.value
        self._last_slot = q
        res = self._views.get(q)
        if res is None:
            v = self._view
            res = (v(self._outv[0].value, (self.capacity,) + self.tail, self.typestr, dev, self),
                   v(self._outv[1].value, (self.nseg + 1,), "<i8", dev, self),
                   v(self._outv[3].value, (2,), "<i8", dev, self))
            if self.want_segid:
                res = res + (v(self._outv[2].value, (self.capacity,), "<i8", dev, self),)
            self._views[q] = res
        return res

    def rows(self) -> int:
        """Row count of the last submit (waits for it)."""
        n = C.c_int64(0)
        _capi.check(self._lib.subg_joiner_rows(self._h, self._last_slot, C.byref(n)))
        return n.value

    def gather(self, edge, ptr: bool = True):
        """gather / pgather / hgather semantics (exact shapes) through the captured graph: one synchronisation."""
        res = self.submit(edge)
        try:
            n = self.rows()
        except MemoryError:   # the batch outgrew the slot: the two-step route with an exact buffer
            out, indptr, segid = _join(edge, self.spg, None, self.arity, encode=self.table, want_segid=not ptr or self.arity == 3)
            if self.table is None:
                out = out.float().unsqueeze(dim=-1)
            return out, (indptr if (ptr and self.arity == 2) else segid)
        xz = res[0][:n]
        if self.table is None:
            xz = xz.float().unsqueeze(dim=-1) if self.spg.value_kind == 0 else xz.unsqueeze(dim=-1)
        if ptr and self.arity == 2:
            return xz, res[1]
        if not self.want_segid:
            raise TypeError("segment ids were not requested (segid=True)")
        return xz, res[3][:n]

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.subg_joiner_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
