"""Drop-in for the reference's `subg_acc` CPython extension (subg_acc/subg_acc.c:1036-1059),
backed by the CUDA library through the C ABI.  Same function names, keyword names,
return layout and exception classes; `sampler/random_walks.py:18` can import it unchanged:

    from subg_acc import gset_sampler, walk_sampler

RNG: the reference draws from glibc rand_r on one word shared by all OpenMP threads
(subg_acc.c:731-732), so it is only reproducible with nthread=1.  Here
  nthread == 1  -> SUBG_RNG_RAND_R: that single-thread stream replayed in parallel on the
                   GPU, bit-identical output to the reference;
  otherwise     -> SUBG_RNG_PHILOX: counter-based Philox4x32-10 (same sampling law, fast path).
`SUBG_RNG=philox|rand_r` overrides the choice.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

from . import _capi
from .spg import DeviceGraph, SpG

__all__ = ["gset_sampler", "walk_sampler", "walk_join", "batch_sampler"]


def _say(msg: str) -> None:
    if os.environ.get("SUBG_QUIET", "0") != "1":
        print(msg, flush=True)


def _rng_mode(nthread: int) -> int:
    env = os.environ.get("SUBG_RNG", "").lower()
    if env == "philox":
        return _capi.SUBG_RNG_PHILOX
    if env == "rand_r":
        return _capi.SUBG_RNG_RAND_R
    return _capi.SUBG_RNG_RAND_R if nthread == 1 else _capi.SUBG_RNG_PHILOX


def gset_sampler(indptr, indices, query, num_walks=100, num_steps=3, bucket=-1, nthread=-1, seed=111413,
                 debug=-1, device="cuda"):
    """Walk-based node-set sampling with the LP structure encoder (subg_acc.c:649-1034).

    Returns the list [nsize int32[n], remap int32[2,T], enc int16[c,num_steps+1]] and, when
    debug > 0, raw_enc int16[T,num_steps+1] -- the reference's layout (subg_acc.c:1017-1024).
    """
    try:
        indptr = np.asarray(indptr)
        indices = np.asarray(indices)
        query = np.asarray(query).astype(np.int32, copy=False)  # NPY_ARRAY_FORCECAST, subg_acc.c:673
        num_walks, num_steps, bucket, nthread, seed, debug = (int(num_walks), int(num_steps), int(bucket),
                                                              int(nthread), int(seed), int(debug))
    except Exception as e:  # subg_acc.c:656-660
        raise TypeError("Input parsing error.\n") from e
    t0 = time.perf_counter()
    graph = DeviceGraph(indptr, indices, device)
    spg = SpG.sample(graph, query, num_walks=num_walks, num_steps=num_steps, bucket=bucket,
                     seed=seed & 0xFFFFFFFF, rng_mode=_rng_mode(nthread))
    t1 = time.perf_counter()
    n = max(len(query), 1)
    stride = num_walks * num_steps + 1 if bucket < 0 else bucket
    if spg.status & _capi.STATUS_BUCKET_OVERFLOW:  # subg_acc.c:835-836 (per key there; once here)
        _say(f"#SubGAcc: some keys exceed the buffer, try a larger bucket size > {stride}.")
    if spg.status & _capi.STATUS_DEAD_END:
        _say("#SubGAcc: rand_r replay met a node without out-neighbours; output is a valid sample "
             "but no longer the reference's nthread=1 stream.")
    _say(f"#SubGAcc: #total {spg.T}; #max_set {spg.max_set} of {stride}; "
         f"buffer usage {spg.T / n / stride * 100:.2f}%; dT_w {t1 - t0:.2f}s")
    out = spg.export_reference(want_raw=debug > 0)
    t2 = time.perf_counter()
    if os.environ.get("SUBG_PROFILE_HOST"):
        print(f"[subg host ms] gset_sampler: graph+sample={1e3 * (t1 - t0):.1f} export={1e3 * (t2 - t1):.1f}", file=sys.stderr, flush=True)
    _say(f"#SubGAcc: #enc_unique {spg.c}; compression ratio {spg.T / max(spg.c, 1):.2f}, dT_e {t2 - t1:.2f}s")
    spg.close()
    graph.close()
    return out


def walk_sampler(ptr, neighs, query, num_walks=100, num_steps=3, nthread=-1, seed=111413, replacement=-1, device="cuda"):
    """SUREL-v1 walk sampler + relative-position encoder (subg_acc.c:144-389).

    Returns [walks, obj] as the reference: walks int32 [n, num_walks*(num_steps+1)] (column 0 of every
    walk is the seed) and obj, an object array [n, 2] with obj[i,0] = the unique nodes of seed i's
    walks (int32, first-visit order of the step-major scan, root first) and obj[i,1] = their int32
    [count, num_steps+1] landing counts.  The reference parses `replacement` with the 'p' (predicate)
    format: truthy selects the first hop WITHOUT replacement (subg_acc.c:359-367), the default -1
    every hop with replacement.  nthread == 1 replays the reference's rand_r stream bit for bit.
    The rows of obj are views into two arrays, obj.ids / obj.rpe are not copied per seed."""
    import ctypes as C
    from .spg import _dev_index, _stream, pinned_empty
    try:
        ptr = np.asarray(ptr)
        neighs = np.asarray(neighs)
        query = np.ascontiguousarray(np.asarray(query).astype(np.int32, copy=False))
        num_walks, num_steps, nthread, seed = int(num_walks), int(num_steps), int(nthread), int(seed)
        without = 1 if (replacement is not None and replacement != -1 and bool(replacement)) else 0
    except Exception as e:  # subg_acc.c:323-327
        raise TypeError("Input parsing error.\n") from e
    lib = _capi.load()
    graph = DeviceGraph(ptr, neighs, device)
    dev = graph.device
    h = C.c_void_p()
    n = query.size
    try:
        _capi.check(lib.subg_walk_sample(graph._h, query.ctypes.data, n, num_walks, num_steps, seed & 0xFFFFFFFF,
                                         _rng_mode(nthread), without, _stream(dev), C.byref(h)))
        T = C.c_int64()
        st = C.c_uint32()
        _capi.check(lib.subg_walkset_info(h, None, C.byref(T), None, None, C.byref(st)))
        T = T.value
        if st.value & _capi.STATUS_DEAD_END:
            _say("#SubGAcc: rand_r replay met a node without out-neighbours; output is a valid sample "
                 "but no longer the reference's nthread=1 stream.")
        ncol = num_steps + 1
        walks = pinned_empty((n, num_walks * ncol), np.int32)
        off = pinned_empty((n + 1,), np.int64)
        ids = pinned_empty((T,), np.int32)
        rpe = pinned_empty((T, ncol), np.int32)
        _capi.check(lib.subg_walkset_export(h, walks.ctypes.data, off.ctypes.data, ids.ctypes.data, rpe.ctypes.data,
                                            _stream(dev)))
    finally:
        if h.value:
            lib.subg_walkset_free(h)
        graph.close()
    obj = np.empty((n, 2), dtype=object)
    o = off.tolist()
    for i in range(n):
        obj[i, 0] = ids[o[i]:o[i + 1]]
        obj[i, 1] = rpe[o[i]:o[i + 1]]
    return [walks, obj]


def walk_join(walk, key, query, nthread=-1, return_idx=-1, device="cuda"):
    """SUREL-v1 walk joining (subg_acc.c:509-647): for every query (u, v) and every walk position, the index of
    the visited node in the concatenated key sets, looked up in the set of u and in the set of v.
    walk: int32 [n, stride] (or [n, M, m+1]); key: sequence of n node-id arrays (walk_sampler's obj[:,0]);
    query: int [Q, 2].  Returns int32 [2, Q*2*stride], and [that, xq int32 [Q,2]] when return_idx is truthy (the
    reference parses it with the 'p' predicate format; left at its default -1 the index is not returned)."""
    import ctypes as C
    from .spg import _dev_index, _stream
    try:
        walk = np.ascontiguousarray(np.asarray(walk), dtype=np.int32)
        if walk.ndim < 2:
            raise ValueError("walk must be at least 2-D")
        n = walk.shape[0]
        stride = int(walk.shape[1] * walk.shape[2]) if walk.ndim > 2 else int(walk.shape[1])   # subg_acc.c:529-530
        keys = [np.asarray(k).astype(np.int32, copy=False).ravel() for k in key]
        q = np.ascontiguousarray(np.asarray(query), dtype=np.int32)
    except Exception as e:
        raise TypeError("Input parsing error.\n") from e
    if len(keys) != n:
        raise AssertionError("Dims do not match between num of walks and keys.\n")                # subg_acc.c:536-540
    off = np.zeros(n + 1, np.int64)
    np.cumsum([len(k) for k in keys], out=off[1:])
    ids = np.ascontiguousarray(np.concatenate(keys)) if n else np.zeros(0, np.int32)
    Q = q.shape[0]
    out = np.empty((2, Q * 2 * stride), np.int32)
    xq = np.empty(q.shape, np.int32)
    dev = _dev_index(device)
    _capi.check(_capi.load().subg_walk_join(walk.ctypes.data, n, stride, off.ctypes.data, ids.ctypes.data, q.ctypes.data, Q,
                                            out.ctypes.data, xq.ctypes.data, dev, _stream(dev)))
    want = bool(return_idx) and not (isinstance(return_idx, (int, np.integer)) and not isinstance(return_idx, bool) and return_idx == -1)
    return [out, xq] if want else out


def batch_sampler(ptr, neighs, query, num_walks=200, num_steps=8, thld=1000, nthread=-1, seed=111413, device="cuda", pid=None):
    """subg_acc.c:391-507 (kwlist :397): the serial walk-based mini-batch node sampler.  Returns the int32 array of the
    distinct nodes visited, in insertion order.  The reference seeds its single rand_r stream with seed + getpid()
    (:423) -- so does this (`pid=` overrides the process id for reproducible runs): within one process the result equals
    the reference's bit for bit (`subg_batch_sample`; one warp replays the stream on the device)."""
    import ctypes as C
    from .spg import _stream
    del nthread  # the reference parses it and never uses it
    if isinstance(ptr, DeviceGraph):
        g, own = ptr, False
    else:
        g, own = DeviceGraph(np.asarray(ptr), np.asarray(neighs), device), True
    try:
        q = np.ascontiguousarray(np.asarray(query).reshape(-1), dtype=np.int32)
        lib = _capi.load()
        state = (int(seed) + (os.getpid() if pid is None else int(pid))) & 0xFFFFFFFF
        cap = int(min(g.N, len(q) * (int(num_walks) * int(num_steps) + 1))) + 1
        out = np.empty(cap, np.int32)
        cnt = C.c_int64(0)
        _capi.check(lib.subg_batch_sample(g._h, q.ctypes.data, len(q), int(num_walks), int(num_steps), int(thld), state,
                                          out.ctypes.data, cap, C.byref(cnt), _stream(g.device)))
        return out[:cnt.value].copy()
    finally:
        if own:
            g.close()
