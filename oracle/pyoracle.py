"""ctypes front-end of oracle/subg_oracle.c plus numpy/scipy restatements of the
Python parts of the reference hot path.  TEST INFRASTRUCTURE ONLY (see
oracle/__init__.py).  Citations are file:line relative to /root/reference.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libsubg_oracle.so")
_lib = None

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the C restatement and (if /root/reference exists) the reference."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build(ref=False)
    L = C.CDLL(_LIB_PATH)
    L.orc_rand_r.restype = C.c_int
    L.orc_rand_r.argtypes = [C.POINTER(C.c_uint32)]
    L.orc_walks_rand_r.restype = C.c_uint32
    L.orc_walks_rand_r.argtypes = [_i64p, _i32p, _i32p, C.c_int64, C.c_int, C.c_int, C.c_uint32, _i32p, _i64p]
    L.orc_gset_from_walks.restype = C.c_int
    L.orc_gset_from_walks.argtypes = [_i32p, C.c_int64, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p,
                                      C.c_void_p, _i16p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32)]
    L.orc_spg_build.restype = C.c_int
    L.orc_spg_build.argtypes = [C.c_int64, _i32p, C.c_int64, _i32p, _i32p, _i32p, _i64p, _i32p, _i32p]
    L.orc_spjoin_pair_i32.restype = C.c_int64
    L.orc_spjoin_pair_i32.argtypes = [_i64p, _i32p, _i32p, _i64p, C.c_int64, _i32p, _i64p, _i64p]
    L.orc_spjoin_pair_f64.restype = C.c_int64
    L.orc_spjoin_pair_f64.argtypes = [_i64p, _i32p, _f64p, _i64p, C.c_int64, _f64p, _i64p, _i64p]
    L.orc_spjoin_triplet_i32.restype = C.c_int64
    L.orc_spjoin_triplet_i32.argtypes = [_i64p, _i32p, _i32p, _i64p, C.c_int64, _i32p, _i64p]
    L.orc_ppr_push.restype = C.c_int64
    L.orc_ppr_push.argtypes = [_i64p, _i32p, _i64p, C.c_int32, C.c_float, C.c_float, _i32p, _f32p, C.c_int64,
                               C.POINTER(C.c_int64)]
    L.orc_ppr_push_many.restype = C.c_int
    L.orc_ppr_push_many.argtypes = [_i64p, _i32p, _i64p, _i32p, C.c_int64, C.c_float, C.c_float, _i32p, _f32p,
                                    C.c_int64, _i64p, _i64p, C.c_int]
    L.orc_walk_sampler_walks.restype = C.c_uint32
    L.orc_walk_sampler_walks.argtypes = [_i64p, _i32p, _i32p, C.c_int64, C.c_int, C.c_int, C.c_uint32, C.c_int, _i32p]
    L.orc_walk_join.restype = C.c_int
    L.orc_walk_join.argtypes = [_i32p, C.c_int64, C.c_int64, _i64p, _i32p, _i32p, C.c_int64, _i32p, _i32p]
    L.orc_rpe_encode.restype = C.c_int64
    L.orc_rpe_encode.argtypes = [_i32p, C.c_int64, C.c_int, C.c_int, _i64p, _i32p, _i32p, C.c_int64]
    L.orc_batch_sampler.restype = C.c_int64
    L.orc_batch_sampler.argtypes = [_i64p, _i32p, C.c_int64, _i32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint32, _i32p, C.c_int64]
    _lib = L
    return L


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# --------------------------------------------------------------------------- RNG
def rand_r_stream(seed: int, k: int) -> np.ndarray:
    """k successive glibc rand_r outputs from `seed` (subg_acc.c:771,807 call sites)."""
    st = C.c_uint32(seed & 0xFFFFFFFF)
    return np.array([lib().orc_rand_r(C.byref(st)) for _ in range(k)], dtype=np.int64)


# ------------------------------------------------------------------ walk sampler
def walks_rand_r(indptr, indices, query, M: int, m: int, seed: int = 111413):
    """Walk traces of the reference's single-thread rand_r stream (subg_acc.c:745-809).
    Returns (walks int32[n,M,m], calls int64[n])."""
    q = _c(query, np.int32)
    walks = np.empty((len(q), M, m), np.int32)
    calls = np.zeros(len(q), np.int64)
    lib().orc_walks_rand_r(_c(indptr, np.int64), _c(indices, np.int32), q, len(q), M, m,
                           seed & 0xFFFFFFFF, walks, calls)
    return walks, calls


def gset_from_walks(query, walks, M: int, m: int, bucket: int = -1, want_raw: bool = False):
    """Trace-driven restatement of set_sampler (subg_acc.c:778-1000).
    Returns dict(nsize int32[n], remap int32[2,T], enc int16[c,m+1], raw, dropped)."""
    q = _c(query, np.int32)
    n = len(q)
    stride = M * m + 1 if bucket < 0 else bucket
    ncol = m + 1
    cap = max(n * stride, 1)
    nsize = np.zeros(n, np.int32)
    nidx = np.zeros(cap, np.int32)
    sf = np.zeros(cap, np.int32)
    raw = np.zeros((cap, ncol), np.int16) if want_raw else None
    enc = np.zeros((cap, ncol), np.int16)
    T = C.c_int64(0)
    c = C.c_int32(0)
    dropped = C.c_int32(0)
    rc = lib().orc_gset_from_walks(q, n, M, m, bucket, _c(walks, np.int32).reshape(-1), nsize, nidx, sf,
                                   raw.ctypes.data if want_raw else None, enc.reshape(-1), cap,
                                   C.byref(T), C.byref(c), C.byref(dropped))
    if rc == -2:
        raise AssertionError("Longer width of type for hasing key needed > INT64.")  # subg_acc.c:913
    if rc != 0:
        raise MemoryError(f"oracle gset_from_walks failed rc={rc}")
    T, c = T.value, c.value
    return dict(nsize=nsize, remap=np.stack([nidx[:T], sf[:T]]), enc=enc[:c].copy(),
                raw=raw[:T].copy() if want_raw else None, dropped=bool(dropped.value))


def gset_sampler_replay(indptr, indices, query, num_walks=100, num_steps=3, bucket=-1, seed=111413, debug=-1):
    """== reference gset_sampler(..., nthread=1) [subg_acc.c:649-1034], except that the
    node id of an isolated seed is the seed itself (the reference leaves it unwritten)."""
    walks, _ = walks_rand_r(indptr, indices, query, num_walks, num_steps, seed)
    r = gset_from_walks(query, walks, num_walks, num_steps, bucket, want_raw=debug > 0)
    out = [r["nsize"], r["remap"], r["enc"]]
    if debug > 0:
        out.append(r["raw"])
    return out


# ------------------------------------------------------- SUREL-v1 walk_sampler
def walk_sampler(ptr, neighs, query, num_walks=100, num_steps=3, seed=111413, replacement=-1):
    """== reference walk_sampler(..., nthread=1) [subg_acc.c:144-389]: returns
    [walks int32[n, M*(m+1)], obj[n,2]] with obj[i,0] = node ids in first-visit order of the
    step-major scan and obj[i,1] = int32 [count, m+1] relative-position counts.
    `replacement` truthy selects the first hop WITHOUT replacement (subg_acc.c:359-367; the
    reference parses it with the 'p' predicate format, default -1 = with replacement)."""
    q = _c(query, np.int32)
    n, M, m = len(q), int(num_walks), int(num_steps)
    without = 1 if (replacement is not None and replacement != -1 and bool(replacement)) else 0
    walks = np.empty((n, M * (m + 1)), np.int32)
    lib().orc_walk_sampler_walks(_c(ptr, np.int64), _c(neighs, np.int32), q, n, M, m, seed & 0xFFFFFFFF, without,
                                 walks.reshape(-1))
    off, ids, rpe = rpe_encode(walks, M, m)
    obj = np.empty((n, 2), dtype=object)
    for i in range(n):
        obj[i, 0] = ids[off[i]:off[i + 1]]
        obj[i, 1] = rpe[off[i]:off[i + 1]]
    return [walks, obj]


def walk_join(walk, key, query, return_idx=False):
    """== reference walk_join(walk, key, query) [subg_acc.c:509-647].  Returns out int32 [2, Q*2*stride]
    (and xq int32 [Q,2] when return_idx)."""
    walk = _c(walk, np.int32)
    n = walk.shape[0]
    stride = int(np.prod(walk.shape[1:]))
    keys = [np.asarray(k, dtype=np.int32).ravel() for k in key]
    assert len(keys) == n, "Dims do not match between num of walks and keys."
    off = np.zeros(n + 1, np.int64)
    np.cumsum([len(k) for k in keys], out=off[1:])
    ids = _c(np.concatenate(keys), np.int32) if n else np.zeros(0, np.int32)
    q = _c(query, np.int32)
    Q = q.shape[0]
    out = np.zeros((2, Q * 2 * stride), np.int32)
    xq = np.zeros((Q, 2), np.int32)
    rc = lib().orc_walk_join(walk.reshape(-1), n, stride, off, ids, q.reshape(-1), Q, out.reshape(-1), xq.reshape(-1))
    if rc != 0:
        raise MemoryError(f"oracle walk_join failed rc={rc}")
    return [out, xq] if return_idx else out


def rpe_encode(walks, M: int, m: int):
    """rpe_encoder over every seed (subg_acc.c:250-314) -> (off int64[n+1], ids int32[T], rpe int32[T,m+1])."""
    walks = _c(walks, np.int32).reshape(-1, M * (m + 1))
    n = walks.shape[0]
    cap = max(n * (M * m + 1), 1)
    off = np.zeros(n + 1, np.int64)
    ids = np.zeros(cap, np.int32)
    rpe = np.zeros((cap, m + 1), np.int32)
    T = lib().orc_rpe_encode(walks.reshape(-1), n, M, m, off, ids, rpe.reshape(-1), cap)
    if T < 0:
        raise MemoryError(f"oracle rpe_encode failed rc={T}")
    return off, ids[:T].copy(), rpe[:T].copy()


# -------------------------------------------------------------------- SpG build
def subg_matrix_from(nsize, remap, enc, query, N: int, num_steps: int):
    """sampler/random_walks.py:79-81: SpG = csr((sfptr+1, (repeat(seed,nsize), node)));
    enc gets an all-zero row 0.  Returns (scipy csr int32, enc int16[c+1,num_steps])."""
    z = sp.csr_matrix((remap[1] + 1, (np.repeat(np.asarray(query), nsize), remap[0])), shape=(N, N))
    assert z.has_sorted_indices
    enc0 = np.concatenate([np.zeros((1, num_steps), enc.dtype), enc], axis=0)
    return z, enc0


def spg_build(N: int, query, nsize, nidx, sfptr):
    """C flavour of the same build: returns (indptr int64[N+1], indices int32[T], data int32[T])."""
    q = _c(query, np.int32)
    T = int(np.sum(nsize))
    indptr = np.zeros(N + 1, np.int64)
    indices = np.zeros(max(T, 1), np.int32)
    data = np.zeros(max(T, 1), np.int32)
    rc = lib().orc_spg_build(N, q, len(q), _c(nsize, np.int32), _c(nidx, np.int32), _c(sfptr, np.int32),
                             indptr, indices, data)
    assert rc == 0
    return indptr, indices[:T], data[:T]


# ---------------------------------------------------------------------- SpJoin
def _csr_parts(x):
    if sp.issparse(x):
        return _c(x.indptr, np.int64), _c(x.indices, np.int32), x.data
    return _c(x[0], np.int64), _c(x[1], np.int32), x[2]


def spjoin_pair(x, edge):
    """Merge-join restatement of train.py:75-85 (bgather).  x: scipy CSR or
    (indptr, indices, data).  Returns (xz [N,2], sizes_l[B], sizes_r[B]); rows are
    all left sets then all right sets (train.py:34-36)."""
    indptr, indices, data = _csr_parts(x)
    e = _c(edge, np.int64)
    B = e.shape[1]
    sz = np.diff(indptr)
    N = int(sz[e[0]].sum() + sz[e[1]].sum())
    sl = np.zeros(B, np.int64)
    sr = np.zeros(B, np.int64)
    if np.issubdtype(np.asarray(data).dtype, np.floating):
        out = np.zeros((max(N, 1), 2), np.float64)
        n = lib().orc_spjoin_pair_f64(indptr, indices, _c(data, np.float64), e.reshape(-1), B, out.reshape(-1), sl, sr)
    else:
        out = np.zeros((max(N, 1), 2), np.int32)
        n = lib().orc_spjoin_pair_i32(indptr, indices, _c(data, np.int32), e.reshape(-1), B, out.reshape(-1), sl, sr)
    assert n == N
    return out[:N], sl, sr


def spjoin_triplet(x, hedge):
    """Restatement of train.py:48-72 (hgather) index part: blocks [u|w],[w|u],[v|w],[w|v]."""
    indptr, indices, data = _csr_parts(x)
    e = _c(hedge, np.int64)
    B = e.shape[1]
    sz = np.diff(indptr)
    N = int(sz[e[0]].sum() + sz[e[1]].sum() + 2 * sz[e[2]].sum())
    sizes = np.zeros(4 * B, np.int64)
    out = np.zeros((max(N, 1), 2), np.int32)
    n = lib().orc_spjoin_triplet_i32(indptr, indices, _c(data, np.int32), e.reshape(-1), B, out.reshape(-1), sizes)
    assert n == N
    return out[:N], sizes


def pair_index(sizes_l, sizes_r, ptr: bool = True):
    """train.py:20-30 / :104-109: CSR-style indptr [2B+1] or per-row segment ids [N]."""
    sizes = np.concatenate([sizes_l, sizes_r]).astype(np.int64)
    if ptr:
        return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    return np.repeat(np.arange(len(sizes), dtype=np.int64), sizes)


def scipy_pair_join(edge, x):
    """The reference's own formulation (scipy CSR algebra), restated for use as the
    SpJoin CPU baseline 'port': x_b*(x_a>0) + (x_a>0) - 1   (train.py:77-84)."""
    a, b = x[edge[0]], x[edge[1]]
    ma, mb = a > 0, b > 0
    ba = b.multiply(ma) + ma
    ab = a.multiply(mb) + mb
    left = np.stack([a.data, ba.data - 1]).T
    right = np.stack([b.data, ab.data - 1]).T
    return left, right, ma.getnnz(axis=1), mb.getnnz(axis=1)


def scipy_pgather(edge, x, njobs: int = 4):
    """pgather (train.py:88-111) restated for the SpJoin CPU baseline: the batch is split into `njobs` column blocks,
    one Python thread per block runs the bgather formulation (scipy_pair_join), and the blocks are concatenated
    [all left | all right] with the cumulative segment pointer.  Returns (xz int [N,2], indptr int64 [2B+1])."""
    import threading
    blocks = np.array_split(np.asarray(edge), njobs, axis=1)
    out = [None] * njobs

    def work(i):
        out[i] = scipy_pair_join(blocks[i], x)
    th = [threading.Thread(target=work, args=(i,)) for i in range(njobs)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    xz = np.vstack([*[o[0] for o in out], *[o[1] for o in out]])
    indptr = np.cumsum(np.concatenate([[0], *[o[2] for o in out], *[o[3] for o in out]])).astype(np.int64)
    return xz, indptr


def scipy_triplet_join(hedge, x):
    """hgather (train.py:48-72) restated with the same scipy CSR algebra: segments [u|w, w|u, v|w, w|v]."""
    u, v, w = x[hedge[0]], x[hedge[1]], x[hedge[2]]
    mu, mv, mw = u > 0, v > 0, w > 0
    parts = []
    for a, b, ma, mb in ((u, w, mu, mw), (v, w, mv, mw)):
        ba = b.multiply(ma) + ma
        ab = a.multiply(mb) + mb
        parts.append(np.stack([a.data, ba.data - 1]).T)
        parts.append(np.stack([b.data, ab.data - 1]).T)
    sizes = np.concatenate([mu.getnnz(axis=1), mw.getnnz(axis=1), mv.getnnz(axis=1), mw.getnnz(axis=1)])
    return np.vstack(parts), np.repeat(np.arange(len(sizes), dtype=np.int64), sizes)


# ------------------------------------------------------------------------- PPR
def ppr_push(indptr, indices, deg, node: int, alpha: float, eps: float, cap: int = 1 << 16):
    """sampler/pprgo.py:9-38 for one seed: (keys int64, vals float32) in p-insertion order."""
    keys = np.zeros(cap, np.int32)
    vals = np.zeros(cap, np.float32)
    npush = C.c_int64(0)
    k = lib().orc_ppr_push(_c(indptr, np.int64), _c(indices, np.int32), _c(deg, np.int64), int(node),
                           np.float32(alpha), np.float32(eps), keys, vals, cap, C.byref(npush))
    if k < 0:
        return ppr_push(indptr, indices, deg, node, alpha, eps, cap * 8)
    return keys[:k].astype(np.int64), vals[:k].copy(), npush.value


def ppr_push_many(indptr, indices, deg, seeds, alpha, eps, cap=4096, nthread=1):
    seeds = _c(seeds, np.int32)
    n = len(seeds)
    while True:
        keys = np.zeros((n, cap), np.int32)
        vals = np.zeros((n, cap), np.float32)
        cnt = np.zeros(n, np.int64)
        pushes = np.zeros(n, np.int64)
        rc = lib().orc_ppr_push_many(_c(indptr, np.int64), _c(indices, np.int32), _c(deg, np.int64), seeds, n,
                                     np.float32(alpha), np.float32(eps), keys.reshape(-1), vals.reshape(-1),
                                     cap, cnt, pushes, nthread)
        if rc == 0:
            return keys, vals, cnt, pushes
        cap *= 4


def ppr_topk_rows(keys, vals, cnt, topk: int):
    """pprgo.py:58-61: per seed keep argsort(val)[-topk:].  Ties at the k-th score are
    unspecified in the reference (unstable sort); here the stable order is used and
    `tie_mask` marks entries equal to the k-th score so callers can exclude them."""
    rows = []
    for i in range(len(cnt)):
        k, v = keys[i, :cnt[i]], vals[i, :cnt[i]]
        order = np.argsort(v, kind="stable")[-topk:]
        kth = v[order[0]] if len(order) else np.float32(0)
        rows.append((k[order].astype(np.int64), v[order], v[order] == kth if cnt[i] > topk else np.zeros(len(order), bool)))
    return rows


def topk_ppr_matrix(adj, alpha, eps, idx, topk, normalization="row", nthread=1):
    """Restatement of pprgo.py:65-111 (ppr_topk + construct_sparse + normalisation)."""
    idx = np.asarray(idx)
    deg_cnt = np.asarray((adj > 0).sum(axis=1)).ravel().astype(np.int64)  # pprgo.py:68
    keys, vals, cnt, _ = ppr_push_many(adj.indptr, adj.indices, deg_cnt, idx, alpha, eps, nthread=nthread)
    rows = ppr_topk_rows(keys, vals, cnt, topk)
    i = np.repeat(np.arange(len(idx)), [len(r[0]) for r in rows])
    j = np.concatenate([r[0] for r in rows])
    w = np.concatenate([r[1] for r in rows])
    mat = sp.coo_matrix((w, (i, j)), (len(idx), adj.shape[0])).tocsr()
    if normalization == "sym":  # pprgo.py:87-96
        deg = np.asarray(adj.sum(1)).ravel()
        dsq = np.sqrt(np.maximum(deg, 1e-12))
        dinv = 1.0 / dsq
        r, c = mat.nonzero()
        mat.data = dsq[idx[r]] * mat.data * dinv[c]
    elif normalization == "col":  # pprgo.py:97-106
        deg = np.asarray(adj.sum(1)).ravel()
        dinv = 1.0 / np.maximum(deg, 1e-12)
        r, c = mat.nonzero()
        mat.data = deg[idx[r]] * mat.data * dinv[c]
    elif normalization != "row":
        raise ValueError(f"Unknown PPR normalization: {normalization}")
    return mat


# -------------------------------------------------------------------- encoders
def encoding_ppr(x):
    """utils.py:35-36: affine rescale by the global max."""
    x = x.copy()
    x.data = (x.data + 0.1) / (x.data.max() + 0.1)
    return x


def encoding_spd(x, adj):
    """utils.py:29-34: 1*[w in N(u)] + 0.5*[w in S_u and w in N2(u)] + 0.3*[w in S_u], diag 2.3."""
    import warnings
    x0 = x > 0
    x1 = adj > 0
    x2 = x1 @ x1  # boolean matrix square (bool dtype keeps it a reachability test)
    out = x1 + x0.multiply(x2 * 0.5) + x0 * 0.3
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out.setdiag(2.3)
    out = out.tocsr()
    out.sort_indices()
    return out


def batch_sampler(ptr, neighs, query, num_walks=200, num_steps=8, thld=1000, seed=111413, pid=0):
    """batch_sampler (subg_acc.c:391-507): distinct nodes, in insertion order, of the serial walk-based mini-batch
    sampler; the rand_r stream starts at seed + pid (the reference adds getpid(), :423)."""
    ptr = _c(ptr, np.int64)
    neighs = _c(neighs, np.int32)
    query = _c(np.asarray(query).reshape(-1), np.int32)
    N = len(ptr) - 1
    cap = int(min(N, len(query) * (num_walks * num_steps + 1))) + 1
    out = np.empty(cap, np.int32)
    cnt = lib().orc_batch_sampler(ptr, neighs, N, query, len(query), num_walks, num_steps, thld, (seed + pid) & 0xFFFFFFFF, out, cap)
    assert cnt >= 0, cnt
    return out[:cnt].copy()
