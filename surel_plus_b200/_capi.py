"""ctypes binding of include/subg_b200.h (libsubg_b200.so).

The product path has no CPU fallback: if the CUDA library is missing or no GPU is
visible, every entry point raises.  Error codes map to the exception classes the
reference raises (subg_acc/subg_acc.c:658, 688-721, 913, 1003).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
#: SUBG_LIB: an experiment build of the same library (surel_plus_b200.build --tag); measurement only
LIB_PATH = os.environ.get("SUBG_LIB") or os.path.join(_HERE, "_lib", "libsubg_b200.so")

SUBG_RNG_PHILOX, SUBG_RNG_RAND_R, SUBG_RNG_TRACE = 0, 1, 2
STATUS_BUCKET_OVERFLOW, STATUS_DEAD_END, STATUS_PPR_SECOND_PASS = 1, 2, 4
SAMPLE_NO_RANKS, SAMPLE_DUMP_WALKS, SAMPLE_NO_COMPACT = 1, 2, 4
ENCODER_NONE, ENCODER_PPR, ENCODER_SPD = 0, 1, 2
TIMING_SAMPLER, TIMING_SPJOIN, TIMING_BUILD, TIMING_PPR, TIMING_EXCHANGE = 0, 1, 2, 3, 4

#: every symbol include/subg_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "subg_abi_version", "subg_last_error",
    "subg_graph_create", "subg_graph_from_edges", "subg_graph_export", "subg_graph_info", "subg_graph_free",
    "subg_gset_sample", "subg_gset_sample_shard", "subg_spg_set_lp_table", "subg_spg_info", "subg_spg_export", "subg_spg_views", "subg_spg_rows", "subg_spg_enc", "subg_spg_walks", "subg_spg_expand_rows",
    "subg_spg_from_csr", "subg_spg_alloc", "subg_spg_seal", "subg_spg_free",
    "subg_xchg_create", "subg_xchg_export", "subg_xchg_open", "subg_xchg_slab", "subg_xchg_pack", "subg_xchg_assemble", "subg_xchg_stage", "subg_xchg_link", "subg_xchg_free",
    "subg_spjoin_plan", "subg_spjoin_run", "subg_spjoin",
    "subg_joiner_create", "subg_joiner_submit", "subg_joiner_rows", "subg_joiner_free",
    "subg_ppr_topk", "subg_spg_encode", "subg_spg_pushes",
    "subg_walk_sample", "subg_walkset_info", "subg_walkset_export", "subg_walkset_views", "subg_walkset_free", "subg_walk_join", "subg_batch_sample",
    "subg_timing_enable", "subg_timing_read", "subg_launch_count",
    "subg_trim_cache", "subg_host_alloc", "subg_host_free",
]

_lib = None


class SubgUnsupported(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m surel_plus_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, u64 = C.c_void_p, C.c_int64, C.c_int, C.c_uint64
    L.subg_abi_version.restype = i32
    L.subg_last_error.restype = C.c_char_p
    L.subg_graph_create.argtypes = [vp, i32, vp, i64, i64, i32, vp, C.POINTER(vp)]
    L.subg_graph_from_edges.argtypes = [vp, vp, i64, i64, i32, i32, i32, vp, C.POINTER(vp)]
    L.subg_graph_export.argtypes = [vp, vp, vp, vp]
    L.subg_graph_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    L.subg_graph_free.argtypes = [vp]
    L.subg_graph_free.restype = None
    L.subg_gset_sample.argtypes = [vp, vp, i64, i32, i32, i32, u64, i32, vp, i32, vp, C.POINTER(vp)]
    L.subg_gset_sample_shard.argtypes = [vp, vp, i64, i64, i64, i32, i32, i32, u64, i32, vp, i32, vp, C.POINTER(vp)]
    L.subg_spg_set_lp_table.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, vp]
    L.subg_spg_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
    L.subg_spg_export.argtypes = [vp, vp, vp, vp, vp, vp]
    L.subg_spg_views.argtypes = [vp, vp] + [C.POINTER(vp)] * 6
    L.subg_spg_enc.argtypes = [vp, vp, C.POINTER(vp)]
    L.subg_spg_expand_rows.argtypes = [vp, i64, vp]
    L.subg_spg_walks.argtypes = [vp, vp, C.POINTER(vp)]
    L.subg_spg_rows.argtypes = [vp] + [C.POINTER(vp)] * 4 + [C.POINTER(i64)]
    L.subg_spg_from_csr.argtypes = [vp, vp, vp, i32, i64, i64, i32, vp, C.POINTER(vp)]
    L.subg_spg_alloc.argtypes = [i64, i64, i32, vp, C.POINTER(vp)]
    L.subg_spg_seal.argtypes = [vp, vp]
    L.subg_spg_free.argtypes = [vp]
    L.subg_spg_free.restype = None
    L.subg_xchg_create.argtypes = [i32, i32, i32, i64, C.POINTER(vp)]
    L.subg_xchg_export.argtypes = [vp, vp]
    L.subg_xchg_open.argtypes = [vp, vp]
    L.subg_xchg_slab.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
    L.subg_xchg_pack.argtypes = [vp, vp, i64, vp, vp]
    L.subg_xchg_assemble.argtypes = [vp, vp, vp, i32, i32, vp, C.POINTER(vp)]
    L.subg_xchg_stage.argtypes = [vp, vp, i64, i64, vp, vp]
    L.subg_xchg_link.argtypes = [vp, vp, vp, i32, i32, vp, C.POINTER(vp)]
    L.subg_xchg_free.argtypes = [vp]
    L.subg_xchg_free.restype = None
    L.subg_spjoin_plan.argtypes = [vp, vp, i64, i32, vp, vp, C.POINTER(i64), vp]
    L.subg_spjoin_run.argtypes = [vp, vp, i64, i32, vp, vp, i32, vp, vp, vp]
    L.subg_spjoin.argtypes = [vp, vp, i64, i32, vp, vp, vp, i32, vp, i64, vp, C.POINTER(i64), C.POINTER(i32), vp]
    L.subg_joiner_create.argtypes = [vp, i64, i32, vp, i32, i64, i32, i32, C.POINTER(vp)]
    L.subg_joiner_submit.argtypes = [vp, vp, i32, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i32)]
    L.subg_joiner_rows.argtypes = [vp, i32, C.POINTER(i64)]
    L.subg_joiner_free.argtypes = [vp]
    L.subg_joiner_free.restype = None
    L.subg_ppr_topk.argtypes = [vp, vp, i64, C.c_float, C.c_float, i32, i32, vp, i32, vp, C.POINTER(vp)]
    L.subg_spg_encode.argtypes = [vp, vp, i32, vp, C.POINTER(vp)]
    L.subg_spg_pushes.argtypes = [vp, C.POINTER(i64)]
    L.subg_walk_sample.argtypes = [vp, vp, i64, i32, i32, u64, i32, i32, vp, C.POINTER(vp)]
    L.subg_walkset_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_uint32)]
    L.subg_walkset_export.argtypes = [vp, vp, vp, vp, vp, vp]
    L.subg_walkset_views.argtypes = [vp] + [C.POINTER(vp)] * 4
    L.subg_walkset_free.argtypes = [vp]
    L.subg_walkset_free.restype = None
    L.subg_walk_join.argtypes = [vp, i64, i64, vp, vp, vp, i64, vp, vp, i32, vp]
    L.subg_batch_sample.argtypes = [vp, vp, i64, i32, i32, i32, C.c_uint32, vp, i64, C.POINTER(i64), vp]
    L.subg_timing_enable.argtypes = [i32]
    L.subg_timing_read.argtypes = [i32, C.POINTER(C.c_double), C.POINTER(i64)]
    L.subg_launch_count.restype = i64
    L.subg_launch_count.argtypes = []
    L.subg_trim_cache.restype = i64
    L.subg_trim_cache.argtypes = []
    L.subg_host_alloc.argtypes = [C.POINTER(vp), i64]
    L.subg_host_free.argtypes = [vp]
    L.subg_host_free.restype = None
    for name in SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("subg_abi_version", "subg_trim_cache"):
            fn.restype = C.c_int
    _lib = L
    return L


def timing_enable(on: bool) -> None:
    load().subg_timing_enable(int(on))


def timing_read(which: int):
    """-> (summed device ms, launches) of kernel class `which` since the last read."""
    ms, cnt = C.c_double(0), C.c_int64(0)
    load().subg_timing_read(which, C.byref(ms), C.byref(cnt))
    return ms.value, cnt.value


def trim_cache() -> int:
    """Release the library's cache of large device blocks (subg_trim_cache); returns the bytes released."""
    return int(load().subg_trim_cache())


def launch_count() -> int:
    return int(load().subg_launch_count())


def check(rc: int) -> None:
    """Translate a SUBG_ERR_* code into the reference's exception class."""
    if rc == 0:
        return
    msg = load().subg_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise TypeError(msg)
    if rc == -2:
        raise MemoryError(msg)
    if rc == -3:
        raise AssertionError(msg)
    if rc == -5:
        raise SubgUnsupported(msg)
    raise RuntimeError(f"libsubg_b200: {msg} (rc={rc})")
