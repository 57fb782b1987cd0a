"""surel_plus_b200 -- B200-native SubGAcc hot path (set sampling -> LP encoding -> SpG -> SpJoin).

Host side: Python over a C ABI (include/subg_b200.h); compute: hand-written CUDA for sm_100a.
No CPU fallback: importing works anywhere, calling needs the built library and a GPU.
"""
from . import _capi
from .spg import DeviceGraph, SpG
from .sampler import subg_matrix, rw_matrix
from .train import gather, hgather, bgather, pgather, JoinStream
from .subg_acc import gset_sampler, walk_sampler
from .pprgo import topk_ppr_matrix, encoding

__all__ = ["DeviceGraph", "SpG", "subg_matrix", "gather", "hgather", "bgather", "pgather", "JoinStream", "gset_sampler", "walk_sampler", "rw_matrix", "topk_ppr_matrix",
           "encoding", "_capi"]
