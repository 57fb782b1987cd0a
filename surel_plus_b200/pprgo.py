"""Host mirror of sampler/pprgo.py:83-111 (`topk_ppr_matrix`) and utils.py:20-39 (`encoding`)
running on the device: ACL forward push per seed, top-k, normalisation and the PPR / SPD
structure encoders are CUDA kernels (csrc/ppr.cu); the result is a float64 value SpG that stays
in HBM and that gather / pgather of surel_plus_b200.train join in place (train.py:38-43)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .spg import DeviceGraph, SpG, _ptr, _stream

_NORMS = {"row": 0, "sym": 1, "col": 2}
_ENCODERS = {None: 0, "PPR": 1, "SPD": 2}

# value SpG -> device graph it was sampled from (lets encoding(x, adj, mode) skip a second upload)
_graph_of: dict = {}


def _device_graph(adj, device):
    if isinstance(adj, DeviceGraph):
        return adj, False
    return DeviceGraph.from_scipy(adj, device), True


def topk_ppr_matrix(adj_matrix, alpha, eps, idx, topk, normalization="row", device="cuda", encoder=None):
    """sampler/pprgo.py:83: SpG whose row i holds the (up to) `topk` largest approximate-PPR
    entries of node idx[i], normalised.  `adj_matrix` is the scipy CSR adjacency (or a
    DeviceGraph, then `normalization` uses the unweighted degree).  `encoder` fuses
    utils.encoding(x, adj, encoder) into the same call."""
    if normalization not in _NORMS:
        raise ValueError(f"Unknown PPR normalization: {normalization}")  # pprgo.py:109
    if encoder not in _ENCODERS:
        raise NotImplementedError(encoder)  # utils.py:37-38
    graph, own = _device_graph(adj_matrix, device)
    lib = _capi.load()
    q = np.ascontiguousarray(np.asarray(idx).astype(np.int32, copy=False))
    ndeg = None
    if not isinstance(adj_matrix, DeviceGraph) and normalization != "row":
        # deg = adj.sum(1) (pprgo.py:89,99).  For an unweighted adjacency (every stored entry 1 / True: what the reference's
        # loaders build) that is the row length, which the device has already; scipy's sum over 61 M entries is 0.3-0.5 s
        # on the host, half of the whole call on the citation2 shape.  Anything else is summed as the reference does.
        data = adj_matrix.data
        uniform = bool(data.all()) if data.dtype == np.bool_ else bool((data == 1).all())
        if not uniform:
            ndeg = np.ascontiguousarray(np.asarray(adj_matrix.sum(1)).ravel(), dtype=np.float64)
    h = C.c_void_p()
    _capi.check(lib.subg_ppr_topk(graph._h, _ptr(q), q.size, C.c_float(np.float32(alpha)), C.c_float(np.float32(eps)),
                                  int(topk), _NORMS[normalization], _ptr(ndeg) if ndeg is not None else None,
                                  _ENCODERS[encoder], _stream(graph.device), C.byref(h)))
    x = SpG(h, graph.device, n_nodes=graph.N)
    if own:
        graph.close()
    return x


def encoding(x, adj, encoding="DEG", device="cuda"):
    """utils.py:20-39 for a device value SpG: returns (SpG, None).  'DEG' is broken in the reference
    (it returns a sparse `agg` that gather would index as a tensor, train.py:37) and is not provided."""
    if encoding not in ("PPR", "SPD"):
        raise NotImplementedError
    if not isinstance(x, SpG):
        x = SpG.from_scipy(x, device)
    lib = _capi.load()
    graph, own = (None, False)
    if encoding == "SPD":
        graph, own = _device_graph(adj, f"cuda:{x.device}")
    h = C.c_void_p()
    _capi.check(lib.subg_spg_encode(graph._h if graph is not None else None, x._h, _ENCODERS[encoding],
                                    _stream(x.device), C.byref(h)))
    out = SpG(h, x.device, n_nodes=x.shape[1])
    if own:
        graph.close()
    return out, None
