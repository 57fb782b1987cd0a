#!/usr/bin/env python
"""PPR set sampler on BASELINE.json configs[2] (citation2 shape, topk=100, alpha=0.1, eps=1e-4, 'sym'):
seeds/s of the device forward push + top-k + normalisation, pushes/s, MRR-style 1-vs-1000 SpJoin queries/s on
the value SpG, and the oracle's C port of the numba kernel on the host cores on a bounded sample.
    python scripts/ppr_bench.py [n_seeds] [graph]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SUBG_QUIET", "1")
from surel_plus_b200 import DeviceGraph, _capi, gather  # noqa: E402
from surel_plus_b200.graphs import named_graph  # noqa: E402
from surel_plus_b200.pprgo import topk_ppr_matrix, encoding  # noqa: E402


def main():
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    shape = sys.argv[2] if len(sys.argv) > 2 else "citation2"
    A = named_graph(shape)
    N = A.shape[0]
    alpha, eps, topk = 0.1, 1e-4, 100
    g = DeviceGraph.from_scipy(A, "cuda:0")
    idx = np.arange(N if n_seeds <= 0 else min(n_seeds, N), dtype=np.int32)
    topk_ppr_matrix(g, alpha, eps, idx[: max(len(idx) // 50, 1)], topk, "sym").close()  # warm-up
    _capi.timing_enable(True)
    _capi.timing_read(_capi.TIMING_PPR)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x = topk_ppr_matrix(g, alpha, eps, idx, topk, "sym")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    k_ms, k_n = _capi.timing_read(_capi.TIMING_PPR)
    out = {"workload": f"synthetic ogbl-{shape}-shape (N={N}, nnz={A.nnz}), PPR alpha={alpha} eps={eps} topk={topk} sym",
           "seeds": int(len(idx)), "seeds_per_s": len(idx) / dt, "wall_ms": dt * 1e3, "push_kernel_ms": k_ms,
           "pushes": x.pushes, "pushes_per_s": x.pushes / max(k_ms, 1e-9) * 1e3, "nnz": x.T, "status": x.status}
    # MRR-style SpJoin on the PPR-encoded value SpG: each source against 1 positive + 1000 negatives
    if len(idx) == N:
        xe, _ = encoding(x, g, "PPR")
        rng = np.random.default_rng(0)
        Q = 64
        src = np.repeat(rng.integers(0, N, Q), 1001)
        dst = rng.integers(0, N, Q * 1001)
        edge = torch.from_numpy(np.stack([src, dst])).cuda()
        for _ in range(2):
            gather(edge, xe, "cuda:0", True, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 10
        for _ in range(reps):
            xz, ptr = gather(edge, xe, "cuda:0", True, None)
        e1.record()
        torch.cuda.synchronize()
        out["spjoin_queries_per_s"] = edge.shape[1] * reps / e0.elapsed_time(e1) * 1e3
        out["spjoin_rows_per_query"] = xz.shape[0] / edge.shape[1]
    # CPU: the oracle's C port of _calc_ppr_node + top-k (pprgo.py:9-62) on all host cores, bounded sample
    try:
        from oracle import pyoracle as po
        cores = os.cpu_count() or 1
        samp = np.random.default_rng(1).choice(N, min(N, 20000), replace=False).astype(np.int32)
        t0 = time.perf_counter()
        po.topk_ppr_matrix(A, alpha, eps, samp, topk, "sym", nthread=cores)
        dtc = time.perf_counter() - t0
        out["cpu_port"] = {"seeds_per_s": len(samp) / dtc, "cores": cores, "sample": f"{len(samp)} random seeds, {dtc:.1f}s"}
    except Exception as ex:  # pragma: no cover
        out["cpu_port"] = {"error": repr(ex)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
