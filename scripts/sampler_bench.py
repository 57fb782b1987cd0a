#!/usr/bin/env python
"""Sampler-only timing for sweeps on the GPU box: python scripts/sampler_bench.py ppa [steps]
Prints one line: workload, env knobs, step ms, sampler-kernel ms, build ms, avg set size."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SUBG_QUIET", "1")
from bench import WORKLOADS  # noqa: E402
from surel_plus_b200 import DeviceGraph, SpG, _capi  # noqa: E402
from surel_plus_b200.graphs import named_graph  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "ppa"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    ranks = os.environ.get("RANKS", "0") == "1"
    shape, M, m, _ = WORKLOADS[wl]
    A = named_graph(shape, 1.0)
    g = DeviceGraph.from_scipy(A, "cuda:0")
    q = torch.arange(A.shape[0], dtype=torch.int32, device="cuda:0")
    for i in range(2):
        SpG.sample(g, q, num_walks=M, num_steps=m, seed=i, first_visit_ranks=ranks).close()
    _capi.timing_enable(True)
    for w in (0, 2):
        _capi.timing_read(w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    T = 0
    for i in range(steps):
        s = SpG.sample(g, q, num_walks=M, num_steps=m, seed=10 + i, first_visit_ranks=ranks)
        T += s.T
        s.close()
    e1.record()
    torch.cuda.synchronize()
    k_ms, k_n = _capi.timing_read(0)
    b_ms, _ = _capi.timing_read(2)
    knobs = {k: v for k, v in os.environ.items() if k.startswith("SUBG_") and k != "SUBG_QUIET"}
    print(f"{wl} ranks={int(ranks)} {knobs} step_ms={e0.elapsed_time(e1) / steps:.3f} kernel_ms={k_ms / max(k_n, 1):.3f} "
          f"build_ms={b_ms / steps:.3f} avg_set={T / steps / A.shape[0]:.1f} seeds_per_s={A.shape[0] * steps / e0.elapsed_time(e1) * 1e3:.3e}",
          flush=True)


if __name__ == "__main__":
    main()
