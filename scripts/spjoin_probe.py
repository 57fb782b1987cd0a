#!/usr/bin/env python
"""Where does a SpJoin batch spend its time?  python scripts/spjoin_probe.py [workload] [B]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SUBG_QUIET", "1")
from bench import WORKLOADS, make_queries  # noqa: E402
from surel_plus_b200 import DeviceGraph, SpG, _capi, gather  # noqa: E402
from surel_plus_b200.graphs import named_graph  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "ppa"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 21504
    shape, M, m, k = WORKLOADS[wl]
    A = named_graph(shape, 1.0)
    g = DeviceGraph.from_scipy(A, "cuda:0")
    spg = SpG.sample(g, torch.arange(A.shape[0], dtype=torch.int32, device="cuda:0"), num_walks=M, num_steps=m, seed=1,
                     first_visit_ranks=False)
    xpe = (torch.from_numpy(spg.enc_table()).float() / M).cuda()
    rng = np.random.default_rng(7)
    edges = [torch.from_numpy(make_queries(A, B, k, rng)).cuda() for _ in range(8)]
    for e in edges[:4]:
        gather(e, spg, "cuda:0", True, xpe)
    _capi.timing_enable(True)
    _capi.timing_read(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 64
    for i in range(reps):
        xz, ptr = gather(edges[i % 8], spg, "cuda:0", True, xpe)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    k_ms, k_n = _capi.timing_read(1)
    _capi.timing_enable(False)
    print(f"{wl} B={B}: {dt * 1e3:.3f} ms/batch = {B / dt:.3e} q/s; join kernel {k_ms / k_n:.3f} ms; rows/batch {xz.shape[0]}")
    os.environ["SUBG_PROFILE_HOST"] = "1"
    for i in range(3):
        t0 = time.perf_counter()
        xz, ptr = gather(edges[i], spg, "cuda:0", True, xpe)
        print(f"  call {i}: {1e3 * (time.perf_counter() - t0):.3f} ms host", file=sys.stderr)


if __name__ == "__main__":
    main()
