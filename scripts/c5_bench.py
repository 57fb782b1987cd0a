#!/usr/bin/env python
"""BASELINE configs[4] on one GPU: synthetic twitter-2010 shape generated + ingested on the device, LP M=100, walk
length 2, every node a seed.  Prints one JSON line (seeds/s, kernel ms, build ms, host phases on stderr)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SUBG_QUIET", "1")
from surel_plus_b200 import DeviceGraph, SpG, _capi, gather  # noqa: E402
from surel_plus_b200.graphs import SHAPES  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    N, E_und, gseed = SHAPES["twitter"]
    M, m = 100, 2
    gen = torch.Generator(device="cuda:0")
    gen.manual_seed(gseed)
    t0 = time.perf_counter()
    src = (torch.rand(E_und, dtype=torch.float64, device="cuda:0", generator=gen).pow_(2.0) * N).to(torch.int64).clamp_(max=N - 1)
    dst = torch.randint(0, N, (E_und,), dtype=torch.int64, device="cuda:0", generator=gen)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    g = DeviceGraph.from_edges(src, dst, num_nodes=N, symmetrize=True, drop_self_loops=True, device="cuda:0")
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    del src, dst
    torch.cuda.empty_cache()
    q = torch.arange(N, dtype=torch.int32, device="cuda:0")
    SpG.sample(g, q[:2_000_000], num_walks=M, num_steps=m, seed=1, first_visit_ranks=False).close()
    _capi.timing_enable(True)
    for w in (0, 2):
        _capi.timing_read(w)
    res = []
    for i in range(steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        spg = SpG.sample(g, q, num_walks=M, num_steps=m, seed=111413 + i, first_visit_ranks=False)
        e1.record()
        torch.cuda.synchronize()
        k_ms, k_n = _capi.timing_read(0)
        b_ms, _ = _capi.timing_read(2)
        res.append({"step_ms": e0.elapsed_time(e1), "kernel_ms": k_ms, "kernel_launches": k_n, "build_ms": b_ms, "T": spg.T,
                    "c": spg.c})
        if i + 1 < steps:
            spg.close()
    # SpJoin on the resident SpG
    B = 21504
    xpe = (torch.from_numpy(spg.enc_table()).float() / M).cuda()
    edges = [torch.randint(0, N, (2, B), dtype=torch.int64, device="cuda:0") for _ in range(8)]
    for e in edges[:3]:
        gather(e, spg, "cuda:0", True, xpe)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(32):
        gather(edges[i % 8], spg, "cuda:0", True, xpe)
    e1.record()
    torch.cuda.synchronize()
    best = min(res, key=lambda r: r["step_ms"])
    print(json.dumps({"workload": f"synthetic twitter-2010 shape N={g.N} directed nnz={g.E}, LP M={M} m={m}, all nodes are seeds, 1 GPU",
                      "edge_gen_s": t1 - t0, "ingest_s": t2 - t1, "ingest_edges_per_s": 2 * E_und / (t2 - t1),
                      "seeds_per_s": N / best["step_ms"] * 1e3, "steps": res,
                      "spjoin_queries_per_s": 32 * B / e0.elapsed_time(e1) * 1e3}))


if __name__ == "__main__":
    main()
