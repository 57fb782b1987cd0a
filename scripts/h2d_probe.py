#!/usr/bin/env python
"""Upload of a pageable ppa-sized column array (243 MB) through staged_h2d: GB/s for the thread count / chunk size in the
environment (SUBG_COPY_THREADS, SUBG_COPY_CHUNK_MB), against a plain pageable copy and a pinned copy."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surel_plus_b200.spg import staged_h2d, pinned_empty
a = np.random.default_rng(0).integers(0, 1 << 20, 60_637_674, dtype=np.int32)
dev = torch.device("cuda", 0)
def t(f, n=8):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
ms = t(lambda: staged_h2d(a, dev)) * 1e3
line = f"threads={os.environ.get('SUBG_COPY_THREADS','default')} chunk={os.environ.get('SUBG_COPY_CHUNK_MB','8')}MB staged {ms:.2f} ms {a.nbytes/ms/1e6:.1f} GB/s"
if os.environ.get("PROBE_BASE"):
    ms2 = t(lambda: torch.from_numpy(a).to(dev)) * 1e3
    p = pinned_empty(a.shape, a.dtype); np.copyto(p, a); pt = torch.from_numpy(p)
    ms3 = t(lambda: pt.to(dev, non_blocking=True)) * 1e3
    line += f" | plain pageable {ms2:.2f} ms {a.nbytes/ms2/1e6:.1f} GB/s | pinned {ms3:.2f} ms {a.nbytes/ms3/1e6:.1f} GB/s | cores {os.cpu_count()}"
print(line)
