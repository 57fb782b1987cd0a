#!/usr/bin/env python
"""SUREL-v1 walk_sampler + walk_join throughput on the collab-shape graph (legacy API)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SUBG_QUIET", "1")
from surel_plus_b200 import subg_acc  # noqa: E402
from surel_plus_b200.graphs import named_graph  # noqa: E402


def main():
    A = named_graph("collab")
    n, M, m = 100_000, 100, 3
    q = np.random.default_rng(0).permutation(A.shape[0])[:n].astype(np.int32)
    indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    subg_acc.walk_sampler(indptr, indices, q[:2000], num_walks=M, num_steps=m, replacement=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    walks, obj = subg_acc.walk_sampler(indptr, indices, q, num_walks=M, num_steps=m, replacement=True)
    t1 = time.perf_counter()
    qq = q[np.random.default_rng(1).integers(0, n, (4096, 2))]
    keys = list(obj[:, 0])
    t2 = time.perf_counter()
    out = subg_acc.walk_join(walks, keys, qq)
    t3 = time.perf_counter()
    res = {"workload": f"collab shape, walk_sampler n={n} M={M} m={m} replacement=True; walk_join Q=4096",
           "walk_sampler_seeds_per_s": n / (t1 - t0), "walk_sampler_s": t1 - t0, "sets_total": int(sum(len(k) for k in keys)),
           "walk_join_queries_per_s": 4096 / (t3 - t2), "walk_join_s": t3 - t2, "walk_join_out_shape": list(out.shape)}
    ref = None
    try:
        from oracle import reference as R
        ref = R.subg_acc()
    except Exception:
        pass
    if ref is not None:  # the compiled reference on the host cores, bounded sample
        qs = q[:20000]
        t0 = time.perf_counter()
        w2, o2 = ref.walk_sampler(indptr, indices, qs, num_walks=M, num_steps=m, replacement=True, nthread=os.cpu_count())
        t1 = time.perf_counter()
        ref.walk_join(w2, list(o2[:, 0]), qs[np.random.default_rng(1).integers(0, len(qs), (1024, 2))], nthread=os.cpu_count())
        t2 = time.perf_counter()
        res["reference_cpu"] = {"cores": os.cpu_count(), "walk_sampler_seeds_per_s": len(qs) / (t1 - t0),
                                "walk_join_queries_per_s": 1024 / (t2 - t1)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
