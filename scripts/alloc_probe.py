"""2-rank probe: does the per-call allocation time depend on NCCL being initialised?  (torchrun, SUBG_PROFILE_HOST=1)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SUBG_QUIET", "1")
import torch.distributed as dist
from surel_plus_b200 import DeviceGraph, SpG
from surel_plus_b200.graphs import named_graph

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
A = named_graph("ppa", float(os.environ.get("PROBE_SCALE", "1.0")))
g = DeviceGraph.from_scipy(A, f"cuda:{local}")
q = torch.arange(A.shape[0], dtype=torch.int32, device=f"cuda:{local}")
def loop(tag):
    for i in range(6):
        t0 = time.perf_counter()
        s = SpG.sample(g, q, num_walks=200, num_steps=3, seed=i, first_visit_ranks=False)
        t1 = time.perf_counter()
        s.close()
        if os.environ.get("PROBE_SYNC", "1") == "1":
            torch.cuda.synchronize()
        print(f"[probe {tag} rank{local}] call_ms={1e3*(t1-t0):.2f}", file=sys.stderr, flush=True)
loop("nodist")
if "RANK" in os.environ:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier()
    loop("dist")
    t = torch.ones(1, device=f"cuda:{local}"); dist.all_reduce(t)
    loop("dist+coll")
    dist.destroy_process_group()
