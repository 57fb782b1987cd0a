#!/usr/bin/env python
"""bench.py -- SubGAcc hot path on B200: sampled node-sets/s (+ SpJoin queries/s).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)
    python bench.py --workload collab|ppa|dblp|citation2-ppr|twitter

Default workload (BASELINE.json configs[1]): synthetic ogbl-ppa-shape graph (576 289 nodes, 30.3 M undirected
edges), LP encoder, CLI num_steps=4 (walk length m=3), num_walks=200; every node is a seed.  One *step* = one pass
of sample -> LP-encode -> SpG build over ALL seeds, graph resident in HBM, the joinable SpG left resident in HBM.

  N = 1   `value` = seeds/s of that pass on one GPU.
  N > 1   `value` = seeds/s of the SAME pass sharded over the N GPUs (strong scaling, BASELINE configs[1] "seeds
          sharded over 8 GPUs"): contiguous seed ranges per rank, the SpG shards exchanged over NVLink (csrc/xchg.cu),
          every rank ends with the full joinable SpG; checked against the single-GPU SpG (`sharded.parity_ok`).
          The N-independent-replicas number of round 1 is kept under "replicas".
`e2e` = the reference-facing call with HOST buffers: subg_matrix(G scipy CSR, idx) -> (joinable SpG, LP table), the unit
of work the reference arm times (gset_sampler + scipy CSR build); `e2e_numpy` = gset_sampler numpy in / numpy out.
SpJoin is timed afterwards on the resident SpG at the reference's batch sizes.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("SUBG_QUIET", "1")

WORKLOADS = {
    # name: graph shape, sampler, join pattern of the reference's driver for that dataset
    "ppa": dict(shape="ppa", kind="lp", M=200, m=3, k=20, join="pair", batches=(1024, 21504)),          # main.py:32,103-106
    "collab": dict(shape="collab", kind="lp", M=200, m=2, k=10, join="pair", batches=(1024, 11264)),    # main.py:99-102
    "dblp": dict(shape="dblp", kind="lp", M=100, m=2, k=10, join="triplet", batches=(2048,)),           # main_horder.py:33
    "citation2-ppr": dict(shape="citation2", kind="ppr", topk=100, alpha=0.1, eps=1e-4, k=1000, join="pair",
                          batches=(1001, 64064)),                                                       # main.py:107-111,181
    "twitter": dict(shape="twitter", kind="lp", M=100, m=2, k=10, join="pair", batches=(1024, 21504), device_graph=True),
}
NVLINK_PEER_GBS = 770.0  # measured peer copy per direction per GPU on this pool (B200_PROFILING.md)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------ helpers
class ClockSampler:
    """SM clock / throttle reasons sampled through NVML every 5 ms while the timed region runs
    (the B200_PROFILING.md clocks line; nvidia-smi itself takes longer to start than a run lasts)."""

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.th, self.stop_flag = gpu_index, [], None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
        except Exception:
            self.nv = None

    def start(self):
        if self.nv is None or os.environ.get("BENCH_NO_CLOCKS"):
            return
        self.th = threading.Thread(target=self._pump, daemon=True)
        self.th.start()

    def _pump(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self) -> dict:
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=2)
        nv = self.nv
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        sm = [float(r[0]) for r in self.rows]
        reasons = sorted(k for k, bit in names.items() if any(r[1] & bit for r in self.rows))
        hi = sorted(sm)[len(sm) // 2:] if sm else []  # samples under load = the upper half
        return {"sm_mhz": float(np.median(hi)) if hi else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


def make_host_graph(W, scale: float):
    from surel_plus_b200.graphs import named_graph
    t0 = time.time()
    A = named_graph(W["shape"], scale)
    log(f"[bench] graph {W['shape']} x{scale}: N={A.shape[0]} nnz={A.nnz} ({time.time() - t0:.1f}s)")
    return A


def make_device_graph(W, scale, dev):
    """Edge list drawn and ingested on the device (the twitter shape: 1.47 G directed entries; a host CSR of that size
    takes minutes to build).  Same law as graphs.synthetic_graph (src = floor(N U^2), dst uniform)."""
    import torch
    from surel_plus_b200 import DeviceGraph
    from surel_plus_b200.graphs import SHAPES
    N, E_und, gseed = SHAPES[W["shape"]]
    N, E_und = max(int(N * scale), 16), max(int(E_und * scale), 16)
    gen = torch.Generator(device=dev)
    gen.manual_seed(gseed)
    t0 = time.time()
    src = (torch.rand(E_und, dtype=torch.float64, device=dev, generator=gen).pow_(2.0) * N).to(torch.int64).clamp_(max=N - 1)
    dst = torch.randint(0, N, (E_und,), dtype=torch.int64, device=dev, generator=gen)
    g = DeviceGraph.from_edges(src, dst, num_nodes=N, symmetrize=True, drop_self_loops=True, device=dev)
    del src, dst
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    log(f"[bench] graph {W['shape']} x{scale} built on the device: N={g.N} nnz={g.E} ({time.time() - t0:.1f}s)")
    return g


def seed_algorithmic_bytes(deg: np.ndarray, M: int, m: int, T: float) -> float:
    """B_seed = 24 + 4*min(d,M) + 12*M*(m-1) + 8*|S_u| (d > 0), 32 for d = 0."""
    live = deg > 0
    return float(24 * live.sum() + 4 * np.minimum(deg[live], M).sum() + 12 * M * (m - 1) * live.sum()
                 + 32 * (~live).sum() + 8 * (T - (~live).sum()))


def make_queries(deg, indptr, indices, N, B: int, k: int, rng) -> np.ndarray:
    """1 positive : k negatives per batch (the reference's training / evaluation pattern, main.py --k)."""
    npos = max(B // (k + 1), 1)
    if indptr is not None:
        rows = rng.integers(0, N, npos * 4)
        rows = rows[deg[rows] > 0][:npos]
        pos_v = indices[indptr[rows] + (rng.integers(0, 1 << 30, len(rows)) % deg[rows])]
        pos = np.stack([rows, pos_v])
    else:
        pos = rng.integers(0, N, (2, npos))
    if k >= 100:   # MRR style (citation2): every source against its 1 positive + k negatives
        src = np.repeat(pos[0], k)[: B - pos.shape[1]]
        neg = np.stack([src, rng.integers(0, N, len(src))])
    else:
        neg = rng.integers(0, N, (2, B - pos.shape[1]))
    e = np.concatenate([pos, neg], axis=1).astype(np.int64)[:, :B]
    return e[:, rng.permutation(e.shape[1])]


def make_triplets(deg, indptr, indices, N, B: int, rng) -> np.ndarray:
    """(u, v, w): (u, v) an existing edge, w uniform (dataloader.py:272-276)."""
    u = rng.integers(0, N, B * 3)
    u = u[deg[u] > 0][:B]
    v = indices[indptr[u] + (rng.integers(0, 1 << 30, len(u)) % deg[u])]
    return np.stack([u, v, rng.integers(0, N, len(u))]).astype(np.int64)


def workload_config(args, W, N, nnz):
    if W["kind"] == "lp":
        what = (f"LP encoder, num_walks={W['M']}, walk length m={W['m']} (CLI num_steps={W['m'] + 1}), all nodes are seeds, "
                f"neg ratio k={W['k']}")
    else:
        what = f"PPR sampler topk={W['topk']} alpha={W['alpha']} eps={W['eps']} 'sym', all nodes are seeds, 1-vs-{W['k']} queries"
    return {"workload": f"synthetic {W['shape']}-shape graph x{args.scale:g} (N={N}, directed nnz={nnz}), {what}",
            "name": args.workload, "rng": "philox4x32-10", "l2": "graph + SpG output >> 126 MB L2; no flush needed",
            "spjoin": f"{W['join']} queries, batches {list(W['batches'])}"}


# ------------------------------------------------------------------------------ reference arm
def reference_sampler(W, A, target_s: float, rng, nthread: int):
    """The UNMODIFIED reference (oracle/_ref, else the oracle port) on a bounded random sample of seeds sized for about
    `target_s` seconds.  Returns (seeds, run, kind, what) with run(q) -> (seconds for gset_sampler alone, seconds incl.
    the subg_matrix CSR build) -- or the PPR pipeline for the PPR workload."""
    from oracle import reference as ref, pyoracle as po
    n = A.shape[0]
    if W["kind"] == "ppr":
        def run(q):
            t0 = time.perf_counter()
            po.topk_ppr_matrix(A, W["alpha"], W["eps"], q, W["topk"], "sym", nthread=nthread)
            t = time.perf_counter() - t0
            return t, t
        kind, what = "port", f"oracle C port of pprgo.py:9-111 (numba prange -> {nthread} threads)"
    else:
        subg = ref.subg_acc()
        M, m = W["M"], W["m"]
        indptr = A.indptr.astype(np.int32)
        indices = A.indices.astype(np.int32)

        def run(q):
            t0 = time.perf_counter()
            if subg is not None:
                nsize, remap, enc = subg.gset_sampler(indptr, indices, q, num_walks=M, num_steps=m, nthread=nthread)
            else:
                nsize, remap, enc = po.gset_sampler_replay(indptr, indices, q, M, m)
            t1 = time.perf_counter()
            po.subg_matrix_from(nsize, remap, enc, q, n, m + 1)  # random_walks.py:79-81
            return t1 - t0, time.perf_counter() - t0
        kind = "reference" if subg is not None else "port"
        what = f"gset_sampler(nthread={nthread}) + scipy CSR build (= subg_matrix, random_walks.py:74-82)"
    perm = rng.permutation(n).astype(np.int32)
    probe = perm[: min(n, 4000 if W["kind"] == "lp" else 1000)]
    t = run(probe)[1]
    S = int(min(n, max(len(probe), len(probe) * target_s / max(t, 1e-3))))
    return perm[:S], run, kind, what


class _Quiet:
    """fd-level stdout redirect: the reference prints '#SubGAcc' lines from C."""

    def __enter__(self):
        self.devnull, self.saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
        sys.stdout.flush()
        os.dup2(self.devnull, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.devnull)
        os.close(self.saved)


def host_graph_for_reference(args, W):
    if not W.get("device_graph"):
        return make_host_graph(W, args.scale)
    # the twitter shape is drawn on the device (a host build takes minutes); the reference reads the exported CSR
    import scipy.sparse as sp
    import torch
    g = make_device_graph(W, args.scale, torch.device("cuda", 0))
    indptr, indices = g.csr()
    g.close()
    return sp.csr_matrix((np.ones(len(indices), dtype=bool), indices, indptr), shape=(len(indptr) - 1,) * 2)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    W = WORKLOADS[args.workload]
    A = host_graph_for_reference(args, W)
    rng = np.random.default_rng(0)
    cores = os.cpu_count() or 1
    with _Quiet():
        q, run, kind, what = reference_sampler(W, A, args.ref_seconds, rng, cores)
        for _ in range(args.warmup):
            run(q[: max(len(q) // 8, 1)])
        t_s = t_all = 0.0
        for _ in range(args.steps):
            a, b = run(q)
            t_s += a
            t_all += b
    val = len(q) * args.steps / t_all
    sample = f"{len(q)} random seeds of {A.shape[0]} per step; {what}"
    line = {
        "impl": "reference", "metric": "sampled_node_sets_per_sec", "value": val, "unit": "seeds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_all / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32" if W["kind"] == "lp" else "f32", "data": "synthetic",
        "config": dict(workload_config(args, W, A.shape[0], A.nnz), reference_sample=sample),
        "cpu_baseline": {"value": val, "unit": "seeds/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "seeds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "api": "subg_matrix: sampler + CSR build (host arrays in, joinable scipy SpG out)"},
        "e2e_numpy": {"value": len(q) * args.steps / t_s, "unit": "seeds/s", "api": "gset_sampler alone (numpy in / numpy out)"},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------ our arm
class Ctx:
    pass


def run_ours(args):
    import torch
    import torch.distributed as dist

    c = Ctx()
    c.args, c.torch, c.dist = args, torch, dist
    c.rank = int(os.environ.get("RANK", "0"))
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    if c.world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", c.local))
    torch.cuda.set_device(c.local)
    c.dev = torch.device("cuda", c.local)
    c.peaks = load_peaks()
    W = c.W = WORKLOADS[args.workload]

    def barrier():
        if c.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if c.world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=c.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(flag: bool) -> bool:
        if c.world == 1:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=c.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t.item()))
    c.barrier, c.max_over_ranks, c.all_ok = barrier, max_over_ranks, all_ok

    from surel_plus_b200 import DeviceGraph
    if W.get("device_graph"):
        c.A = None
        c.graph = make_device_graph(W, args.scale, c.dev)
        indptr_d, _ = c.graph.csr(device=c.dev)
        c.deg = (indptr_d[1:] - indptr_d[:-1]).cpu().numpy()
        del indptr_d
        torch.cuda.empty_cache()
        c.indptr = c.indices = None
    else:
        c.A = make_host_graph(W, args.scale)
        c.graph = DeviceGraph.from_scipy(c.A, c.dev)
        c.deg = np.diff(c.A.indptr)
        c.indptr, c.indices = c.A.indptr, c.A.indices
    c.n = c.graph.N
    c.nnz = c.graph.E
    c.query = np.arange(c.n, dtype=np.int32)
    c.q_dev = torch.arange(c.n, dtype=torch.int32, device=c.dev)
    c.clocks = ClockSampler(c.local)

    if W["kind"] == "ppr":
        head, spg = bench_ppr(c)
    elif c.world == 1:
        head, spg = bench_single(c)
    else:
        head, spg = bench_sharded(c)

    # ---- e2e: the reference-facing calls with host buffers -----------------------------------------
    e2e = e2e_numpy = None
    if args.quick:
        e2e = {"value": None, "why": "--quick"}
    elif c.A is not None:
        try:
            e2e, e2e_numpy = bench_e2e(c)
        except Exception as ex:  # noqa: BLE001
            e2e = {"error": repr(ex)}
            log(f"[bench] e2e block failed on rank {c.rank}: {ex!r}")
    else:
        e2e = {"value": None, "unit": "seeds/s", "why": "the graph of this workload is drawn and ingested on the device "
               "(1.47 G entries); there is no host CSR to pass through the reference-facing call"}

    # ---- SpJoin on the resident SpG ----------------------------------------------------------------
    spjoin = []
    try:
        if args.quick:
            raise RuntimeError("--quick: SpJoin not timed")
        if W["kind"] == "ppr" and c.world > 1:
            raise RuntimeError("PPR rows are partitioned over the ranks, not replicated: SpJoin is timed at N = 1")
        for B in (args.spjoin_batch or W["batches"]):
            spjoin.append(bench_spjoin(c, spg, int(B)))
    except Exception as ex:  # noqa: BLE001 -- the seeds/s headline must survive a failure of the secondary block
        spjoin.append({"error": repr(ex)})
        log(f"[bench] SpJoin block failed on rank {c.rank}: {ex!r}")

    # ---- the no-replication alternative: linked shards (remote rows read over NVLink at join time) ------------------
    linked = None
    if c.world > 1 and W["kind"] == "lp" and not args.quick and args.linked:
        try:
            linked = bench_linked(c)
        except Exception as ex:  # noqa: BLE001
            linked = {"error": repr(ex)}
            log(f"[bench] linked block failed on rank {c.rank}: {ex!r}")

    cpu = None
    if c.rank == 0 and c.world == 1 and not args.no_cpu_baseline and not args.quick and c.A is not None:
        cpu = cpu_baseline(c, spg, spjoin)
    if spg is not None:
        spg.close()

    if c.rank == 0:
        line = {
            "metric": "sampled_node_sets_per_sec", "unit": "seeds/s", "n_gpus": c.world, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
            "dtype": "int32" if W["kind"] == "lp" else "f32", "data": "synthetic",
            "config": workload_config(args, W, c.n, c.nnz),
        }
        line.update(head)
        line["e2e"] = e2e
        if e2e_numpy is not None:
            line["e2e_numpy"] = e2e_numpy
        line["spjoin"] = spjoin[-1] if spjoin else None      # the large batch; all sizes under spjoin_batches
        line["spjoin_batches"] = spjoin
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if linked is not None:
            line["linked"] = linked
        print(json.dumps(line), flush=True)
    if c.world > 1:
        from surel_plus_b200.parallel import close_exchanges
        barrier()
        close_exchanges()
        dist.destroy_process_group()
    return 0


def sampler_roofline(c, deg, M, m, T_avg, n_seeds, k_ms, k_launches, b_ms, steps, ms_total, lp_rows):
    alg_bytes = seed_algorithmic_bytes(deg, M, m, T_avg)
    k_avg_ms = k_ms / max(k_launches, 1)
    achieved = alg_bytes / (k_avg_ms / 1e3) / 1e9
    # what explains `frac`: the walk is random 4-byte gathers, so the kernel is compared with the measured
    # rate of that access pattern (profiles/r1_gather_micro.txt: plain random gathers over an array of the CSR's size, one B200)
    live = int((deg > 0).sum())
    visits = float(live) * M * m + n_seeds
    csr_mb = (4.0 * float(c.graph.E) + 8.0 * float(c.graph.N)) / 2 ** 20
    ceiling = 285e9 if csr_mb <= 100 else (78e9 if csr_mb <= 400 else 42e9)
    later = float(live) * M * (m - 1)   # hops 2..m: row info of a random node, then a uniformly random entry of its row
    gather = {"visits_per_s": visits / (k_avg_ms / 1e3), "random_column_gathers_per_s": later / (k_avg_ms / 1e3),
              "csr_MB": csr_mb, "random_gather_rate_measured": ceiling,
              "frac_of_gather_rate": later / (k_avg_ms / 1e3) / ceiling,
              "what": "gathers of the later hops only (the first hop reads the seed's own row) over the rate of plain random "
                      "4-byte gathers from an array of the CSR's size with nothing else in the kernel",
              "regime": "instruction issue (CSR inside the L2)" if csr_mb <= 100 else "random DRAM sector gathers",
              "source": "profiles/r1_gather_micro.txt (L2-resident 285 G/s, 243 MB 78 G/s, 1 GB 42 G/s)"}
    return {"bound": "hbm", "kernel": "gset_sample_kernel", "achieved": achieved, "peak": c.peaks["hbm_gbs"],
            "unit": "GB/s", "frac": achieved / c.peaks["hbm_gbs"], "traffic": load_traffic("gset_sample", c.args),
            "peak_source": c.peaks["source"], "algorithmic_bytes_per_launch": alg_bytes,
            "kernel_ms_per_launch": k_avg_ms, "kernel_share_of_step": k_ms / ms_total,
            "spg_build_ms_per_step": b_ms / steps, "avg_set_size": T_avg / max(n_seeds, 1), "unique_lp_rows": int(lp_rows),
            "gather": gather}


def bench_single(c, with_clocks=True):
    """One GPU: sample + encode + SpG build over all seeds, graph and SpG resident."""
    torch, args, W = c.torch, c.args, c.W
    from surel_plus_b200 import SpG, _capi
    M, m = W["M"], W["m"]
    base_seed = 111413 + 1000003 * c.rank

    def step(i):  # what subg_matrix builds: the SpG (sorted CSR-of-sets + LP table), resident and joinable
        return SpG.sample(c.graph, c.q_dev, num_walks=M, num_steps=m, seed=base_seed + i, rng_mode=_capi.SUBG_RNG_PHILOX,
                          first_visit_ranks=False)

    for i in range(max(args.warmup, 3)):
        step(i).close()
    _capi.timing_enable(True)
    for w in (0, 1, 2):
        _capi.timing_read(w)
    if c.rank == 0 and with_clocks:
        c.clocks.start()
    launches0 = _capi.launch_count()
    c.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    T_sum, spg = 0, None
    for i in range(args.steps):
        if spg is not None:
            spg.close()
        spg = step(args.warmup + i)
        T_sum += spg.T
    ev1.record()
    c.barrier()
    launches = _capi.launch_count() - launches0
    ms_total = c.max_over_ranks(ev0.elapsed_time(ev1))
    k_ms, k_launches = _capi.timing_read(_capi.TIMING_SAMPLER)
    b_ms, _ = _capi.timing_read(_capi.TIMING_BUILD)
    _capi.timing_enable(False)
    clk = c.clocks.stop() if c.rank == 0 and with_clocks else None
    T_avg = T_sum / args.steps
    roof = sampler_roofline(c, c.deg, M, m, T_avg, c.n, k_ms, k_launches, b_ms, args.steps, ms_total, spg.c)
    # total work (one pass over all seeds) is the same at every N: the N = 1 point of the strong-scaling series
    head = {"value": c.world * c.n * args.steps / (ms_total / 1e3), "ms_per_step": ms_total / args.steps, "scaling": "strong",
            "clocks": clk, "gpu_launches": int(launches), "roofline": roof}
    return head, spg


def bench_sharded(c):
    """N GPUs, ONE pass over all seeds: seed ranges per rank + exchange over NVLink; every rank holds the full SpG."""
    torch, W = c.torch, c.W
    from surel_plus_b200 import SpG, _capi
    from surel_plus_b200.parallel import partition_by_work, sharded_sample
    M, m = W["M"], W["m"]
    # contiguous seed ranges balanced by a degree-based estimate of the set size (the synthetic generator puts the hubs
    # at the low ids; equal-count ranges would leave rank 0 with the largest sets)
    bounds = partition_by_work(0.5 * M * m + 1.5 * np.minimum(c.deg, M), c.world)
    lo, hi = int(bounds[c.rank]), int(bounds[c.rank + 1])

    def timed(mode, steps, warm):
        for i in range(warm):
            sharded_sample(c.graph, c.q_dev, num_walks=M, num_steps=m, seed=7 + i, bounds=bounds, mode=mode).close()
        _capi.timing_enable(True)
        for w in (0, 2, 4):
            _capi.timing_read(w)
        launches0 = _capi.launch_count()
        c.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        spg, x_ev, T_sum = None, [], 0
        for i in range(steps):
            if spg is not None:
                spg.close()
            spg = sharded_sample(c.graph, c.q_dev, num_walks=M, num_steps=m, seed=111413 + i, bounds=bounds, mode=mode)
            x_ev.append(spg.exchange_events)
            T_sum += spg.T
        s1.record()
        c.barrier()
        ms = c.max_over_ranks(s0.elapsed_time(s1))
        launches = _capi.launch_count() - launches0
        k_ms, k_n = _capi.timing_read(_capi.TIMING_SAMPLER)
        b_ms, _ = _capi.timing_read(_capi.TIMING_BUILD)
        p_ms, p_n = _capi.timing_read(_capi.TIMING_EXCHANGE)
        _capi.timing_enable(False)
        x_ms = c.max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in x_ev])))
        pull_ms = c.max_over_ranks(p_ms / max(p_n, 1))
        if os.environ.get("BENCH_PER_RANK"):
            log(f"[bench] rank {c.rank} mode {mode}: pull {p_ms / max(p_n, 1):.2f} ms, sampler kernel {k_ms / max(k_n, 1):.2f} ms, "
                f"exchange {float(np.mean([a.elapsed_time(b) for a, b in x_ev])):.2f} ms")
        recv = float(spg.exchange_received)
        info = {"mode": spg.exchange_mode, "ms_per_pass": ms / steps, "seeds_per_s": c.n / (ms / steps / 1e3),
                "sampler_kernel_ms": c.max_over_ranks(k_ms / max(k_n, 1)),
                "exchange_ms": x_ms, "what_exchange_ms_covers": "pack + header all-gather + LP-table merge + pull/unpack + end barrier",
                "wire_bytes_per_entry": spg.exchange_entry_bytes, "packed_bytes_total": float(spg.exchange_bytes),
                "received_bytes_per_gpu": recv, "pull_kernel_ms": pull_ms,
                "exchange_GBps_per_gpu": recv / (x_ms / 1e3) / 1e9,
                "pull_GBps_per_gpu": recv / (pull_ms / 1e3) / 1e9 if pull_ms > 0 else None}
        return spg, info, ms, launches, (k_ms, k_n, b_ms, T_sum / steps)

    want = (os.environ.get("SUBG_EXCHANGE") or "peer").lower()
    if c.rank == 0:
        c.clocks.start()
    spg, info, ms_total, launches, (k_ms, k_n, b_ms, T_avg) = timed(want, c.args.steps, max(c.args.warmup, 3))
    clk = c.clocks.stop() if c.rank == 0 else None
    # ---- the sharded result must BE the single-GPU SpG (indices are global; LP ids in global first-occurrence order)
    parity_err = None
    try:
        torch.cuda.synchronize()
        _capi.trim_cache()           # the check holds a second full SpG: give the cached blocks back first
        torch.cuda.empty_cache()
        ref = SpG.sample(c.graph, c.q_dev, num_walks=M, num_steps=m, seed=111413 + c.args.steps - 1, rng_mode=_capi.SUBG_RNG_PHILOX,
                         first_visit_ranks=False)
        same = (ref.n, ref.T, ref.c, ref.max_set) == (spg.n, spg.T, spg.c, spg.max_set) and np.array_equal(ref.enc_table(), spg.enc_table())
        if same:
            # rows compared in place (no compaction of the GB-sized arrays): a checksum per row over its (node, LP id) pairs
            same = bool(torch.equal(row_checksums(torch, ref), row_checksums(torch, spg)))
        ref.close()
    except Exception as ex:  # noqa: BLE001 -- e.g. no room for a second full SpG beside the exchanged one
        same, parity_err = False, repr(ex)
        log(f"[bench] parity check failed to run on rank {c.rank}: {ex!r}")
    parity_ok = c.all_ok(same)
    torch.cuda.empty_cache()
    other = None
    if info["mode"] == "peer" and not c.args.no_exchange_compare:
        spg.close()
        spg, other, _, _, _ = timed("nccl", max(2, min(c.args.steps, 3)), 2)
    deg_r = c.deg[lo:hi]
    roof = sampler_roofline(c, deg_r, M, m, T_avg * (hi - lo) / c.n, hi - lo, k_ms, k_n, b_ms, c.args.steps, ms_total, spg.c)
    roof["note"] = "rank 0's seed range (1/N of the seeds per launch)"
    pull = info["pull_GBps_per_gpu"] or 0.0
    info.update({"value": info["seeds_per_s"], "unit": "seeds/s", "scaling": "strong", "parity_ok": parity_ok, "parity_error": parity_err,
                 "parity_check": "every rank: the exchanged SpG equals SpG.sample of all seeds on one GPU (sizes, LP table, "
                                 "per-row checksums of the (node, LP id) pairs)",
                 "what": "seed ranges sampled per rank, shards packed and pulled by the peers over NVLink; every rank holds "
                         "the full joinable SpG",
                 "nvlink": {"bound": "nvlink", "achieved": pull, "peak": NVLINK_PEER_GBS, "unit": "GB/s",
                            "frac": pull / NVLINK_PEER_GBS, "peak_source": "B200_PROFILING.md measured peer copy per direction"}})
    if other is not None:
        info["nccl_staged"] = other
    # secondary: N independent replicas (what round 1 reported as the headline)
    rep = None
    if not c.args.no_replicas:
        spg.close()
        rep_head, spg = bench_single(c, with_clocks=False)
        rep = {"value": rep_head["value"], "ms_per_step": rep_head["ms_per_step"], "scaling": "weak",
               "what": "every rank runs its own pass over all seeds on a full replica; no collective (round-1 headline)"}
    head = {"value": info["value"], "ms_per_step": ms_total / c.args.steps, "scaling": "strong", "clocks": clk,
            "gpu_launches": int(launches), "roofline": roof, "sharded": info}
    if rep is not None:
        head["replicas"] = rep
    return head, spg


def row_checksums(torch, spg):
    """int64 [n]: order-independent checksum of every row's (node, LP id) pairs, computed on the layout SpJoin reads
    (rows are visited in chunks of about 64 M entries: the twitter shape has 5.4 G of them)."""
    r = spg.rows()
    n = spg.n
    dev = r["nsize"].device
    sizes = r["nsize"].long()
    ends = torch.cumsum(sizes, 0)
    off = ends - sizes
    total = int(ends[-1]) if n else 0
    out = torch.zeros(n, dtype=torch.int64, device=dev)
    step = 1 << 26
    cuts = torch.searchsorted(ends, torch.arange(step, max(total, 1), step, device=dev)) + 1 if total > step else ends[:0]
    bounds = [0] + sorted(set(int(x) for x in cuts.tolist())) + [n]
    for a, b in zip(bounds[:-1], bounds[1:]):
        if a >= b:
            continue
        sz = sizes[a:b]
        tot = int(sz.sum())
        row_of = torch.repeat_interleave(torch.arange(a, b, device=dev), sz, output_size=tot)
        pos = torch.arange(tot, device=dev) - (off[row_of] - off[a]) + r["rowbeg"][row_of]
        v = r["indices"][pos].long() * 1000003 + r["data"][pos].long() * 7919 + 1
        out.index_add_(0, row_of, v * v)
    return out


def bench_linked(c):
    """N > 1: one pass with LINKED shards (parallel.linked_sample: no bulk exchange, every rank keeps its rows; the LP
    tables are merged and 12 bytes of row metadata per seed cross NVLink), and SpJoin on the linked SpG, whose remote rows
    come over NVLink inside the join kernel.  Reported beside the replicated pass: the trade is pass time against join rate."""
    torch, args, W = c.torch, c.args, c.W
    from surel_plus_b200.parallel import linked_exchange, linked_sample, partition_by_work
    M, m = W["M"], W["m"]
    bounds = partition_by_work(0.5 * M * m + 1.5 * np.minimum(c.deg, M), c.world)
    xc = linked_exchange(c.graph, int(np.max(np.diff(bounds))), M, m)   # one slab, re-staged by every pass
    for i in range(2):
        lk = linked_sample(c.graph, c.q_dev, num_walks=M, num_steps=m, seed=7 + i, bounds=bounds, exchange=xc)
        c.barrier()
        lk.close()
        c.barrier()
    steps = max(2, min(args.steps, 5))
    c.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lk = None
    for i in range(steps):
        if lk is not None:
            lk.close()
            c.barrier()
        lk = linked_sample(c.graph, c.q_dev, num_walks=M, num_steps=m, seed=111413 + i, bounds=bounds, exchange=xc)
    e1.record()
    c.barrier()
    ms = c.max_over_ranks(e0.elapsed_time(e1)) / steps
    out = {"ms_per_pass": ms, "seeds_per_s": c.n / (ms / 1e3), "unit": "seeds/s",
           "what": "seed ranges sampled per rank, shards staged unpacked in the IPC-mapped slabs and linked (LP tables merged, ids "
                   "relabelled in place, row offsets / sizes fetched); rows are NOT copied: SpJoin reads remote rows over NVLink",
           "spjoin": []}
    for B in (args.spjoin_batch or W["batches"]):
        try:
            b = bench_spjoin(c, lk, int(B))
            out["spjoin"].append({k: b.get(k) for k in ("batch", "pattern", "value", "unit", "ms_per_batch")}
                                 | {"stream": (b.get("stream") or {}).get("value"),
                                    "kernel_ms": (b.get("roofline") or {}).get("kernel_ms_per_launch")})
        except Exception as ex:  # noqa: BLE001
            out["spjoin"].append({"batch": int(B), "error": repr(ex)})
    c.barrier()
    lk.close()
    torch.cuda.synchronize()
    c.barrier()
    xc.close()
    return out


def bench_ppr(c):
    """configs[2]: PPR top-k set sampler (forward push + top-k + 'sym' normalisation + PPR encoder)."""
    torch, args, W = c.torch, c.args, c.W
    from surel_plus_b200 import _capi
    from surel_plus_b200.parallel import partition
    from surel_plus_b200.pprgo import topk_ppr_matrix
    lo, hi = partition(c.n, c.world, c.rank)          # seeds partitioned over the ranks; rows are independent
    idx = c.query[lo:hi]
    topk_ppr_matrix(c.graph, W["alpha"], W["eps"], idx[: max(len(idx) // 50, 1)], W["topk"], "sym", encoder="PPR").close()
    _capi.timing_enable(True)
    _capi.timing_read(_capi.TIMING_PPR)
    if c.rank == 0:
        c.clocks.start()
    launches0 = _capi.launch_count()
    c.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    x, pushes = None, 0
    steps = max(1, min(args.steps, 3))
    for i in range(steps):
        if x is not None:
            x.close()
        x = topk_ppr_matrix(c.graph, W["alpha"], W["eps"], idx, W["topk"], "sym", encoder="PPR")
        pushes += x.pushes
    ev1.record()
    c.barrier()
    launches = _capi.launch_count() - launches0
    ms_total = c.max_over_ranks(ev0.elapsed_time(ev1))
    k_ms, k_n = _capi.timing_read(_capi.TIMING_PPR)
    _capi.timing_enable(False)
    clk = c.clocks.stop() if c.rank == 0 else None
    # sum over pushes (8 + 8 deg(u)) + 8 topk out; the degree of every pushed node is not recorded, so the
    # edge-weighted mean degree (a push lands on u with probability ~ deg(u)) stands in for it, and that is stated
    d = c.deg[c.deg > 0].astype(np.float64)
    deg_push = float((d * d).sum() / d.sum())
    alg = pushes / steps * (8 + 8 * deg_push) + 8.0 * x.T
    # one launch of ppr_push_fast_kernel takes every seed of the step; the general kernel's one or two small relaunches (the
    # handful of seeds that outgrew the fast path) are counted into the same time
    k_avg = k_ms / steps
    per_launch = alg
    ach = per_launch / (k_avg / 1e3) / 1e9
    roof = {"bound": "hbm", "kernel": "ppr_push_fast_kernel (+ ppr_push_kernel for the seeds that outgrow its shared-memory queue / p-list)", "achieved": ach, "peak": c.peaks["hbm_gbs"], "unit": "GB/s",
            "frac": ach / c.peaks["hbm_gbs"], "traffic": load_traffic("ppr_push", args), "peak_source": c.peaks["source"],
            "algorithmic_bytes_per_launch": per_launch, "kernel_ms_per_launch": k_avg, "kernel_launches_per_step": k_n / steps,
            "kernel_share_of_step": k_ms / ms_total, "pushes_per_step": pushes / steps, "pushes_per_s": pushes / (k_ms / 1e3),
            "bytes_model": f"pushes x (8 + 8 x {deg_push:.1f} edge-weighted mean degree) + 8 x nnz"}
    head = {"value": c.n * steps / (ms_total / 1e3), "ms_per_step": ms_total / steps, "steps": steps,
            "scaling": "strong", "clocks": clk, "gpu_launches": int(launches), "roofline": roof}
    return head, x


def bench_e2e(c):
    """(ii) subg_matrix(G, idx): pageable scipy CSR in -> joinable device SpG + LP table out (what the reference arm
    times as gset_sampler + CSR build); (i) gset_sampler: numpy in / numpy out."""
    torch, args, W = c.torch, c.args, c.W
    from surel_plus_b200 import subg_acc
    steps = max(1, args.e2e_steps or args.steps)
    A = c.A
    if W["kind"] == "ppr":
        from surel_plus_b200.parallel import partition
        from surel_plus_b200.pprgo import topk_ppr_matrix
        lo, hi = partition(c.n, c.world, c.rank)
        idx = c.query[lo:hi]

        def call(i):
            x = topk_ppr_matrix(A, W["alpha"], W["eps"], idx, W["topk"], "sym", device=c.dev, encoder="PPR")
            x.close()
            return 8
        api = "topk_ppr_matrix(adj scipy CSR, alpha, eps, idx, topk, 'sym') + encoding 'PPR' -> device value SpG"
        steps = min(steps, 2)
    elif c.world == 1:
        from surel_plus_b200.sampler import subg_matrix

        def call(i):
            with _Quiet():
                z, enc = subg_matrix(A, c.query, num_walks=W["M"], num_steps=W["m"] + 1, device=c.dev, seed=1000 + i)
            z.close()
            return enc.nbytes
        api = "surel_plus_b200.sampler.subg_matrix(G scipy CSR, idx, num_walks, num_steps) -> (device SpG, enc)"
    else:
        from surel_plus_b200.parallel import sharded_subg_matrix

        def call(i):
            z, enc = sharded_subg_matrix(A, c.query, num_walks=W["M"], num_steps=W["m"] + 1, device=c.dev, seed=1000 + i)
            z.close()
            return enc.nbytes
        api = "surel_plus_b200.parallel.sharded_subg_matrix(G scipy CSR, idx, ...) -> (replicated device SpG, enc)"
    d2h = 0
    for i in range(2):
        d2h = call(i)
    c.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        d2h = call(10 + i)
    torch.cuda.synchronize()
    c.barrier()
    ms = c.max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e = {"value": c.n * steps / (ms / 1e3), "unit": "seeds/s", "ms_per_step": ms / steps, "steps": steps,
           "h2d_bytes_per_step": int(A.indptr.nbytes + A.indices.nbytes + c.query.nbytes), "d2h_bytes_per_step": int(d2h),
           "inputs": "pageable numpy / scipy arrays", "api": api}
    e2e_numpy = None
    if W["kind"] == "lp" and c.world == 1:
        os.environ["SUBG_RNG"] = "philox"
        indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
        out = None
        for i in range(3):  # the pinned result buffers alternate
            out = subg_acc.gset_sampler(indptr, indices, c.query, num_walks=W["M"], num_steps=W["m"], seed=5 + i, device=c.dev)
        t0 = time.perf_counter()
        for i in range(steps):
            out = subg_acc.gset_sampler(indptr, indices, c.query, num_walks=W["M"], num_steps=W["m"], seed=100 + i, device=c.dev)
        torch.cuda.synchronize()
        ms2 = (time.perf_counter() - t0) * 1e3
        e2e_numpy = {"value": c.n * steps / (ms2 / 1e3), "unit": "seeds/s", "ms_per_step": ms2 / steps, "steps": steps,
                     "h2d_bytes_per_step": int(indptr.nbytes + indices.nbytes + c.query.nbytes),
                     "d2h_bytes_per_step": int(sum(a.nbytes for a in out)), "inputs": "pageable numpy arrays",
                     "api": "surel_plus_b200.subg_acc.gset_sampler(indptr, indices, query, ...) numpy in / numpy out"}
        del out
    return e2e, e2e_numpy


def bench_spjoin(c, spg, B):
    torch, args, W = c.torch, c.args, c.W
    from surel_plus_b200 import _capi, gather, hgather
    rng = np.random.default_rng(7 + c.rank)
    nb = 8
    triplet = W["join"] == "triplet"
    if spg.value_kind == 0:
        xpe = (torch.from_numpy(spg.enc_table()).float() / W["M"]).to(c.dev)
        kdim = xpe.shape[1]
        out_desc = f"float32 [N,2,{kdim}] fused LP lookup"
    else:
        xpe, kdim = None, 1
        out_desc = "float32 [N,2,1] values (PPR encoder)"
    if triplet:
        batches = [make_triplets(c.deg, c.indptr, c.indices, c.n, B, rng) for _ in range(nb)]
    else:
        batches = [make_queries(c.deg, c.indptr, c.indices, c.n, B, W["k"], rng) for _ in range(nb)]
    B = int(batches[0].shape[1])
    batches = [b[:, :B] for b in batches if b.shape[1] >= B]
    nb = len(batches)
    dev_batches = [torch.from_numpy(b).to(c.dev) for b in batches]
    pin_batches = [torch.from_numpy(b).pin_memory() for b in batches]
    sizes = spg.set_sizes().cpu().numpy()
    if triplet:
        rows = [int(sizes[b[0]].sum() + sizes[b[1]].sum() + 2 * sizes[b[2]].sum()) for b in batches]
        nseg = 4
        join = lambda e: hgather(e, spg, c.dev, xpe)                      # noqa: E731
    else:
        rows = [int(sizes[b[0]].sum() + sizes[b[1]].sum()) for b in batches]
        nseg = 2
        join = lambda e: gather(e, spg, c.dev, True, xpe)                 # noqa: E731
    for i in range(2 * nb):  # every batch shape once: the output-size estimate and torch's block cache settle
        join(dev_batches[i % nb])
    _capi.timing_enable(True)
    _capi.timing_read(1)
    c.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = int(min(max(args.steps, 1) * 40 * max(1.0, 4096 / B), 2000))   # tens of ms of work per block
    per_call = []
    e0.record()
    for i in range(reps):
        t_call = time.perf_counter()
        xz, ptr = join(dev_batches[i % nb])     # returns after its one stream synchronisation
        per_call.append(time.perf_counter() - t_call)
    e1.record()
    c.barrier()
    ms = c.max_over_ranks(e0.elapsed_time(e1))
    k_ms, k_n = _capi.timing_read(1)
    _capi.timing_enable(False)
    rows_avg = float(np.mean([rows[i % nb] for i in range(reps)]))
    if spg.value_kind == 0:
        alg = 24.0 * nseg * B + (8 + 8 * kdim) * rows_avg  # SURVEY 8(d), fused-feature form
    else:
        alg = 24.0 * nseg * B + (12 + 8) * rows_avg         # int32 id + float64 value in, float32 pair out
    ach = alg / (k_ms / max(k_n, 1) / 1e3) / 1e9
    # e2e: pinned host edges in, checksum scalar back (what a training step does with the loss)
    for i in range(nb):
        join(pin_batches[i])
    c.barrier()
    t0 = time.perf_counter()
    chk = 0.0
    for i in range(reps):
        xz, ptr = join(pin_batches[i % nb])
        chk += float(xz.reshape(-1)[-1].item())
    c.barrier()
    e2e_s = c.max_over_ranks(time.perf_counter() - t0)
    # ---- the same batches through a JoinStream: CUDA-graph replay per batch, host edges in, no host synchronisation
    stream_blk = None
    try:
        from surel_plus_b200 import JoinStream
        js = JoinStream(spg, B, c.dev, encode=xpe, arity=3 if triplet else 2, segid=triplet, depth=4)
        for i in range(2 * nb):
            js.submit(batches[i % nb])
        assert js.rows() == rows[(2 * nb - 1) % nb]
        c.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s0.record()
        for i in range(reps):
            js.submit(batches[i % nb])
        s1.record()
        t_submit = time.perf_counter() - t0
        last_rows = js.rows()
        c.barrier()
        ms_s = c.max_over_ranks(s0.elapsed_time(s1))
        k_per = k_ms / max(k_n, 1)
        stream_blk = {"value": c.world * B * reps / (ms_s / 1e3), "unit": "queries/s", "ms_per_batch": ms_s / reps,
                      "host_us_per_submit": t_submit / reps * 1e6, "kernel_share_of_batch": k_per / (ms_s / reps),
                      "rows_last_batch": int(last_rows), "h2d_bytes_per_step": 8 * batches[0].shape[0] * B, "d2h_bytes_per_step": 16,
                      "what": "JoinStream.submit: pageable host edges -> pinned staging -> [plan, join, row count to pinned memory] "
                              "replayed as one CUDA graph; output stays on the device, nothing is synchronised per batch"}
        js.close()
    except Exception as ex:  # noqa: BLE001
        stream_blk = {"error": repr(ex)}
    return {"value": c.world * B * reps / (ms / 1e3), "unit": "queries/s", "batch": B, "pattern": W["join"], "output": out_desc,
            "stream": stream_blk,
            "avg_rows_per_batch": rows_avg, "avg_set_size": rows_avg / ((4 if triplet else 2) * B), "ms_per_batch": ms / reps,
            "ms_per_batch_median": float(np.median(per_call)) * 1e3, "ms_per_batch_max": float(np.max(per_call)) * 1e3, "calls": reps,
            "e2e": {"value": c.world * B * reps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": 8 * batches[0].shape[0] * B,
                    "d2h_bytes_per_step": 4},
            "roofline": {"bound": "hbm", "kernel": "spjoin_kernel", "achieved": ach, "peak": c.peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": ach / c.peaks["hbm_gbs"], "traffic": load_traffic("spjoin", args) if B == 21504 else None,
                         "algorithmic_bytes_per_launch": alg, "kernel_ms_per_launch": k_ms / max(k_n, 1),
                         "kernel_share_of_batch": k_ms / ms}}


def cpu_baseline(c, spg, spjoin):
    args, W, A = c.args, c.W, c.A
    rng = np.random.default_rng(0)
    cores = os.cpu_count() or 1
    with _Quiet():
        q, run, kind, what = reference_sampler(W, A, args.ref_seconds, rng, cores)
        t_s, t_all = run(q)
    out = {"value": len(q) / t_all, "unit": "seeds/s", "cores": cores, "kind": kind,
           "sample": f"{len(q)} random seeds of {A.shape[0]}; {what}; {t_all:.1f}s",
           "sampler_only": {"value": len(q) / t_s, "unit": "seeds/s"}}
    # SpJoin CPU baseline: the reference's scipy formulation on the same SpG at the same batch sizes: gather
    # (1 thread, train.py:13-45) and pgather (njobs = 4 threads, train.py:88-111); hgather for triplets
    try:
        from oracle import pyoracle as po
        z = spg.to_scipy()
        res = []
        for blk in spjoin:
            B = blk.get("batch")
            if not B:
                continue
            if W["join"] == "triplet":
                e = make_triplets(c.deg, c.indptr, c.indices, c.n, B, rng)
                fns = {"hgather (1 thread)": lambda: po.scipy_triplet_join(e, z)}
            else:
                e = make_queries(c.deg, c.indptr, c.indices, c.n, B, W["k"], rng)
                fns = {"gather (1 thread)": lambda: po.scipy_pair_join(e, z), "pgather(njobs=4)": lambda: po.scipy_pgather(e, z, 4)}
            for name, fn in fns.items():
                t0 = time.perf_counter()
                cnt = 0
                while time.perf_counter() - t0 < 3.0:
                    fn()
                    cnt += 1
                dt = time.perf_counter() - t0
                res.append({"batch": B, "call": name, "value": cnt * B / dt, "unit": "queries/s", "kind": "port",
                            "cores": 1 if "1 thread" in name else 4,
                            "sample": f"{cnt} batches of {B} queries, scipy CSR algebra as train.py:77-84"})
        out["spjoin"] = res
    except Exception as ex:  # pragma: no cover
        out["spjoin"] = [{"error": repr(ex)}]
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "MEASURED_PEAKS.json (measured)"}
    return {"hbm_gbs": 6650.0, "source": "B200_PROFILING.md fallback"}


def load_traffic(kernel: str, args):
    """Per-launch dram bytes from the committed ncu capture of this workload (profiles/), else null."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
        return d.get(f"{args.workload}:{kernel}") if args.scale == 1.0 else None
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ppa", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (smoke runs only)")
    ap.add_argument("--spjoin-batch", type=int, nargs="*", default=None, help="queries per SpJoin call (default: the workload's sizes)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = --steps")
    ap.add_argument("--ref-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="kernel iteration: the sampling pass only (no e2e, SpJoin, CPU baseline)")
    ap.add_argument("--no-replicas", action="store_true", help="N > 1: skip the secondary independent-replicas number")
    ap.add_argument("--no-exchange-compare", action="store_true", help="N > 1: skip the NCCL-staged exchange comparison")
    ap.add_argument("--linked", action="store_true", help="N > 1: also time the linked-shards pass (no replication, remote rows read "
                    "over NVLink at join time) and SpJoin on it; measured at 2 GPUs (profiles/r3m_bench_ppa_2gpu.json), opt-in so that "
                    "an unmeasured GPU count never sits between the headline and its JSON line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
