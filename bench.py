#!/usr/bin/env python
"""bench.py -- SubGAcc hot path on B200: sampled node-sets/s (+ SpJoin queries/s).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)

Workload (BASELINE.json configs[1]): synthetic ogbl-ppa-shape graph (576 289 nodes, 30.3 M
undirected edges), LP encoder, CLI num_steps=4 (walk length m=3), num_walks=200; every node
is a seed.  One *step* = one pass of sample -> LP-encode -> SpG build over all seeds of the
rank, graph resident in HBM, SpG left resident in HBM.  `value` = seeds/s over all ranks.
N>1 is weak scaling: the graph is replicated, rank r runs sampling round r (its own Philox
stream) over all seeds, no data-path collective.  SpJoin is timed afterwards on the resident
SpG and reported under "spjoin".  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("SUBG_QUIET", "1")

WORKLOADS = {
    # name: (graph shape, num_walks, walk length m, neg ratio k)
    "ppa": ("ppa", 200, 3, 20),
    "collab": ("collab", 200, 2, 10),
    "dblp": ("dblp", 100, 2, 10),
}


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


def make_graph(workload: str, scale: float):
    from surel_plus_b200.graphs import named_graph
    shape, M, m, k = WORKLOADS[workload]
    t0 = time.time()
    A = named_graph(shape, scale)
    log(f"[bench] graph {shape} x{scale}: N={A.shape[0]} nnz={A.nnz} ({time.time() - t0:.1f}s)")
    return A, M, m, k


def seed_algorithmic_bytes(deg: np.ndarray, M: int, m: int, T: int) -> float:
    """B_seed = 24 + 4*min(d,M) + 12*M*(m-1) + 8*|S_u| (d > 0), 32 for d = 0."""
    live = deg > 0
    return float(24 * live.sum() + 4 * np.minimum(deg[live], M).sum() + 12 * M * (m - 1) * live.sum()
                 + 32 * (~live).sum() + 8 * (T - (~live).sum()))


def make_queries(A, B: int, k: int, rng) -> np.ndarray:
    """1 positive : k negatives per batch (ogbl-ppa training pattern, main.py --k)."""
    npos = max(B // (k + 1), 1)
    rows = rng.integers(0, A.shape[0], npos * 4)
    rows = rows[np.diff(A.indptr)[rows] > 0][:npos]
    pos_v = A.indices[A.indptr[rows] + (rng.integers(0, 1 << 30, len(rows)) % np.diff(A.indptr)[rows])]
    pos = np.stack([rows, pos_v])
    neg = rng.integers(0, A.shape[0], (2, B - pos.shape[1]))
    e = np.concatenate([pos, neg], axis=1).astype(np.int64)
    return e[:, rng.permutation(e.shape[1])]


# ------------------------------------------------------------------------------ reference arm
def reference_sampler_rate(A, M, m, target_s: float, rng, nthread: int):
    """Time the UNMODIFIED reference gset_sampler (oracle/_ref) + the subg_matrix CSR build on a
    bounded random sample of seeds sized for about `target_s` seconds."""
    from oracle import reference as ref, pyoracle as po
    subg = ref.subg_acc()
    kind = "reference"
    n = A.shape[0]
    indptr = A.indptr.astype(np.int32)
    indices = A.indices.astype(np.int32)

    def run(q):
        t0 = time.perf_counter()
        if subg is not None:
            nsize, remap, enc = subg.gset_sampler(indptr, indices, q, num_walks=M, num_steps=m, nthread=nthread)
        else:
            nsize, remap, enc = po.gset_sampler_replay(indptr, indices, q, M, m)
        po.subg_matrix_from(nsize, remap, enc, q, n, m + 1)  # random_walks.py:79-81
        return time.perf_counter() - t0

    if subg is None:
        kind = "port"
    perm = rng.permutation(n).astype(np.int32)
    probe = perm[: min(n, 4000)]
    t = run(probe)
    S = int(min(n, max(len(probe), len(probe) * target_s / max(t, 1e-3))))
    return perm[:S], run, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    A, M, m, k = make_graph(args.workload, args.scale)
    rng = np.random.default_rng(0)
    cores = os.cpu_count() or 1
    # fd-level redirect: the reference prints '#SubGAcc' lines from C
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)
    try:
        q, run, kind = reference_sampler_rate(A, M, m, args.ref_seconds, rng, cores)
        for _ in range(args.warmup):
            run(q[: max(len(q) // 8, 1)])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run(q)
        dt = time.perf_counter() - t0
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
    val = len(q) * args.steps / dt
    sample = f"{len(q)} random seeds of {A.shape[0]} per step, gset_sampler(nthread={cores}) + scipy CSR build"
    line = {
        "impl": "reference", "metric": "sampled_node_sets_per_sec", "value": val, "unit": "seeds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, A, M, m, k),
        "cpu_baseline": {"value": val, "unit": "seeds/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "seeds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, A, M, m, k):
    return {"workload": f"synthetic ogbl-{args.workload}-shape graph x{args.scale:g} (N={A.shape[0]}, "
                        f"directed nnz={A.nnz}), LP encoder, num_walks={M}, walk length m={m} (CLI num_steps={m + 1}), "
                        f"all nodes are seeds, neg ratio k={k}",
            "rng": "philox4x32-10", "l2": "graph + SpG output >> 126 MB L2; no flush needed",
            "spjoin_batch": args.spjoin_batch}


# ------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    from surel_plus_b200 import DeviceGraph, SpG, _capi, gather, subg_acc
    from surel_plus_b200 import graphs  # noqa: F401

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    A, M, m, k = make_graph(args.workload, args.scale)
    n = A.shape[0]
    deg = np.diff(A.indptr)
    query = np.arange(n, dtype=np.int32)
    graph = DeviceGraph.from_scipy(A, dev)
    q_dev = torch.from_numpy(query).to(dev)
    base_seed = 111413 + 1000003 * rank

    # ---- value: graph resident in HBM, SpG left in HBM --------------------------------------
    def step(i):
        # what subg_matrix builds: the SpG (sorted CSR-of-sets + LP table), resident and joinable
        return SpG.sample(graph, q_dev, num_walks=M, num_steps=m, seed=base_seed + i, rng_mode=_capi.SUBG_RNG_PHILOX,
                          first_visit_ranks=False)

    for i in range(args.warmup):
        step(i).close()
    _capi.timing_enable(True)
    for w in (0, 1, 2):
        _capi.timing_read(w)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = _capi.launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    T_sum, spg = 0, None
    for i in range(args.steps):
        if spg is not None:
            spg.close()
        spg = step(args.warmup + i)
        T_sum += spg.T
    ev1.record()
    barrier()
    launches = _capi.launch_count() - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    k_ms, k_launches = _capi.timing_read(_capi_const("SAMPLER"))
    b_ms, _ = _capi.timing_read(_capi_const("BUILD"))
    _capi.timing_enable(False)
    clk = clocks.stop() if rank == 0 else None
    value = world * n * args.steps / (ms_total / 1e3)
    T_avg = T_sum / args.steps
    alg_bytes = seed_algorithmic_bytes(deg, M, m, T_avg)
    peaks = load_peaks()
    k_avg_ms = k_ms / max(k_launches, 1)
    achieved = alg_bytes / (k_avg_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": "gset_sample_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": load_traffic("gset_sample", args),
                "peak_source": peaks["source"], "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms_per_launch": k_avg_ms, "kernel_share_of_step": k_ms / ms_total,
                "spg_build_ms_per_step": b_ms / args.steps, "avg_set_size": T_avg / n, "unique_lp_rows": int(spg.c)}

    # ---- e2e: reference-facing call with host buffers (H2D graph + seeds, D2H nsize/remap/enc) --
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    h_indptr, h_indices, h_query = pin(A.indptr.astype(np.int32)), pin(A.indices.astype(np.int32)), pin(query)
    os.environ["SUBG_RNG"] = "philox"
    d2h = 0
    out = None
    for i in range(max(3, args.warmup)):  # same holding pattern as the timed loop: the pinned result buffers alternate
        out = subg_acc.gset_sampler(h_indptr, h_indices, h_query, num_walks=M, num_steps=m, seed=base_seed + i, device=dev)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.e2e_steps):
        out = subg_acc.gset_sampler(h_indptr, h_indices, h_query, num_walks=M, num_steps=m, seed=base_seed + 100 + i,
                                    device=dev)
        d2h = sum(a.nbytes for a in out)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    e2e = {"value": world * n * args.e2e_steps / (e2e_ms / 1e3), "unit": "seeds/s",
           "h2d_bytes_per_step": int(h_indptr.nbytes + h_indices.nbytes + h_query.nbytes),
           "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
           "api": "surel_plus_b200.subg_acc.gset_sampler(indptr, indices, query, ...) numpy in / numpy out"}
    del out

    # ---- SpJoin on the resident SpG ------------------------------------------------------------
    try:
        spjoin = bench_spjoin(args, torch, dev, spg, A, M, k, gather, _capi, peaks, barrier, max_over_ranks, world)
    except Exception as ex:  # noqa: BLE001 -- the seeds/s headline must survive a failure of the secondary block
        spjoin = {"error": repr(ex)}
        log(f"[bench] SpJoin block failed on rank {rank}: {ex!r}")

    # ---- N > 1: the partitioned form of the same job (BASELINE configs[1]: "seeds sharded over the GPUs"): every rank
    # samples its contiguous seed range, the SpG shards are all-gathered (NCCL over NVLink) and every rank ends with the
    # full SpG.  Reported beside the weak-scaling headline; strong scaling of one sampling pass.
    sharded = None
    if world > 1:
        try:  # the headline above must survive a failure of this secondary block
            from surel_plus_b200.parallel import partition_by_work, sharded_sample
            # contiguous seed ranges balanced by a degree-based estimate of the set size (the synthetic generator puts
            # the hubs at the low ids; equal-count ranges would leave rank 0 with the largest sets)
            bounds = partition_by_work(300.0 + 1.5 * np.minimum(deg, M), world)
            spg.close()
            spg = None
            torch.cuda.empty_cache()
            for i in range(2):
                sharded_sample(graph, query, num_walks=M, num_steps=m, seed=base_seed + i, bounds=bounds).close()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(2, min(args.steps, 5))
            s0.record()
            for i in range(reps):
                if spg is not None:
                    spg.close()
                spg = sharded_sample(graph, query, num_walks=M, num_steps=m, seed=111413 + i, bounds=bounds)
            s1.record()
            barrier()
            sh_ms = max_over_ranks(s0.elapsed_time(s1)) / reps
            xbytes = float(getattr(spg, "exchange_bytes", 0))
            torch.cuda.synchronize()
            x_ms = max_over_ranks(spg.exchange_events[0].elapsed_time(spg.exchange_events[1]))
            sharded = {"value": n / (sh_ms / 1e3), "unit": "seeds/s", "ms_per_pass": sh_ms, "scaling": "strong",
                       "what": "seed ranges sampled per rank + NCCL all-gather of the SpG shards; every rank holds the full SpG",
                       "allgather_bytes_total": xbytes, "received_bytes_per_gpu": xbytes * (world - 1) / world,
                       "exchange_ms": x_ms, "exchange_GBps_per_gpu": xbytes * (world - 1) / world / (x_ms / 1e3) / 1e9}
        except Exception as ex:  # noqa: BLE001
            sharded = {"error": repr(ex)}
            log(f"[bench] sharded block failed on rank {rank}: {ex!r}")

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, A, M, m, spg, spjoin)
    if spg is not None:
        spg.close()

    if rank == 0:
        line = {
            "metric": "sampled_node_sets_per_sec", "value": value, "unit": "seeds/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": workload_config(args, A, M, m, k), "clocks": clk,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "spjoin": spjoin,
        }
        if sharded is not None:
            line["sharded"] = sharded
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _capi_const(name):
    return {"SAMPLER": 0, "SPJOIN": 1, "BUILD": 2}[name]


def bench_spjoin(args, torch, dev, spg, A, M, k, gather, _capi, peaks, barrier, max_over_ranks, world):
    rng = np.random.default_rng(7)
    B = args.spjoin_batch
    nb = 8
    xpe = (torch.from_numpy(spg.enc_table()).float() / M).to(dev)
    kdim = xpe.shape[1]
    batches = [make_queries(A, B, k, rng) for _ in range(nb)]
    dev_batches = [torch.from_numpy(b).to(dev) for b in batches]
    pin_batches = [torch.from_numpy(b).pin_memory() for b in batches]
    sizes = spg.set_sizes().cpu().numpy()
    rows = [int(sizes[b[0]].sum() + sizes[b[1]].sum()) for b in batches]
    # device-resident edges, fused fp32 feature output [N,2,k]
    for i in range(2 * nb):  # every batch shape once: the output-size estimate and torch's block cache settle
        gather(dev_batches[i % nb], spg, dev, True, xpe)
    _capi.timing_enable(True)
    _capi.timing_read(1)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(args.steps, 1) * 40     # ~50 ms of work: one scheduling hiccup of the host must not decide the number
    per_call = []
    e0.record()
    for i in range(reps):
        t_call = time.perf_counter()
        xz, ptr = gather(dev_batches[i % nb], spg, dev, True, xpe)     # returns after its one stream synchronisation
        per_call.append(time.perf_counter() - t_call)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    k_ms, k_n = _capi.timing_read(1)
    _capi.timing_enable(False)
    rows_avg = float(np.mean([rows[i % nb] for i in range(reps)]))
    alg = 48.0 * B + (8 + 8 * kdim) * rows_avg  # SURVEY 8(d), fused-feature form
    ach = alg / (k_ms / max(k_n, 1) / 1e3) / 1e9
    # e2e: pinned host edges in, checksum scalar back (what a training step does with the loss)
    for i in range(nb):
        gather(pin_batches[i], spg, dev, True, xpe)
    barrier()
    t0 = time.perf_counter()
    chk = 0.0
    for i in range(reps):
        xz, ptr = gather(pin_batches[i % nb], spg, dev, True, xpe)
        chk += float(xz[-1, 0, 0].item())
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    return {"value": world * B * reps / (ms / 1e3), "unit": "queries/s", "batch": B, "output": f"float32 [N,2,{kdim}] fused LP lookup",
            "avg_rows_per_batch": rows_avg, "avg_set_size": rows_avg / (2 * B), "ms_per_batch": ms / reps,
            "ms_per_batch_median": float(np.median(per_call)) * 1e3, "ms_per_batch_max": float(np.max(per_call)) * 1e3, "calls": reps,
            "e2e": {"value": world * B * reps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": 16 * B, "d2h_bytes_per_step": 12},
            "roofline": {"bound": "hbm", "kernel": "spjoin_kernel", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": ach / peaks["hbm_gbs"], "traffic": load_traffic("spjoin", args),
                         "algorithmic_bytes_per_launch": alg, "kernel_ms_per_launch": k_ms / max(k_n, 1),
                         "kernel_share_of_batch": k_ms / ms}}


def cpu_baseline(args, A, M, m, spg, spjoin):
    rng = np.random.default_rng(0)
    cores = os.cpu_count() or 1
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)
    try:
        q, run, kind = reference_sampler_rate(A, M, m, args.ref_seconds, rng, cores)
        dt = run(q)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
    out = {"value": len(q) / dt, "unit": "seeds/s", "cores": cores, "kind": kind,
           "sample": f"{len(q)} random seeds of {A.shape[0]}, gset_sampler(nthread={cores}) + scipy CSR build, {dt:.1f}s"}
    # SpJoin CPU baseline: the reference's scipy formulation (train.py:77-84) on the same SpG, 1 thread
    try:
        from oracle import pyoracle as po
        z = spg.to_scipy()
        B = 1024
        e = make_queries(A, B, 20, rng)
        t0 = time.perf_counter()
        cnt = 0
        while time.perf_counter() - t0 < 5.0:
            po.scipy_pair_join(e, z)
            cnt += 1
        out["spjoin"] = {"value": cnt * B / (time.perf_counter() - t0), "unit": "queries/s", "cores": 1, "kind": "port",
                         "sample": f"{cnt} batches of {B} pair queries, scipy CSR algebra as train.py:77-84"}
    except Exception as ex:  # pragma: no cover
        out["spjoin"] = {"error": repr(ex)}
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
    ap.add_argument("--spjoin-batch", type=int, default=21504, help="queries per SpJoin call (1024 x (1 pos + 20 neg))")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
