"""torchrun worker of tests/test_gpu_shard.py::test_sharded_sample_nccl_two_gpus (one rank per GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SUBG_QUIET", "1")


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from surel_plus_b200 import DeviceGraph, SpG, _capi, gather
    from surel_plus_b200.graphs import synthetic_graph
    from surel_plus_b200.parallel import sharded_sample, partition
    A = synthetic_graph(20_000, 120_000, seed=7, gamma=2.0, isolated=3)
    n = A.shape[0]
    q = np.arange(n, dtype=np.int32)
    g = DeviceGraph.from_scipy(A, f"cuda:{local}")
    modes_seen = set()
    for xmode in ("peer", "nccl"):
        for mode in (_capi.SUBG_RNG_RAND_R, _capi.SUBG_RNG_PHILOX):
            full = SpG.sample(g, q, 60, 3, seed=5, rng_mode=mode)
            rep = sharded_sample(g, q, 60, 3, seed=5, rng_mode=mode, mode=xmode)
            modes_seen.add(rep.exchange_mode)
            xz0, p0 = gather(np.stack([q[:500], q[::-1][:500]]).astype(np.int64), full, f"cuda:{local}", True, None)
            xz1, p1 = gather(np.stack([q[:500], q[::-1][:500]]).astype(np.int64), rep, f"cuda:{local}", True, None)
            assert torch.equal(xz0, xz1) and torch.equal(p0, p1)     # the scattered result joins in place
            fv, rv = full.views(), rep.views()
            for k in ("indptr", "indices", "data"):
                assert torch.equal(fv[k], rv[k]), k
            assert np.array_equal(full.enc_table(), rep.enc_table())
            assert rep.exchange_bytes >= 4 * full.T
            if not (xmode == "nccl" and mode == _capi.SUBG_RNG_PHILOX):
                rep.close()
                full.close()
    print("EXCHANGE_MODES", sorted(modes_seen), flush=True)
    # queries sliced per rank, joined locally on the replicated SpG, gathered for comparison
    xpe = torch.from_numpy(rep.enc_table()).float().cuda() / 60
    edge = np.random.default_rng(0).integers(0, n, (2, 1024))
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = partition(1024, world, rank)
    mine, _ = gather(edge[:, lo:hi], rep, f"cuda:{local}", True, xpe)
    whole, ptr = gather(edge, full, f"cuda:{local}", True, xpe)
    B = 1024
    ip = ptr.cpu().numpy()
    left = whole[ip[lo]:ip[hi]]
    right = whole[ip[B + lo]:ip[B + hi]]
    assert torch.equal(mine, torch.cat([left, right]))
    # the no-replication alternative: linked shards, remote rows read over NVLink by the join kernel (TMA bulk copies)
    from surel_plus_b200.parallel import linked_sample
    for mode in (_capi.SUBG_RNG_RAND_R, _capi.SUBG_RNG_PHILOX):
        ref = SpG.sample(g, q, 60, 3, seed=5, rng_mode=mode, first_visit_ranks=False)
        lk = linked_sample(g, q, 60, 3, seed=5, rng_mode=mode)
        assert (lk.n, lk.T, lk.c, lk.max_set) == (ref.n, ref.T, ref.c, ref.max_set)
        assert np.array_equal(lk.enc_table(), ref.enc_table())
        e2 = np.stack([q[:800], q[::-1][:800]]).astype(np.int64)
        x0, p0 = gather(e2, ref, f"cuda:{local}", True, None)
        x1, p1 = gather(e2, lk, f"cuda:{local}", True, None)
        assert torch.equal(x0, x1) and torch.equal(p0, p1)
        torch.cuda.synchronize()
        dist.barrier()
        lk.close()
        ref.close()
        dist.barrier()
    print("LINKED_OK", rank, flush=True)
    dist.barrier()
    print("SHARDED_OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
