"""torchrun --nproc-per-node 2 scripts/nccl_probe.py : what do the all-gather flavours reach on this box?"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    per = 284_000_000 // world  # int32 entries per rank (ppa shard)
    mine = per + 1000 * rank
    counts = [per + 1000 * r for r in range(world)]
    local_t = torch.full((mine,), rank, dtype=torch.int32, device="cuda")
    out = torch.empty(sum(counts), dtype=torch.int32, device="cuda")
    offs = [sum(counts[:r]) for r in range(world + 1)]
    slices = [out[offs[r]:offs[r + 1]] for r in range(world)]
    t_uneven = timeit(lambda: dist.all_gather(slices, local_t))
    eq_local = local_t[:per]
    eq_out = torch.empty(world * per, dtype=torch.int32, device="cuda")
    t_equal = timeit(lambda: dist.all_gather_into_tensor(eq_out, eq_local))
    eq_slices = [eq_out[r * per:(r + 1) * per] for r in range(world)]
    t_list_equal = timeit(lambda: dist.all_gather(eq_slices, eq_local))

    def p2p():
        ops = []
        for r in range(world):
            if r != rank:
                ops.append(dist.P2POp(dist.isend, local_t, r))
                ops.append(dist.P2POp(dist.irecv, slices[r], r))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    t_p2p = timeit(p2p)
    if rank == 0:
        gb = per * 4 * (world - 1) / 1e9
        print(f"[nccl probe] world={world} env={ {k: v for k, v in os.environ.items() if k.startswith('NCCL_') and 'FILE' not in k} } {gb:.2f} GB received per GPU: all_gather(list, uneven)={t_uneven:.2f} ms ({gb / t_uneven * 1e3:.0f} GB/s), "
              f"all_gather_into_tensor={t_equal:.2f} ms ({gb / t_equal * 1e3:.0f} GB/s), all_gather(list, equal)={t_list_equal:.2f} ms, "
              f"send/recv={t_p2p:.2f} ms ({gb / t_p2p * 1e3:.0f} GB/s)", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
