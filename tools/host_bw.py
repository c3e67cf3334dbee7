"""Host <-> device copy ceiling of the box, all ranks at once (run under torchrun; development tool).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/host_bw.py

Every rank copies 256 MiB pinned buffers H2D, D2H and both at once, concurrently with the other ranks; rank 0 prints
per-rank and aggregate GB/s plus where each rank's GPU and pinned memory live.  This is the denominator of bench.py's
`e2e` at N > 1: what the host memory / PCIe complex delivers with NO kernels running.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "memc-net_b200"))


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    from memc_b200.host_pipeline import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local) if os.environ.get("BIND", "1") == "1" else {"bound": False}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 256 << 20
    h_a, h_b = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a, d_b = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=8):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        for s in (s1, s2):
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    res = [timed(h2d), timed(d2h), timed(both)]
    t = torch.tensor(res, device="cuda", dtype=torch.float64)
    allr = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, t)
    else:
        allr = [t]
    if rank == 0:
        rows = [[round(float(x), 1) for x in r.tolist()] for r in allr]
        print(json.dumps({"world": world, "per_rank_gbs [h2d, d2h, each way when both run]": rows,
                          "aggregate_h2d": sum(r[0] for r in rows), "aggregate_d2h": sum(r[1] for r in rows),
                          "aggregate_each_way_bidirectional": sum(r[2] for r in rows), "numa_rank0": numa,
                          "cpus_allowed_rank0": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
