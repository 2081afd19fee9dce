#!/usr/bin/env python
"""What the host of a multi-GPU box can feed: every rank copies its own pinned buffer to its GPU in a
loop (plain cudaMemcpyAsync, nothing else running) and the aggregate is reported.  bench.py's
end-to-end number moves 2316 B per cluster over this path, so N x 54 GB/s (PCIe, one GPU) or this
aggregate, whichever is smaller, bounds it.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_ceiling.py"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = 2 << 30
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.fill_(rank + 1)
    devb = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = {}
    for mode in ("all_ranks", "rank0_only"):
        for _ in range(2):
            devb.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        reps = 12
        t0 = time.perf_counter()
        if mode == "all_ranks" or rank == 0:
            for _ in range(reps):
                devb.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gbs = torch.tensor([reps * nbytes / dt / 1e9 if (mode == "all_ranks" or rank == 0) else 0.0], device="cuda")
        if world > 1:
            dist.barrier()
            tot = gbs.clone()
            dist.all_reduce(tot)
            mn = gbs.clone() if mode == "all_ranks" else tot.clone()
            if mode == "all_ranks":
                dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        else:
            tot = mn = gbs
        out[mode] = {"aggregate_gbs": float(tot.item()), "slowest_rank_gbs": float(mn.item())}
    if rank == 0:
        print(json.dumps({"n_gpus": world, "pinned_bytes_per_rank": nbytes, "h2d": out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
