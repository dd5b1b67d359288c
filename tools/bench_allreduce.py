#!/usr/bin/env python
"""Cost of the data-parallel gradient exchange in isolation: NCCL all_reduce of the flat fp32 gradient
buffer of the Electronics shape (12.5 MB) and of its two halves, CUDA-event timed, max over ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_allreduce.py

NCCL reads its environment (NCCL_ALGO, NCCL_PROTO, NCCL_MAX_NCHANNELS, ...) at communicator creation: run once per
setting.  Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    sizes = {"flat_12.5MB": 3_127_000, "user_plane_6.7MB": 1_680_000, "item_plane_5.8MB": 1_447_000, "dense_18KB": 4456}
    bufs = {k: torch.ones(n, dtype=torch.float32, device="cuda") for k, n in sizes.items()}
    out = {"world": world, "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}
    for name, t in bufs.items():
        for _ in range(20):
            dist.all_reduce(t)
        torch.cuda.synchronize()
        dist.barrier()
        iters = 200
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(iters):
            dist.all_reduce(t)
        ev[1].record()
        torch.cuda.synchronize()
        us = torch.tensor([ev[0].elapsed_time(ev[1]) * 1e3 / iters], device="cuda")
        dist.all_reduce(us, op=dist.ReduceOp.MAX)
        out[name + "_us"] = round(float(us.item()), 1)
    # the split exchange: two back-to-back calls
    torch.cuda.synchronize(); dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(200):
        dist.all_reduce(bufs["user_plane_6.7MB"]); dist.all_reduce(bufs["item_plane_5.8MB"])
    ev[1].record(); torch.cuda.synchronize()
    us = torch.tensor([ev[0].elapsed_time(ev[1]) * 1e3 / 200], device="cuda")
    dist.all_reduce(us, op=dist.ReduceOp.MAX)
    out["two_calls_us"] = round(float(us.item()), 1)
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
