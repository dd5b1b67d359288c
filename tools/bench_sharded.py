"""BASELINE config 5: synthetic 10 M-item catalogue, item tables row-sharded over the ranks (tlsan_b200/sharded.py).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded.py [--items 10000000]

Prints one JSON line: train samples/s (all ranks, device-resident batches), distinct ids and exchange bytes per
rank and step, and the per-phase split of a step (CUDA events, max over ranks)."""
import argparse, json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import tlsan_oracle as O          # default_config only
from tlsan_b200.sharded import ShardedModel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=10_000_000)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--partition", default="mod")
    args = ap.parse_args()
    rank, world, lr_ = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr_)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
        pg = dist.group.WORLD
    NU, NC, L, B = 40000, 673, 10, args.batch
    bench.NI = args.items                      # the generators draw item ids from bench.NI
    bench.NU, bench.NC = NU, NC
    cfg = O.default_config(NU, args.items, NC, Ls=L)
    icl = np.random.default_rng(1234).integers(0, NC, args.items).astype(np.int32)
    m = ShardedModel(cfg, icl, process_group=pg, partition=args.partition)
    hb = bench.synth_batches(np.random.default_rng(1234 + 1000 * rank), 4, B, L)
    dbs = [m.stage_batch(b) for b in hb]
    for w in range(args.warmup):
        m.train_staged(dbs[w % 4], 1.0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    m.profile = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        m.train_staged(dbs[k % 4], 1.0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    marks, per = m.profile, 6
    names = [marks[i][0] for i in range(1, per)]
    ph = np.zeros(per - 1)
    for s_ in range(args.steps):
        for i in range(1, per):
            ph[i - 1] += marks[s_ * per + i - 1][1].elapsed_time(marks[s_ * per + i][1])
    t = torch.tensor([ms] + list(ph / args.steps), device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        t = t.tolist()
        print(json.dumps({"metric": "train_samples_per_s", "value": B * world / (t[0] * 1e-3), "n_gpus": world,
                          "ms_per_step": t[0], "config": {"workload": "TLSAN synthetic %d-item catalogue, item tables "
                          "row-sharded (%s), NU 40000, NC 673, Ls 10" % (args.items, args.partition), "per_gpu_batch": B},
                          "distinct_ids_per_rank": m.last_unique, "exchange_bytes_per_rank_step": 2 * m.last_exchange_bytes,
                          "phases_ms": dict(zip(names, t[1:])), "loss": float(m._stats[0].item())}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
