"""BASELINE config 4: batched scoring sweep (eval_auc-style, 2 candidates per row), Electronics-shape tables.

    python tools/bench_scoring_sweep.py                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_scoring_sweep.py

B in {1k, 4k, 16k, 64k} x Ls in {10, 30, 50, 70, 90}; "full" = every row at full length (the roofline number),
"mix" = the empirical Digital-Music length law scaled to Ls.  Rows are sharded over ranks with no collective
(SURVEY 8e); time = CUDA events, max over ranks.  Algorithmic bytes per sequence: SURVEY 8d scoring-2 formula."""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import tlsan_oracle as O          # default_config only
from tlsan_b200.model import Model


def scoring2_bytes(batch, L):
    sl = np.asarray(batch[6], np.int64); s = np.asarray(batch[7], np.int64); S = batch[4].shape[1]
    R = 2 * (sl + s) + 4
    return int((4 * (2 * L + S + 6) + 4 * (sl + s + 1) + 128 * R + (4 * L + 4) + 4 + 272).sum())


def main():
    rank, world, lr_ = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr_)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
    icl = np.random.default_rng(1234).integers(0, bench.NC, bench.NI).astype(np.int32)
    rows = []
    for L in (10, 30, 50, 70, 90):
        model = Model(O.default_config(bench.NU, bench.NI, bench.NC, Ls=L), icl, seed=1234)
        for B in (1024, 4096, 16384, 65536):
            for kind in ("full", "mix"):
                rng = np.random.default_rng(7 + rank)
                b = list(bench.synth_batches(rng, 1, B, L)[0])
                if kind == "full":
                    b[6] = np.full(B, L, np.int64)
                    b[3] = rng.integers(0, bench.NI, (B, L)).astype(np.int64)
                    b[5] = (1.0 / np.sort(rng.integers(1, 13, (B, L)), axis=1)[:, ::-1]).astype(np.float32)
                b[2] = rng.integers(0, bench.NI, B).astype(np.int64)          # negative item
                db = model.stage_batch(tuple(b), is_test=True)
                for _ in range(3):
                    model.score_staged(db, 2)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                n = 20 if B <= 16384 else 10
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    model.score_staged(db, 2)
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
                gbs = scoring2_bytes(b, L) / (ms * 1e-3) / 1e9
                rows.append({"Ls": L, "B_per_gpu": B, "rows": kind, "ms": round(ms, 4),
                             "seqs_per_s": round(B * world / (ms * 1e-3)), "algorithmic_GBps_per_gpu": round(gbs, 1),
                             "frac_of_measured_hbm": round(gbs / peak, 4)})
    if rank == 0:
        print(json.dumps({"metric": "eval_seqs_per_s", "n_gpus": world, "candidates": 2, "hbm_peak_gbs": peak, "sweep": rows}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
