#!/usr/bin/env python
"""bench.py -- TLSAN train-step throughput on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--batch B] [--Ls L] [--workload electronics|movies] [--strong]
                    [--no-cpu-baseline] [--skip-extras] [--no-pipeline]

Headline workload = BASELINE.json configs[1]: "TLSAN Electronics-shape synthetic (40k users, 22k items, 673 cates)
fp32": NU 39 991, NI 22 048, NC 673, generators of SURVEY.md 8d (tlsan_b200/synth.py, numpy default_rng(1234 + rank)),
per-GPU batch 65 536, Ls 10.  A step = one Model.train step (forward, loss, backward, segmented reduce,
L2 + clip + SGD over every table row).  ONE JSON line on stdout:

  value            train samples/s, all ranks, device-resident batches (several distinct batches cycled, the next
                   batch's occurrence sort pipelined behind the current step)
  e2e              the same step through Model.train(sess, batch, lr) with HOST numpy batches: pack -> pinned ->
                   H2D -> step -> loss D2H inside the timed region (>= 200 steps), batch k+1 staged while step k runs
  sustained        >= 3 s of back-to-back steps with the SM clock seen under that load
  roofline         SURVEY 8d byte model: the whole step (headline) and the dominant kernel, against MEASURED_PEAKS.json
  cpu_baseline     the oracle port on the host cores, bounded sample (rank 0, N = 1)
  dp_parity        N > 1: a fixed global batch trained data-parallel (nccl / p2p exchange) and row-sharded must give
                   bit-identical weights on every rank and the weights of a 1-rank step (also with
                   --skip-extras --dp-parity)
  dataset_resident, eval, eval_rank, movies (configs[2], weak + strong), scoring_sweep (configs[3]),
  sharded_10M (configs[4]): the other BASELINE configurations, compact.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tlsan_b200.synth import S_MAX, WORKLOADS, algorithmic_bytes, synth_batches as _synth, table_bytes  # noqa: E402

WL_NAME, NU, NI, NC = WORKLOADS["electronics"]
N_SAMPLES = 561100


def synth_batches(rng, n_batches, B, L, **kw):
    """Batches of the currently selected workload (module globals NU / NI / NC)."""
    return _synth(rng, n_batches, B, L, NU, NI, NC, **kw)


_SAMPLER_SRC = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
bus = sys.argv[1]
try:
    h = nv.nvmlDeviceGetHandleByPciBusId(bus.encode())
except Exception:
    h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[2]))
print("max", nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM), flush=True)
while True:
    try:
        bits = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
    except Exception:
        bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    try:
        pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
    except Exception:
        pw = 0.0
    print(time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), bits, pw, flush=True)
    time.sleep(0.002)
"""


class ClockSampler:
    """SM clock, power and throttle reasons DURING a timed region: a side process polls NVML every ~2 ms (started
    early); window(t0, t1) summarises the samples that fall inside a wall-clock window."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.proc, self.lines = None, []
        try:
            import torch
            p = torch.cuda.get_device_properties(index)
            bus = "%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, bus, str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def pump():
                for ln in self.proc.stdout:
                    self.lines.append(ln)
            self.t = threading.Thread(target=pump, daemon=True)
            self.t.start()
        except Exception as e:                      # pragma: no cover
            self.err = repr(e)

    def wait_ready(self):
        t_end = time.time() + 5.0                  # the side process needs ~1 s to import + nvmlInit
        while self.proc is not None and not self.lines and time.time() < t_end:
            time.sleep(0.01)

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml sampler unavailable"]}
        time.sleep(0.01)
        sm, pw, bits, mx = [], [], 0, None
        for ln in list(self.lines):
            f = ln.split()
            try:
                if f[0] == "max":
                    mx = int(f[1])
                elif t0 <= float(f[0]) <= t1:
                    sm.append(int(f[1])); bits |= int(f[2]); pw.append(float(f[3]))
            except Exception:
                pass
        reasons = sorted(name for bit, name in self.REASONS.items() if bits & bit)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "power_w": statistics.median(pw) if pw else None, "reasons": reasons}

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()


def cpu_port_throughput(L, budget_s, B=1024, seed=99):
    """Oracle (torch-CPU restatement of model.py) train steps on a bounded sample of the workload."""
    import torch
    from oracle import tlsan_oracle as O
    rng = np.random.default_rng(seed)
    cfg = O.default_config(NU, NI, NC, Ls=L)
    params = O.init_params(cfg, seed=1234)
    icl = rng.integers(0, NC, NI).astype(np.int32)
    batches = synth_batches(rng, 4, B, L)
    O.train_step(params, icl, batches[0], 1.0, cfg)                     # warm-up
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget_s:
        r = O.train_step(params, icl, batches[n % 4], 1.0, cfg)
        params = r["new_params"]
        n += 1
    dt = time.perf_counter() - t0
    return n * B / dt, torch.get_num_threads(), "%d steps of B=%d, L=%d, %s (%.1f s)" % (n, B, L, WL_NAME, dt)


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path = the oracle port
    (TF 1.8 cannot be installed), all host threads, bounded sample per step."""
    if rank != 0:
        return
    import torch
    from oracle import tlsan_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    B = 2048
    rng = np.random.default_rng(1234)
    cfg = O.default_config(NU, NI, NC, Ls=args.Ls)
    params = O.init_params(cfg, seed=1234)
    icl = rng.integers(0, NC, NI).astype(np.int32)
    batches = synth_batches(rng, 4, B, args.Ls)
    for w in range(args.warmup):
        params = O.train_step(params, icl, batches[w % 4], 1.0, cfg)["new_params"]
    t0 = time.perf_counter()
    for k in range(args.steps):
        params = O.train_step(params, icl, batches[k % 4], 1.0, cfg)["new_params"]
    dt = time.perf_counter() - t0
    v = args.steps * B / dt
    sample = "each step = B=%d rows of the %s workload (bounded sample of the 65536-row step)" % (B, WL_NAME)
    _emit({
        "impl": "reference", "metric": "train_samples_per_s", "value": v, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, B),
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def workload_config(args, B):
    return {"workload": "%s (NU %d, NI %d, NC %d), train step (fwd+bwd+L2+clip+SGD)" % (WL_NAME, NU, NI, NC),
            "per_gpu_batch": B, "global_batch": B * args.gpus,
            "Ls": args.Ls, "short_max": S_MAX, "lr": 1.0, "optimizer": "sgd", "parallelism": "dp%d" % args.gpus,
            "l2_policy": "several distinct device-resident batches cycled; per-step working set "
                         "(gradient rows + tables + batch) exceeds the 126 MB L2",
            "pipeline": "off" if getattr(args, "no_pipeline", False) else
                        "occurrence sort of batch k+1 enqueued behind the backward kernels of step k"}


_REAL_STDOUT = None


def _quiet_stdout():
    """Route everything libraries print on fd 1 (e.g. NCCL's version banner) to stderr; the ONE JSON
    line of the contract is written to the saved real stdout by _emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


class Ctx:
    """Process-wide handles of one bench run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.pg = None
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.pg = dist.group.WORLD
        assert self.world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.peaks = peaks
        self.hbm = float(peaks.get("hbm_gbs", 6650.0))
        self.hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, n):
        """n calls of fn between barriers, CUDA events on the current stream; ms per call, max over ranks."""
        self.barrier()
        self.e0.record()
        for k in range(n):
            fn(k)
        self.e1.record()
        self.barrier()
        return self.max_over_ranks(self.e0.elapsed_time(self.e1)) / n


def train_loop(model, dbs, steps, pipe, first=0, global_batch=None):
    n = len(dbs)
    for k in range(steps):
        model.train_staged(dbs[(first + k) % n], 1.0, global_batch=global_batch,
                           next_db=dbs[(first + k + 1) % n] if pipe else None)


def make_model(ctx, L, seed=1234, pg="ctx", **kw):
    from oracle import tlsan_oracle as O            # flag defaults (train.py:26-49) only
    from tlsan_b200.model import Model
    cfg = O.default_config(NU, NI, NC, Ls=L)
    icl = np.random.default_rng(1234).integers(0, NC, NI).astype(np.int32)     # same on every rank
    return Model(cfg, icl, seed=seed, process_group=ctx.pg if pg == "ctx" else pg, **kw), cfg, icl


# ---------------------------------------------------------------------------------------------------- dp parity
def dp_parity(ctx):
    """A fixed global batch, trained (a) data-parallel with the NCCL exchange, (b) with the peer-memory exchange,
    (c) with row-sharded item tables, against 1-rank steps on the whole batch computed redundantly on every rank."""
    torch, dist = ctx.torch, ctx.dist
    from tlsan_b200.parallel import shard_rows
    from tlsan_b200.sharded import ShardedModel
    Bg, L = 8192, 10
    rng = np.random.default_rng(4242)
    batches = synth_batches(rng, 2, Bg, L)
    g = torch.Generator().manual_seed(77)
    ref, cfg, icl = make_model(ctx, L, pg=None)
    sd0 = ref.state_dict()
    for k in ("usert_emb", "item_b"):
        sd0[k] = sd0[k] + 0.3 * torch.randn(sd0[k].shape, generator=g)
    for k in list(sd0):
        if k.endswith("/bias"):
            sd0[k] = sd0[k] + 0.1 * torch.randn(sd0[k].shape, generator=g)
    ref.load_state_dict(sd0)
    ref_sd = []
    for b in batches:
        ref.train(None, b, 1.0)
        ref_sd.append({k: v.clone() for k, v in ref.state_dict().items()})
    step = {k: float((ref_sd[-1][k] - sd0[k]).abs().max()) for k in sd0}
    # the gradient of a second map's bias is identically zero (softmax shift invariance): what arrives there is the
    # rounding noise of terms as large as the sibling kernel's gradient, so that tensor's step sets the scale
    scale = {k: max(step[k], step.get(k.replace("/bias", "/W"), 0.0) if k.endswith("bn_dense_map2/linear_map/bias") else 0.0)
             + 1e-12 for k in sd0}
    out = {}

    def compare(sd, want, label):
        err = max(float((sd[k] - want[k]).abs().max()) / (float(want[k].abs().max()) + 1e-12) for k in want)
        err_step = max(float((sd[k] - want[k]).abs().max()) / scale[k] for k in want)
        flat = torch.cat([v.reshape(-1) for v in sd.values()]).cuda().view(torch.int32)
        hi, lo = flat.clone(), flat.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        same = bool(torch.equal(hi, lo))
        e = torch.tensor([err, err_step], device="cuda")
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        out[label] = {"ok": bool(float(e[0].item()) < 1e-5 and float(e[1].item()) < 1e-3 and same),
                      "err": float(e[0].item()), "err_vs_step": float(e[1].item()), "bit_identical_across_ranks": same}

    for mode in ("nccl", "p2p"):
        try:
            m, _, _ = make_model(ctx, L, dp_mode=mode)
            m.load_state_dict(sd0)
            for b in batches:
                local, _ = shard_rows(b, ctx.rank, ctx.world)
                m.train_staged(m.stage_batch(local), 1.0, global_batch=Bg)
            if mode == "p2p":
                m.check_dp_health()
            compare(m.state_dict(), ref_sd[-1], mode)
            del m
        except Exception as e:                                    # reported, not hidden
            out[mode] = {"ok": False, "error": repr(e)[:200]}
    try:
        sm = ShardedModel(cfg, icl, process_group=ctx.pg)
        sm.load_full_state(sd0)
        local, _ = shard_rows(batches[0], ctx.rank, ctx.world)
        sm.train_staged(sm.stage_batch(local), 1.0, global_batch=Bg)
        compare(sm.gather_full_state(), ref_sd[0], "sharded")
        del sm
    except Exception as e:
        out["sharded"] = {"ok": False, "error": repr(e)[:200]}
    out["what"] = ("global batch %d, %s shape, 2 steps (sharded: 1); err = max |w - w_1rank| / max |w_1rank|, err_vs_step = "
                   "max |w - w_1rank| / max |w_1rank - w_0| (fp32 summation order differs between 1 and N ranks), both "
                   "over all variables and ranks; ok = err < 1e-5, err_vs_step < 1e-3 and bit-identical weights on every rank" % (Bg, WL_NAME))
    torch.cuda.synchronize()
    return out


# ---------------------------------------------------------------------------------------------------- extra legs
def leg_other_workload(ctx, name, L, pipe, steps):
    """BASELINE configs[2]: Movies-TV shape, data parallel; weak (per-GPU batch 65 536) and strong (global 65 536)."""
    global WL_NAME, NU, NI, NC
    keep = (WL_NAME, NU, NI, NC)
    WL_NAME, NU, NI, NC = WORKLOADS[name]
    out = {"workload": "%s (NU %d, NI %d, NC %d)" % (WL_NAME, NU, NI, NC)}
    try:
        model, _, _ = make_model(ctx, L)
        for label, B in (("weak", 65536), ("strong", max(1, 65536 // ctx.world))):
            hb = synth_batches(np.random.default_rng(99 + 1000 * ctx.rank), 3, B, L)
            dbs = [model.stage_batch(b) for b in hb]
            train_loop(model, dbs, 4, pipe, global_batch=B * ctx.world)
            ms = ctx.timed(lambda k: model.train_staged(dbs[k % 3], 1.0, global_batch=B * ctx.world,
                                                        next_db=dbs[(k + 1) % 3] if pipe else None), steps)
            bts = float(np.mean([algorithmic_bytes(b, L)["train"] for b in hb])) + 2 * table_bytes(L, NU, NI, NC) + 2 * 4 * 4449
            out[label] = {"per_gpu_batch": B, "global_batch": B * ctx.world, "ms_per_step": ms,
                          "value": B * ctx.world / (ms * 1e-3), "unit": "samples/s",
                          "algorithmic_GBps_per_gpu": bts / (ms * 1e-3) / 1e9, "frac_of_hbm": bts / (ms * 1e-3) / 1e9 / ctx.hbm}
            del dbs
        del model
    finally:
        WL_NAME, NU, NI, NC = keep
    return out


def leg_scoring_sweep(ctx):
    """BASELINE configs[3]: eval_auc-style scoring (2 candidates), rows sharded over the ranks, no collective;
    four corner points of the sweep, every row at full length (the roofline variant)."""
    torch = ctx.torch
    pts = []
    for L in (10, 90):
        model, _, _ = make_model(ctx, L, pg=None)
        for B in (1024, 65536):
            b = synth_batches(np.random.default_rng(7 + ctx.rank), 1, B, L, full=True, is_test=True)[0]
            db = model.stage_batch(b, is_test=True)
            for _ in range(3):
                model.score_staged(db, 2)
            n = 20 if B <= 4096 else 8
            ms = ctx.timed(lambda k: model.score_staged(db, 2), n)
            gbs = algorithmic_bytes(b, L)["scoring2"] / (ms * 1e-3) / 1e9
            pts.append({"Ls": L, "B_per_gpu": B, "ms": ms, "seqs_per_s": B * ctx.world / (ms * 1e-3),
                        "algorithmic_GBps_per_gpu": gbs, "frac_of_hbm": gbs / ctx.hbm})
        del model
    torch.cuda.synchronize()
    return {"what": "eval_auc scoring, 2 candidates, full-length rows, rows sharded (no collective)", "points": pts}


def leg_sharded(ctx, steps=12):
    """BASELINE configs[4]: 10 M-item catalogue, item tables row-sharded over the ranks + all-to-all."""
    global WL_NAME, NU, NI, NC
    from oracle import tlsan_oracle as O
    from tlsan_b200.sharded import ShardedModel
    keep = (WL_NAME, NU, NI, NC)
    WL_NAME, NU, NI, NC = WORKLOADS["items10m"]
    try:
        L, B = 10, 65536
        cfg = O.default_config(NU, NI, NC, Ls=L)
        icl = np.random.default_rng(1234).integers(0, NC, NI).astype(np.int32)
        m = ShardedModel(cfg, icl, process_group=ctx.pg, partition="mod")
        hb = synth_batches(np.random.default_rng(1234 + 1000 * ctx.rank), 3, B, L)
        dbs = [m.stage_batch(b) for b in hb]
        for w in range(3):
            m.train_staged(dbs[w % 3], 1.0, global_batch=B * ctx.world)
        ms = ctx.timed(lambda k: m.train_staged(dbs[k % 3], 1.0, global_batch=B * ctx.world), steps)
        out = {"workload": "%s (NU %d, NI %d, NC %d), item_emb / item_b / icl row-sharded (id mod N)" % (WL_NAME, NU, NI, NC),
               "per_gpu_batch": B, "ms_per_step": ms, "value": B * ctx.world / (ms * 1e-3), "unit": "samples/s",
               "distinct_ids_per_rank": int(m.last_unique), "exchange_bytes_per_rank_step": int(2 * m.last_exchange_bytes),
               "item_shard_bytes_per_rank": int(m.n_local) * 33 * 4}
        del m, dbs
        return out
    finally:
        WL_NAME, NU, NI, NC = keep


# ---------------------------------------------------------------------------------------------------- main
def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--Ls", type=int, default=10)
    ap.add_argument("--resident", type=int, default=6, help="distinct device-resident batches")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--sustain-s", type=float, default=3.0, help="seconds of back-to-back steps in the sustained leg")
    ap.add_argument("--skip-extras", action="store_true", help="profiling runs: only the device-resident leg")
    ap.add_argument("--dp-parity", action="store_true", help="with --skip-extras at N > 1: still run the dp_parity check")
    ap.add_argument("--no-pipeline", action="store_true", help="do not presort the next batch behind the current step")
    ap.add_argument("--workload", default="electronics", choices=sorted(WORKLOADS))
    ap.add_argument("--strong", action="store_true", help="fixed GLOBAL batch: every rank gets batch / N rows")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global NU, NI, NC, WL_NAME
    WL_NAME, NU, NI, NC = WORKLOADS[args.workload]
    if args.strong:
        args.batch = max(1, args.batch // max(args.gpus, 1))
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return

    ctx = Ctx(args)
    torch, dist, rank, world = ctx.torch, ctx.dist, ctx.rank, ctx.world
    from tlsan_b200 import _lib
    lib = _lib.lib()
    B, L = args.batch, args.Ls
    clocks = ClockSampler(ctx.local_rank) if rank == 0 else None       # side process; started early
    model, cfg, icl = make_model(ctx, L)
    rng = np.random.default_rng(1234 + 1000 * rank)
    host_batches = synth_batches(rng, args.resident, B, L)
    dev_batches = [model.stage_batch(b) for b in host_batches]
    torch.cuda.synchronize()
    pipe = not args.no_pipeline          # name the next batch: its occurrence sort runs behind this step's backward
    Bg = B * world
    nres = len(dev_batches)

    # ---------------- device-resident timed region (the contract's K steps)
    train_loop(model, dev_batches, args.warmup, pipe, global_batch=Bg)
    ctx.barrier()
    launches0 = lib.tlsan_launch_count()
    _lib.check(lib.tlsan_profile_begin(args.steps))
    ctx.barrier()
    if clocks:
        clocks.wait_ready()
    t_w0 = time.time()
    ctx.e0.record()
    t_h0 = time.perf_counter()
    train_loop(model, dev_batches, args.steps, pipe, first=args.warmup, global_batch=Bg)
    t_enqueue = time.perf_counter() - t_h0             # host time to ENQUEUE the K steps (launch-bound if ~ device time)
    ctx.e1.record()
    ctx.barrier()
    t_w1 = time.time()
    ms = ctx.max_over_ranks(ctx.e0.elapsed_time(ctx.e1))
    phase = np.zeros((args.steps, len(_lib.PHASES)), np.float32)
    nrec = C.c_int32()
    _lib.check(lib.tlsan_profile_end(phase.ctypes.data, C.byref(nrec)))
    launches = lib.tlsan_launch_count() - launches0
    clk = clocks.window(t_w0, t_w1) if clocks else None
    loss = float(model._stats[0].item())
    ph = phase[:nrec.value].mean(axis=0)
    phases_ms = {n: float(v) for n, v in zip(_lib.PHASES, ph)}

    if args.skip_extras:
        short = {"ms_per_step": ms / args.steps, "phases_ms": phases_ms,
                 "host_enqueue_ms_per_step": 1e3 * t_enqueue / args.steps}
        if args.dp_parity and world > 1:          # profiling-style run that still checks multi-GPU correctness
            try:
                short["dp_parity"] = dp_parity(ctx)
            except Exception as e:
                short["dp_parity"] = {"error": repr(e)[:300]}
            ctx.barrier()
        if rank == 0:
            _emit(short)
        if clocks:
            clocks.stop()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- sustained: >= sustain_s seconds of back-to-back steps
    chunk, done = 500, 0
    ctx.barrier()
    t_s0 = time.time()
    ctx.e0.record()
    while True:
        train_loop(model, dev_batches, chunk, pipe, first=done, global_batch=Bg)
        done += chunk
        torch.cuda.synchronize()
        stop = torch.tensor([1.0 if time.time() - t_s0 >= args.sustain_s else 0.0], device="cuda")
        if world > 1:
            dist.all_reduce(stop, op=dist.ReduceOp.MAX)
        if float(stop.item()) > 0 or done >= 40000:
            break
    ctx.e1.record()
    ctx.barrier()
    t_s1 = time.time()
    ms_sus = ctx.max_over_ranks(ctx.e0.elapsed_time(ctx.e1)) / done
    clk_sus = clocks.window(t_s0, t_s1) if clocks else None
    sustained = {"seconds": t_s1 - t_s0, "steps": done, "ms_per_step": ms_sus, "value": Bg / (ms_sus * 1e-3),
                 "unit": "samples/s", "clocks": clk_sus}

    # ---------------- end to end through Model.train with host batches (batch k+1 staged while step k runs)
    e2e_steps = max(200, args.steps)

    def e2e_loop(batches, lazy):
        """K calls of Model.train(sess, batch, lr, prefetch=next batch); every step's loss is read back (D2H) inside the
        timed region -- right after the call (lazy=False), or after the NEXT call has been enqueued (lazy=True)."""
        for w in range(3):
            model.train(None, batches[w % nres], 1.0, global_batch=Bg)
        model.drop_prefetch()
        torch.cuda.synchronize()
        ctx.barrier()
        t0 = time.perf_counter()
        model.prefetch(batches[0])
        total, pending = 0.0, None
        for k in range(e2e_steps):
            cur = model.train(None, batches[k % nres], 1.0, global_batch=Bg, prefetch=batches[(k + 1) % nres], lazy_loss=lazy)
            if pending is not None:
                total += float(pending)
            pending = cur
        total += float(pending)
        torch.cuda.synchronize()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        model.drop_prefetch()
        assert np.isfinite(total)
        return dt

    e2e_s = e2e_loop(host_batches, False)
    h2d, d2h = model.last_h2d_bytes, model.last_d2h_bytes
    # the same loop fed with int32 batches (tlsan_b200.input index_dtype=np.int32): no 64 -> 32 bit narrowing pass
    hb32 = [tuple(np.ascontiguousarray(f, dtype=np.int32) if n not in (2, 5) else f for n, f in enumerate(b))
            for b in host_batches]
    e2e32_s = e2e_loop(hb32, False)
    del hb32
    # the packed feed: batches as DataInput(..., packed=True) yields them -- already in page-locked memory in the
    # staging layout -- so the timed region holds the H2D copy from pinned memory, the step and the loss read-back
    from tlsan_b200.input import PackedBatch
    pk = [PackedBatch.from_tuple(b) for b in host_batches]
    e2ep_sync_s = e2e_loop(pk, False)
    e2ep_s = e2e_loop(pk, True)
    h2d_p = model.last_h2d_bytes
    del pk

    # ---------------- epoch loop over a device-resident dataset: GPU batch assembly + train step
    from tlsan_b200.dataset import DeviceDataset
    from tlsan_b200.input import CsrDataset
    sl_all = np.concatenate([b[6] for b in host_batches]); ns_all = np.concatenate([b[7] for b in host_batches])
    pre_off = np.zeros(len(sl_all) + 1, np.int64); np.cumsum(sl_all, out=pre_off[1:])
    new_off = np.zeros(len(ns_all) + 1, np.int64); np.cumsum(ns_all, out=new_off[1:])
    mask_l = np.arange(L)[None, :] < sl_all[:, None]
    hi_all = np.concatenate([b[3] for b in host_batches]); ht_all = np.concatenate([b[5] for b in host_batches])
    hn_all = np.concatenate([np.pad(b[4], ((0, 0), (0, S_MAX - b[4].shape[1]))) for b in host_batches])
    mask_s = np.arange(S_MAX)[None, :] < ns_all[:, None]
    csr = CsrDataset(np.concatenate([b[0] for b in host_batches]), pre_off, hi_all[mask_l], ht_all[mask_l], new_off,
                     hn_all[mask_s], np.concatenate([b[1] for b in host_batches]),
                     np.concatenate([b[2] for b in host_batches]), np.concatenate([b[8] for b in host_batches]), False)
    dds = DeviceDataset(csr, is_test=False)
    perm = torch.randperm(len(dds), device="cuda", dtype=torch.int32)
    nb = len(dds) // B

    def ds_batch(k):
        return dds.batch(perm[(k % nb) * B:(k % nb + 1) * B], L, width="max")
    state = {"cur": ds_batch(0)}

    def ds_step(k):                      # batch k+1 is assembled (and, pipelined, sorted) while step k runs
        nxt = ds_batch(k + 1)
        model.train_staged(state["cur"], 1.0, global_batch=Bg, next_db=nxt if pipe else None)
        state["cur"] = nxt
    for k in range(6):                   # warm-up with the same call pattern (both workspaces get allocated here)
        ds_step(k)
    ds_steps = max(100, min(args.steps, 400))
    t_h0 = time.perf_counter()
    ms_ds = ctx.timed(ds_step, ds_steps)
    ds_wall = (time.perf_counter() - t_h0) / ds_steps

    # ---------------- scoring (eval_auc-style, 2 candidates) device-resident
    test_b = list(host_batches[0]); test_b[2] = host_batches[1][1]
    test_b = tuple(test_b)
    db = model.stage_batch(test_b, is_test=True)
    for _ in range(3):
        model.score_staged(db, 2)
    ms_score = ctx.timed(lambda k: model.score_staged(db, 2), 20)
    score_bytes = algorithmic_bytes(test_b, L)["scoring2"]

    # ---------------- full-catalogue ranking (eval_prec / eval_recall hot kernel, tcgen05 3xTF32 GEMM)
    rb = min(B, 65536)
    ut = torch.randn(rb, 64, device="cuda")
    lab = torch.randint(0, NI, (rb,), dtype=torch.int32, device="cuda")
    rk = torch.empty(rb, dtype=torch.int32, device="cuda")
    rdims = model._dims(rb, 1)
    need = C.c_size_t()
    _lib.check(lib.tlsan_rank_workspace_bytes(C.byref(rdims), C.byref(need)))
    rws = torch.empty(need.value, dtype=torch.uint8, device="cuda")

    def rank_once(k):
        _lib.check(lib.tlsan_label_rank_ws(C.byref(rdims), C.byref(model._params), ut.data_ptr(), lab.data_ptr(),
                                           rk.data_ptr(), rws.data_ptr(), rws.numel(), None))
    for _ in range(3):
        rank_once(0)
    ms_rank = ctx.timed(rank_once, 10)

    # ---------------- the other BASELINE configurations + multi-GPU correctness
    extras = {}
    for name, fn in (("movies", lambda: leg_other_workload(ctx, "movies", L, pipe, 60)),
                     ("scoring_sweep", lambda: leg_scoring_sweep(ctx)),
                     ("sharded_10M", lambda: leg_sharded(ctx)),
                     ("dp_parity", (lambda: dp_parity(ctx)) if world > 1 else None)):
        if fn is None:
            continue
        try:
            extras[name] = fn()
        except Exception as e:                                    # a failing extra leg is reported, the line still prints
            extras[name] = {"error": repr(e)[:300]}
        ctx.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline (SURVEY 8d byte model; CUDA events recorded by the library on the launching stream)
    per_batch = [algorithmic_bytes(b, L) for b in host_batches]
    used = [per_batch[(args.warmup + k) % len(per_batch)] for k in range(args.steps)]
    mean_bytes = {k: float(np.mean([u[k] for u in used])) for k in used[0]}
    train_bytes = mean_bytes["train"] + 2 * table_bytes(L, NU, NI, NC) + 2 * 4 * 4449
    step_ms = ms / args.steps
    # dominant kernel = the longest of the per-sample gather kernels (each phase below is ONE kernel; the forward
    # phase also holds the 12 us metadata pre-pass)
    pname = max(("long_fwd", "short", "bwd_long", "reduce"), key=lambda n: phases_ms[n])
    kname, kbytes, kms = _lib.PHASE_KERNEL[pname], mean_bytes[pname], phases_ms[pname]
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kname)
    except Exception:
        pass
    achieved = kbytes / (kms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": ctx.hbm, "unit": "GB/s",
                "frac": achieved / ctx.hbm, "traffic": traffic, "peak_source": ctx.hbm_src,
                "kernel_ms": kms, "algorithmic_bytes_per_launch": kbytes, "phases_ms": phases_ms,
                "note": "per-kernel bytes follow SURVEY 8d: what each kernel must read / write once; the backward's "
                        "re-read of the token rows, scratch and sort traffic are not credited",
                "per_kernel": {_lib.PHASE_KERNEL[n]: {"ms": phases_ms[n], "algorithmic_GBps": mean_bytes[n] / (phases_ms[n] * 1e-3) / 1e9,
                                                      "frac": mean_bytes[n] / (phases_ms[n] * 1e-3) / 1e9 / ctx.hbm}
                               for n in ("long_fwd", "short", "bwd_long", "reduce")},
                "step": {"algorithmic_bytes": train_bytes, "achieved": train_bytes / (step_ms * 1e-3) / 1e9,
                         "frac": train_bytes / (step_ms * 1e-3) / 1e9 / ctx.hbm,
                         "sustained_frac": train_bytes / (ms_sus * 1e-3) / 1e9 / ctx.hbm}}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample = cpu_port_throughput(L, args.cpu_budget)
        cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample}

    tf32_peak = float(ctx.peaks.get("bf16_tflops", 1590.0)) / 2.0
    issued = 3 * 2.0 * rb * NI * 72 / (ms_rank * 1e-3) / 1e12
    algo = 2.0 * rb * NI * 64 / (ms_rank * 1e-3) / 1e12
    line = {
        "metric": "train_samples_per_s", "value": args.steps * Bg / (ms * 1e-3), "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, B),
        "e2e": {"value": e2e_steps * Bg / e2ep_s, "unit": "samples/s", "h2d_bytes_per_step": h2d_p,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": 1e3 * e2ep_s / e2e_steps,
                "how": "Model.train(sess, batch, lr, prefetch=next batch, lazy_loss=True) with the batches "
                       "tlsan_b200.input.DataInput(..., packed=True) yields: PackedBatch = the reference 9-tuple's fields "
                       "in ONE page-locked int32 buffer in the staging layout; per step: H2D from pinned memory (batch "
                       "k+1's copy and occurrence sort overlap step k) + session expansion + train step + 4-byte loss "
                       "read-back, every loss read by the host one call later (double-buffered loop)",
                "sync_loss": {"value": e2e_steps * Bg / e2ep_sync_s, "ms_per_step": 1e3 * e2ep_sync_s / e2e_steps,
                              "what": "same packed feed, loss read synchronously inside every call (the reference's "
                                      "`loss = model.train(...)`): the GPU idles while the host enqueues the next step"},
                "tuple_int64_feed": {"value": e2e_steps * Bg / e2e_s, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                                     "h2d_bytes_per_step": h2d,
                                     "what": "synchronous loop fed with the reference's own 9-tuples of int64 numpy arrays "
                                             "(TLSAN/input.py:54): adds the multi-threaded int64->int32 cast + range "
                                             "checks + pack into pinned memory per step (host-memory-bound)"},
                "int32_feed": {"value": e2e_steps * Bg / e2e32_s, "ms_per_step": 1e3 * e2e32_s / e2e_steps,
                               "what": "9-tuples whose integer fields are already int32 (index_dtype=np.int32)"}},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clk,
        "sustained": sustained,
        "host_enqueue_ms_per_step": 1e3 * t_enqueue / args.steps,
        "dataset_resident": {"metric": "train_samples_per_s", "value": Bg / (ms_ds * 1e-3), "unit": "samples/s",
                             "ms_per_step": ms_ds, "host_wall_ms_per_step": 1e3 * ds_wall, "steps": ds_steps,
                             "what": "shuffled epoch loop over a CSR dataset resident in HBM: tlsan_collate (GPU batch "
                                     "assembly in the input.py layout) + train step, no host batcher; when "
                                     "host_wall_ms_per_step ~ ms_per_step the loop is bound by the host enqueueing "
                                     "~30 launches per step, not by the GPU"},
        "eval": {"metric": "eval_seqs_per_s", "value": Bg / (ms_score * 1e-3), "unit": "seqs/s", "candidates": 2,
                 "ms": ms_score,
                 "roofline": {"bound": "hbm", "achieved": score_bytes / (ms_score * 1e-3) / 1e9, "peak": ctx.hbm,
                              "unit": "GB/s", "frac": score_bytes / (ms_score * 1e-3) / 1e9 / ctx.hbm,
                              "algorithmic_bytes_per_seq": score_bytes / B}},
        "eval_rank": {"metric": "full_catalogue_rank_seqs_per_s", "value": rb * world / (ms_rank * 1e-3), "unit": "seqs/s",
                      "what": "eval_prec/eval_recall hot kernel: [B,64]x[64,NI] 3xTF32 tcgen05 GEMM + count-only "
                              "epilogue (k_build_catalogue + k_label_rank_tc), B=%d, NI=%d" % (rb, NI),
                      "roofline": {"bound": "tensor", "achieved": algo, "peak": tf32_peak, "unit": "TFLOP/s",
                                   "frac": algo / tf32_peak, "issued_TFLOPs": issued, "issued_frac": issued / tf32_peak,
                                   "note": "achieved = ALGORITHMIC flops 2*B*NI*64; issued = 3 tf32 terms x K padded "
                                           "64->72; peak = measured bf16 cuBLAS burst / 2 (tf32 runs at half the bf16 rate)"}},
        "final_loss": loss,
    }
    line.update(extras)
    _emit(line)
    if clocks:
        clocks.stop()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
