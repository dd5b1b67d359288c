#!/usr/bin/env python
"""bench.py -- TLSAN train-step throughput on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--batch B] [--Ls L] [--no-cpu-baseline] [--workload electronics|movies] [--strong]

--workload movies = BASELINE.json configs[2] (Movies-TV shape: NU 35 896, NI 28 589, NC 15); --strong keeps the
GLOBAL batch at --batch and gives every rank batch/N rows (strong scaling); the default is weak scaling.

Workload (BASELINE.json configs[1]): "TLSAN Electronics-shape synthetic (40k users, 22k items,
673 cates) fp32": NU 39 991, NI 22 048, NC 673, generators of SURVEY.md section 8d config 2
(numpy default_rng(1234 + rank)), per-GPU batch 65 536, Ls 10.  A step = one Model.train step
(forward, loss, backward, segmented reduce, L2 + clip + SGD over every table row).

  value : train samples/s, all ranks, device-resident batches (several distinct batches cycled)
  e2e   : the same step through Model.train(sess, batch, lr) with HOST numpy batches
          (pack -> pinned -> H2D -> step -> loss D2H inside the timed region)
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement".
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

NU, NI, NC = 39991, 22048, 673
N_SAMPLES = 561100
WORKLOADS = {"electronics": ("TLSAN Electronics-shape synthetic", 39991, 22048, 673),
             "movies": ("TLSAN Movies-TV-shape synthetic", 35896, 28589, 15)}
WL_NAME = WORKLOADS["electronics"][0]
# Digital-Music empirical laws: P(min(len,10) = k), k = 1..10 ; short length pmf
P_LONG = np.array([7.0, 6.9, 6.7, 6.5, 6.3, 5.9, 5.2, 4.5, 3.9, 47.1]) / 100.0
P_SHORT_HEAD = np.array([.8724, .0886, .0223, .0085, .0040])
S_MAX = 18


def synth_batches(rng, n_batches, B, L):
    """Electronics-shape synthetic batches in the TLSAN/input.py layout."""
    p_long = P_LONG / P_LONG.sum()
    tail = np.full(S_MAX - 5, (1.0 - P_SHORT_HEAD.sum()) / (S_MAX - 5))
    p_short = np.concatenate([P_SHORT_HEAD, tail])
    p_short /= p_short.sum()
    out = []
    for _ in range(n_batches):
        frac = rng.choice(10, B, p=p_long) + 1                           # law of min(len, 10)
        sl = np.maximum(1, np.round(frac * (L / 10.0))).astype(np.int64) if L != 10 else frac.astype(np.int64)
        new_sl = (rng.choice(S_MAX, B, p=p_short) + 1).astype(np.int64)
        S = int(new_sl.max())
        hist_i = rng.integers(0, NI, (B, L)).astype(np.int64)
        hist_i_new = rng.integers(0, NI, (B, S)).astype(np.int64)
        n = np.sort(rng.integers(1, 13, (B, L)), axis=1)[:, ::-1]        # bucket non-increasing in t
        hist_t = (1.0 / n).astype(np.float32)
        col = np.arange(L)[None, :]
        hist_i[col >= sl[:, None]] = 0
        hist_t[col >= sl[:, None]] = 0
        hist_i_new[np.arange(S)[None, :] >= new_sl[:, None]] = 0
        out.append((rng.integers(0, NU, B).astype(np.int64), rng.integers(0, NI, B).astype(np.int64),
                    rng.integers(0, 2, B).astype(np.int64), hist_i, hist_i_new, hist_t, sl, new_sl,
                    rng.integers(0, NC, B).astype(np.int64)))
    return out


def algorithmic_bytes(batch, L):
    """SURVEY.md 8d byte model, evaluated on the actual lengths of `batch` (totals per batch).
    Per sample: R = 2(l+s)+4 embedding rows of 128 B.  The per-kernel figures split the train-step
    formula by which kernel touches what; scratch traffic is never credited."""
    sl = np.asarray(batch[6], np.int64); s = np.asarray(batch[7], np.int64)
    S = batch[4].shape[1]
    R = 2 * (sl + s) + 4
    scoring1 = 4 * (2 * L + S + 6) + 4 * (sl + s + 1) + 128 * R + (4 * L + 4) + 4
    train = scoring1 + 2 * (128 * R + 4 * sl + 4) + 4
    long_fwd = 4 * (2 * L + 2) + 4 * sl + 128 * 2 * sl + 4 * L                 # ids, hist_t, icl, rows, usert row
    short = 4 * (S + 6) + 4 * (s + 1) + 128 * (2 * s + 4) + 4 + 128 * (2 * s + 4) + 8   # reads + gradient rows written
    bwd_long = 4 * 2 * L + 4 * sl + 128 * 2 * sl + 4 * L + 128 * 2 * sl + 4 * sl
    reduce_ = 128 * R + 4 * sl + 4                                              # every gradient row read once
    return {"scoring1": int(scoring1.sum()), "train": int(train.sum()), "long_fwd": int(long_fwd.sum()),
            "short": int(short.sum()), "bwd_long": int(bwd_long.sum()), "reduce": int(reduce_.sum())}


def table_bytes(L):
    return 4 * (33 * NI + 32 * NU + L * NU + 32 * NC)


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
    print(time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), bits, flush=True)
    time.sleep(0.002)
"""


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region: a side process polls NVML every
    ~2 ms (started early; samples are filtered to the [mark_begin, mark_end] window)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.proc, self.t0, self.t1 = None, None, None
        try:
            import torch
            p = torch.cuda.get_device_properties(index)
            bus = "%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, bus, str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.lines = []

            def pump():
                for ln in self.proc.stdout:
                    self.lines.append(ln)
            self.t = threading.Thread(target=pump, daemon=True)
            self.t.start()
        except Exception as e:                      # pragma: no cover
            self.err = repr(e)

    def mark_begin(self):
        t_end = time.time() + 5.0                  # the side process needs ~1 s to import + nvmlInit
        while self.proc is not None and not self.lines and time.time() < t_end:
            time.sleep(0.01)
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml sampler unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=1)
        sm, bits, mx = [], 0, None
        for ln in self.lines:
            f = ln.split()
            try:
                if f[0] == "max":
                    mx = int(f[1])
                elif self.t0 <= float(f[0]) <= self.t1:
                    sm.append(int(f[1])); bits |= int(f[2])
            except Exception:
                pass
        reasons = sorted(name for bit, name in self.REASONS.items() if bits & bit)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": reasons}


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
    return n * B / dt, torch.get_num_threads(), "%d steps of B=%d, L=%d, Electronics-shape synthetic (%.1f s)" % (
        n, B, L, dt)


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
    sample = "each step = B=%d rows of the Electronics-shape workload (bounded sample of the 65536-row step)" % B
    _emit({
        "impl": "reference", "metric": "train_samples_per_s", "value": v, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, B),
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
    ap.add_argument("--skip-extras", action="store_true", help="profiling runs: no e2e / scoring / cpu legs")
    ap.add_argument("--no-pipeline", action="store_true", help="do not presort the next batch behind the current step")
    ap.add_argument("--workload", default="electronics", choices=sorted(WORKLOADS))
    ap.add_argument("--strong", action="store_true", help="fixed GLOBAL batch: every rank gets batch / N rows")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global NU, NI, NC, WL_NAME
    WL_NAME, NU, NI, NC = WORKLOADS[args.workload]
    if args.strong:
        args.batch = max(1, args.batch // max(args.gpus, 1))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from tlsan_b200 import _lib
    from tlsan_b200.model import Model
    from oracle import tlsan_oracle as O            # config defaults + cpu_baseline leg only

    torch.cuda.set_device(local_rank)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        pg = dist.group.WORLD
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus

    B, L = args.batch, args.Ls
    cfg = O.default_config(NU, NI, NC, Ls=L)
    icl = np.random.default_rng(1234).integers(0, NC, NI).astype(np.int32)     # same on every rank
    clocks = ClockSampler(local_rank) if rank == 0 else None       # side process; started early
    model = Model(cfg, icl, seed=1234, process_group=pg)
    rng = np.random.default_rng(1234 + 1000 * rank)
    host_batches = synth_batches(rng, args.resident, B, L)
    dev_batches = [model.stage_batch(b) for b in host_batches]
    torch.cuda.synchronize()
    lib = _lib.lib()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timed region
    nres = len(dev_batches)
    pipe = not args.no_pipeline          # name the next batch: its occurrence sort runs behind this step's backward
    for w in range(args.warmup):
        model.train_staged(dev_batches[w % nres], 1.0, next_db=dev_batches[(w + 1) % nres] if pipe else None)
    barrier()
    launches0 = lib.tlsan_launch_count()
    _lib.check(lib.tlsan_profile_begin(args.steps))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if clocks:
        clocks.mark_begin()
    e0.record()
    for k in range(args.steps):
        model.train_staged(dev_batches[(args.warmup + k) % nres], 1.0,
                           next_db=dev_batches[(args.warmup + k + 1) % nres] if pipe else None)
    e1.record()
    barrier()
    if clocks:
        clocks.mark_end()
    ms = e0.elapsed_time(e1)
    phase = np.zeros((args.steps, len(_lib.PHASES)), np.float32)
    nrec = C.c_int32()
    _lib.check(lib.tlsan_profile_end(phase.ctypes.data, C.byref(nrec)))
    launches = lib.tlsan_launch_count() - launches0
    clk = clocks.stop() if clocks else None
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    loss = float(model._stats[0].item())

    # ---------------- end to end through Model.train with host batches
    if args.skip_extras:
        if rank == 0:
            _emit({"ms_per_step": ms / args.steps, "phases_ms": dict(zip(_lib.PHASES, map(float, phase[:nrec.value].mean(axis=0))))})
        return
    for w in range(2):
        model.train(None, host_batches[w % len(host_batches)], 1.0)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        model.train(None, host_batches[k % len(host_batches)], 1.0)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e2e.item())
    h2d, d2h = model.last_h2d_bytes, model.last_d2h_bytes

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
    cur = ds_batch(0)
    for k in range(4):                   # warm-up with the same call pattern (both workspaces get allocated here)
        nxt = ds_batch(k + 1)
        model.train_staged(cur, 1.0, next_db=nxt if pipe else None)
        cur = nxt
    barrier()
    ds_steps = max(3, min(args.steps, 50))
    cur = ds_batch(0)
    barrier()
    e0.record()
    for k in range(ds_steps):            # batch k+1 is assembled (and, pipelined, sorted) while step k runs
        nxt = ds_batch(k + 1)
        model.train_staged(cur, 1.0, next_db=nxt if pipe else None)
        cur = nxt
    e1.record()
    barrier()
    t_ds = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t_ds, op=dist.ReduceOp.MAX)
    ms_ds = float(t_ds.item()) / ds_steps

    # ---------------- scoring (eval_auc-style, 2 candidates) device-resident
    test_b = list(host_batches[0]); test_b[2] = host_batches[1][1]
    db = model.stage_batch(tuple(test_b), is_test=True)
    for _ in range(3):
        model.score_staged(db, 2)
    barrier()
    e0.record()
    for _ in range(10):
        model.score_staged(db, 2)
    e1.record()
    barrier()
    ms_score = e0.elapsed_time(e1) / 10

    # ---------------- full-catalogue ranking (eval_prec / eval_recall hot kernel, tcgen05 3xTF32 GEMM)
    rb = min(B, 65536)
    ut = torch.randn(rb, 64, device="cuda")
    lab = torch.randint(0, NI, (rb,), dtype=torch.int32, device="cuda")
    rk = torch.empty(rb, dtype=torch.int32, device="cuda")
    rdims = model._dims(rb, 1)
    need = C.c_size_t()
    _lib.check(lib.tlsan_rank_workspace_bytes(C.byref(rdims), C.byref(need)))
    rws = torch.empty(need.value, dtype=torch.uint8, device="cuda")

    def rank_once():
        _lib.check(lib.tlsan_label_rank_ws(C.byref(rdims), C.byref(model._params), ut.data_ptr(), lab.data_ptr(),
                                           rk.data_ptr(), rws.data_ptr(), rws.numel(), None))
    for _ in range(3):
        rank_once()
    barrier()
    e0.record()
    for _ in range(10):
        rank_once()
    e1.record()
    barrier()
    ms_rank = e0.elapsed_time(e1) / 10

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (CUDA events recorded around it on its stream)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    ph = phase[:nrec.value].mean(axis=0)
    phases_ms = {n: float(v) for n, v in zip(_lib.PHASES, ph)}
    per_batch = [algorithmic_bytes(b, L) for b in host_batches]
    used = [per_batch[(args.warmup + k) % len(per_batch)] for k in range(args.steps)]
    mean_bytes = {k: float(np.mean([u[k] for u in used])) for k in used[0]}
    train_bytes = mean_bytes["train"] + 2 * table_bytes(L) + 2 * 4 * 4449
    # dominant kernel = the longest of the per-sample gather kernels (each phase below is ONE kernel)
    pname = max(("long_fwd", "short", "bwd_long", "reduce"), key=lambda n: phases_ms[n])
    kname, kbytes, kms = _lib.PHASE_KERNEL[pname], mean_bytes[pname], phases_ms[pname]
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kname)
    except Exception:
        pass
    achieved = kbytes / (kms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": kms, "algorithmic_bytes_per_launch": kbytes, "phases_ms": phases_ms,
                "per_kernel": {_lib.PHASE_KERNEL[n]: {"ms": phases_ms[n], "algorithmic_GBps": mean_bytes[n] / (phases_ms[n] * 1e-3) / 1e9,
                                                      "frac": mean_bytes[n] / (phases_ms[n] * 1e-3) / 1e9 / peak}
                               for n in ("long_fwd", "short", "bwd_long", "reduce")},
                "step": {"algorithmic_bytes": train_bytes,
                         "achieved": train_bytes / (ms / args.steps * 1e-3) / 1e9,
                         "frac": train_bytes / (ms / args.steps * 1e-3) / 1e9 / peak}}

    cpu = None
    if not args.no_cpu_baseline:
        v, cores, sample = cpu_port_throughput(L, args.cpu_budget)
        cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample}

    line = {
        "metric": "train_samples_per_s", "value": args.steps * B * world / (ms * 1e-3), "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, B),
        "e2e": {"value": e2e_steps * B * world / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clk,
        "dataset_resident": {"metric": "train_samples_per_s", "value": B * world / (ms_ds * 1e-3), "unit": "samples/s",
                             "what": "shuffled epoch loop over a CSR dataset resident in HBM: tlsan_collate (GPU batch "
                                     "assembly in the input.py layout) + train step, no host batcher", "steps": ds_steps},
        "eval": {"metric": "eval_seqs_per_s", "value": B * world / (ms_score * 1e-3), "unit": "seqs/s",
                 "candidates": 2},
        "eval_rank": {"metric": "full_catalogue_rank_seqs_per_s", "value": rb * world / (ms_rank * 1e-3), "unit": "seqs/s",
                      "what": "eval_prec/eval_recall hot kernel: [B,64]x[64,NI] 3xTF32 tcgen05 GEMM + count-only "
                              "epilogue (k_build_catalogue + k_label_rank_tc), B=%d, NI=%d" % (rb, NI),
                      "roofline": {"bound": "tensor", "achieved": 3 * 2.0 * rb * NI * 72 / (ms_rank * 1e-3) / 1e12,
                                   "peak": float(peaks.get("bf16_tflops", 1590.0)) / 2.0, "unit": "TFLOP/s",
                                   "frac": 3 * 2.0 * rb * NI * 72 / (ms_rank * 1e-3) / 1e12 /
                                           (float(peaks.get("bf16_tflops", 1590.0)) / 2.0),
                                   "note": "issued tf32 flops (3 terms, K padded 64->72); peak = measured bf16 "
                                           "cuBLAS burst / 2 (tf32 runs at half the bf16 rate)"}},
        "final_loss": loss,
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
