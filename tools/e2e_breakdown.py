"""Where does Model.train(host batch) spend its time?  (pack / H2D / step / sync)  -- diagnostic, not a bench."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import tlsan_oracle as O
from tlsan_b200.model import Model, pack_batch, _pack_offsets

B, L = 65536, 10
cfg = O.default_config(bench.NU, bench.NI, bench.NC, Ls=L)
icl = np.random.default_rng(1234).integers(0, bench.NC, bench.NI).astype(np.int32)
m = Model(cfg, icl)
hb = bench.synth_batches(np.random.default_rng(1), 4, B, L)
print("cpu_count", os.cpu_count())
for k in range(3):
    m.train(None, hb[k % 4], 1.0)
def t(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n): fn(k)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
S = hb[0][4].shape[1]
offs, total = _pack_offsets(B, L, S)
host = torch.empty(total, dtype=torch.int32).pin_memory()
dims = m._dims(B, S)
print("pack ms", t(lambda k: pack_batch(m._lib, hb[0], dims, False, host.numpy(), True)))
dev = torch.empty(total, dtype=torch.int32, device="cuda")
print("h2d ms", t(lambda k: dev.copy_(host, non_blocking=True)), "MB", total * 4 / 1e6)
dbs = [m.stage_batch(b) for b in hb]
print("step ms", t(lambda k: m.train_staged(dbs[k % 4], 1.0)))
print("stage ms", t(lambda k: m.stage_batch(hb[k % 4])))
print("train ms", t(lambda k: m.train(None, hb[k % 4], 1.0)))

# double-buffered feed: where the wall time of one call goes
import time as _t
m.prefetch(hb[0])
acc = np.zeros(4)
torch.cuda.synchronize()
N = 50
t00 = _t.perf_counter()
for k in range(N):
    t0 = _t.perf_counter()
    db, slot = m._staged(hb[k % 4], False)
    t1 = _t.perf_counter()
    st = m.train_staged(db, 1.0)
    if slot is not None:
        slot[1] = torch.cuda.Event(); slot[1].record()
    t2 = _t.perf_counter()
    m.prefetch(hb[(k + 1) % 4])
    t3 = _t.perf_counter()
    st[0].item()
    t4 = _t.perf_counter()
    acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3]
print("prefetch loop ms/step", (_t.perf_counter() - t00) / N * 1e3, "take/enqueue/prefetch/item ms", acc / N * 1e3)
hb32 = [tuple(np.ascontiguousarray(f, dtype=np.int32) if n != 5 and n != 2 else f for n, f in enumerate(b)) for b in hb]
print("train int32 feed ms", t(lambda k: m.train(None, hb32[k % 4], 1.0, prefetch=hb32[(k + 1) % 4]), 50))
print("train int64 feed ms", t(lambda k: m.train(None, hb[k % 4], 1.0, prefetch=hb[(k + 1) % 4]), 50))
