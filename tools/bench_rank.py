"""eval_prec / eval_recall hot kernel: full-catalogue label rank, tcgen05 vs CUDA-core formulation (diagnostic)."""
import ctypes as C, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import tlsan_oracle as O
from tlsan_b200 import _lib
from tlsan_b200.model import Model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
NI = int(sys.argv[2]) if len(sys.argv) > 2 else bench.NI
cfg = O.default_config(bench.NU, NI, bench.NC, Ls=10)
rng = np.random.default_rng(0)
icl = rng.integers(0, bench.NC, NI).astype(np.int32)
m = Model(cfg, icl)
lib = _lib.lib()
ut = torch.randn(B, 64, device="cuda")
label = torch.randint(0, NI, (B,), dtype=torch.int32, device="cuda")
rank = torch.empty(B, dtype=torch.int32, device="cuda")
dims = m._dims(B, 1)
need = C.c_size_t(); _lib.check(lib.tlsan_rank_workspace_bytes(C.byref(dims), C.byref(need)))
ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
def tc():
    _lib.check(lib.tlsan_label_rank_ws(C.byref(dims), C.byref(m._params), ut.data_ptr(), label.data_ptr(), rank.data_ptr(), ws.data_ptr(), ws.numel(), None))
def ff():
    _lib.check(lib.tlsan_label_rank(C.byref(dims), C.byref(m._params), ut.data_ptr(), label.data_ptr(), rank.data_ptr(), None))
out = {"B": B, "NI": NI}
for name, fn, n in (("tcgen05", tc, 10), ("ffma", ff, 2)):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out[name] = {"ms": ms, "useful_TFLOPs": 2.0 * B * NI * 64 / ms / 1e9, "issued_tf32_TFLOPs": (3 * 2.0 * B * NI * 72 / ms / 1e9) if name == "tcgen05" else None,
                 "seqs_per_s": B / ms * 1e3}
    out[name + "_ranks_sum"] = int(rank.sum().item())
print(json.dumps(out))
