#!/usr/bin/env python
"""The kernels of the peer-memory exchange (tlsan_dp_exchange) on ONE GPU with world = 1: every wait is trivially
satisfied and every "peer" read is local, so an `ncu --metrics gpu__time_duration.sum` launch list of this script
gives the kernels' own durations without NVLink or rank skew (profiles/r02_dp_exchange_kernels.txt).

  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file out.csv python tools/dp_exchange_1gpu.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tlsan_b200 import synth  # noqa: E402
from tlsan_b200.model import Model  # noqa: E402
from tlsan_b200._lib import check  # noqa: E402


def main():
    name, NU, NI, NC = synth.WORKLOADS["electronics"]
    L, B = 10, 65536
    rng = np.random.default_rng(0)
    cfg = dict(hidden_units=64, num_blocks=1, num_heads=8, Ls=L, dropout=0.0, regulation_rate=0.00005,
               itemid_embedding_size=32, userid_embedding_size=32, cateid_embedding_size=32, optimizer="sgd",
               learning_rate=1.0, max_gradient_norm=5.0, train_batch_size=32, test_batch_size=128,
               model_dir="/tmp/tlsan_dpx", user_count=NU, item_count=NI, cate_count=NC)
    icl = rng.integers(0, NC, NI).astype(np.int32)
    m = Model(cfg, icl)
    batches = synth.synth_batches(rng, 2, B, L, NU, NI, NC)
    dbs = [m.stage_batch(b) for b in batches]
    dims = m._dims(B, dbs[0].S, B)
    need = C.c_size_t()
    check(m._lib.tlsan_dp_arena_bytes(C.byref(dims), 1, C.byref(need)))
    mine, handle = C.c_void_p(), C.create_string_buffer(64)
    check(m._lib.tlsan_dp_arena_create(need.value, C.byref(mine), handle))
    ptrs = (C.c_void_p * 1)(mine.value)
    st = m._stream()
    for k in range(6):
        db = dbs[k % 2]
        dims = m._dims(db.B, db.S, db.B)
        ws = m._workspace(dims, 0)
        check(m._lib.tlsan_step_grads_pipelined(C.byref(dims), C.byref(m._params), C.byref(db.c), None, ws.data_ptr(),
                                                ws.numel(), mine, st))
        check(m._lib.tlsan_dp_exchange(C.byref(dims), C.byref(m._params), ptrs, 0, 1, k + 1, 1.0, m.reg, m.clip,
                                       ws.data_ptr(), ws.numel(), m._stats.data_ptr(), st))
    torch.cuda.synchronize()
    print("loss", float(m._stats[0].item()))


if __name__ == "__main__":
    main()
