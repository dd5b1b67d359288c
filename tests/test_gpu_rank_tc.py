"""Full-catalogue label rank on the tcgen05 tensor cores (csrc/tlsan_rank_tc.cu) against the fp64
oracle ranks (reference model.py:140-156 semantics: top_k order, ties -> lower index) and against the CUDA-core
kernel, on ragged shapes: B and NI not multiples of the 128 x 128 tile, one and several item ranges per user tile,
duplicated catalogue rows (exact score ties)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import tlsan_oracle as O

pytestmark = pytest.mark.gpu


def _ranks(model, ut, label, impl):
    from tlsan_b200 import _lib
    lib = _lib.lib()
    B = ut.shape[0]
    dims = model._dims(B, 1)
    rank = torch.full((B,), -7, dtype=torch.int32, device="cuda")
    if impl == "ffma":
        _lib.check(lib.tlsan_label_rank(C.byref(dims), C.byref(model._params), ut.data_ptr(), label.data_ptr(),
                                        rank.data_ptr(), None))
    else:
        need = C.c_size_t()
        _lib.check(lib.tlsan_rank_workspace_bytes(C.byref(dims), C.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
        _lib.check(lib.tlsan_label_rank_ws(C.byref(dims), C.byref(model._params), ut.data_ptr(), label.data_ptr(),
                                           rank.data_ptr(), ws.data_ptr(), ws.numel(), None))
    torch.cuda.synchronize()
    return rank.cpu().numpy()


@pytest.mark.parametrize("B,NI,NC", [(1, 1, 1), (7, 100, 3), (128, 128, 5), (129, 129, 5), (300, 1583, 53),
                                     (1000, 5000, 40), (2048, 22048, 673)])
def test_rank_tc_matches_oracle_and_ffma(B, NI, NC):
    from tlsan_b200.model import Model
    rng = np.random.default_rng(B * 31 + NI)
    cfg = O.default_config(50, NI, NC, Ls=10)
    icl = rng.integers(0, NC, NI).astype(np.int32)
    model = Model(cfg, icl, seed=3)
    model.item_b.copy_(torch.from_numpy(rng.normal(0, 0.3, NI).astype(np.float32)))
    ut = torch.from_numpy(rng.normal(0, 1.0, (B, 64)).astype(np.float32)).cuda()
    label = torch.from_numpy(rng.integers(0, NI, B).astype(np.int32)).cuda()
    all_emb = torch.cat([model.item_emb, model.cate_emb[torch.from_numpy(icl.astype(np.int64)).cuda()]], 1)
    scores = (ut.double() @ all_emb.double().T + model.item_b.double()[None, :]).cpu().numpy()
    ref = O.label_ranks(scores, label.cpu().numpy())
    tc = _ranks(model, ut, label, "tc")
    ff = _ranks(model, ut, label, "ffma")
    # fp32-level arithmetic against fp64: a rank may move only where two scores differ by rounding noise
    s_lab = scores[np.arange(B), label.cpu().numpy()][:, None]
    near = (np.abs(scores - s_lab) < 1e-5 * (1 + np.abs(s_lab))).sum(1) - 1
    assert np.all(np.abs(tc - ref) <= near), (tc - ref)[np.abs(tc - ref) > near]
    assert np.all(np.abs(ff - ref) <= near)
    assert (tc == ref).mean() > 0.99 and tc.min() >= 0 and tc.max() < NI


def test_rank_tc_breaks_exact_ties_by_index():
    """Items with identical rows and bias score identically in every arithmetic: the order must be by index."""
    from tlsan_b200.model import Model
    rng = np.random.default_rng(5)
    NI, NC, B = 700, 4, 260
    cfg = O.default_config(10, NI, NC, Ls=10)
    icl = np.zeros(NI, np.int32)
    model = Model(cfg, icl, seed=1)
    model.item_emb.copy_(model.item_emb[:1].expand(NI, 32).clone())       # every item identical
    ut = torch.from_numpy(rng.normal(0, 1.0, (B, 64)).astype(np.float32)).cuda()
    label = torch.from_numpy(rng.integers(0, NI, B).astype(np.int32)).cuda()
    tc = _ranks(model, ut, label, "tc")
    assert np.array_equal(tc, label.cpu().numpy())                         # exactly the lower-index items are ahead
