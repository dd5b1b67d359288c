"""Shared helpers for the test-suite (loads fixtures, builds synthetic batches)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class DigitalMusic:
    pass


def load_digital_music():
    from oracle.make_golden import from_csr
    d = np.load(os.path.join(GOLD, "digital_music.npz"))
    out = DigitalMusic()
    out.raw = d
    out.counts = tuple(int(x) for x in d["counts"])
    out.icl = d["icl"]
    out.train_set = from_csr(d, "train_", False)
    out.test_set = from_csr(d, "test_", True)
    return out


def synth_batch(rng, B, L, S, NI, NU, NC, is_test=False, full=False, dup_items=False):
    """Random batch in the input.py layout (empirical Digital-Music length law unless full)."""
    if full:
        sl = np.full(B, L, np.int64)
    else:
        sl = np.minimum(rng.integers(1, 2 * L + 1, B), L).astype(np.int64)
    new_sl = np.minimum(rng.geometric(0.8, B), S).astype(np.int64)
    new_sl[rng.integers(0, B)] = S
    hi_items = NI if not dup_items else min(NI, 7)
    hist_i = rng.integers(0, hi_items, (B, L)).astype(np.int64)
    hist_i_new = rng.integers(0, hi_items, (B, S)).astype(np.int64)
    n = np.sort(rng.integers(1, 13, (B, L)), axis=1)[:, ::-1]            # non-increasing buckets
    hist_t = (1.0 / n).astype(np.float32)
    col = np.arange(L)[None, :]
    hist_i[col >= sl[:, None]] = 0
    hist_t[col >= sl[:, None]] = 0
    hist_i_new[np.arange(S)[None, :] >= new_sl[:, None]] = 0
    u = rng.integers(0, NU, B).astype(np.int64)
    c = rng.integers(0, NC, B).astype(np.int64)
    i = rng.integers(0, NI, B).astype(np.int64)
    if is_test:
        second = rng.integers(0, NI, B).astype(np.int64)
    else:
        second = rng.integers(0, 2, B).astype(np.int64)
    return (u, i, second, hist_i, hist_i_new, hist_t, sl, new_sl, c)


def model_from_params(params, icl, config, **kw):
    """tlsan_b200.Model carrying exactly the oracle's weights."""
    import torch
    from tlsan_b200.model import Model
    m = Model(config, icl, **kw)
    m.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in params.items()})
    return m


def rel_err(a, b, floor=1e-6, tol=1e-4):
    """ELEMENT-WISE error in units of the tolerance `|a - b| <= tol * |b| + floor` (north star: 1e-4 relative in
    fp32; the absolute floor covers values next to zero): returns max_i |a_i - b_i| / (|b_i| + floor / tol), so
    `rel_err(a, b) < tol` holds exactly when every element is inside its own tolerance."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor / tol)))


def assert_step_matches(got_sd, old_params, ref_new, lr, tol=1e-4):
    """Updated weights of one train step against the fp64 oracle, ELEMENT-WISE on the step lr * grad:
    |got - ref| <= tol * |step_i| + 1e-5 * max|step| + 2e-7 * |w_i|   for every element i of every variable.
    (Relative term per element; the second term is the floor for gradient elements that cancel to ~0, a tenth of
    the tolerance on the tensor's largest step; the third is the fp32 rounding of w - lr * g itself.)"""
    worst = {}
    smax = {k: float(np.max(np.abs(np.asarray(old_params[k], np.float64) - np.asarray(v, np.float64))))
            for k, v in ref_new.items()}
    for k, v in ref_new.items():
        got = np.asarray(got_sd[k], np.float64)
        v = np.asarray(v, np.float64)
        step = np.asarray(old_params[k], np.float64) - v
        # the gradient of the second map's bias is identically 0 (softmax over the sequence is shift invariant,
        # model.py:383-386): what the kernels return there is the rounding noise of terms as large as those of the
        # sibling kernel's gradient, so that tensor's largest step sets the floor
        sib = k.replace("/bias", "/W")
        scale = max(smax[k], smax.get(sib, 0.0)) if k.endswith("bn_dense_map2/linear_map/bias") else smax[k]
        bound = tol * np.abs(step) + 1e-5 * scale + 2e-7 * np.abs(v) + 1e-12
        ratio = np.abs(got - v) / bound
        worst[k] = float(np.max(ratio))
        assert worst[k] <= 1.0, (k, worst[k], float(np.max(np.abs(got - v))), float(np.max(np.abs(step))))
    return worst
