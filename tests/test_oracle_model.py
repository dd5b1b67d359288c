"""The oracle restatement itself: committed pins, internal consistency of its backward, and the
TF-internal choices (clip norm definitions) documented in oracle/tlsan_oracle.py."""
import os

import numpy as np
import pytest
import torch

from oracle import tlsan_oracle as O
from tests.util import GOLD, synth_batch


def _setup(dm):
    cfg = O.default_config(*dm.counts)
    return cfg, O.randomize_params(O.init_params(cfg, seed=1234), seed=7)


def test_fp32_oracle_reproduces_committed_fp64_pins(dm):
    g = np.load(os.path.join(GOLD, "model_golden.npz"))
    cfg, params = _setup(dm)
    r = O.train_step(params, dm.icl, O.collate_train(dm.train_set[:32], 10), 1.0, cfg)
    assert abs(r["loss"] - float(g["train_f64_loss"])) < 1e-6
    assert np.max(np.abs(r["logits"] - g["train_f64_logits"])) < 1e-5
    assert abs(r["norm_tf"] - float(g["train_f64_norm_tf"])) < 1e-6
    auc, r1, r2 = O.eval_auc(params, dm.icl, O.collate_test(dm.test_set[:128], 10), config=cfg)
    assert auc == float(g["test_f64_auc"])
    assert np.max(np.abs(r1 - g["test_f64_pos"])) < 1e-5


def test_backward_matches_finite_differences():
    rng = np.random.default_rng(0)
    NU, NI, NC, L, S, B = 6, 20, 3, 4, 3, 5
    cfg = O.default_config(NU, NI, NC, Ls=L)
    params = O.randomize_params(O.init_params(cfg), seed=3)
    icl = rng.integers(0, NC, NI).astype(np.int32)
    batch = synth_batch(rng, B, L, S, NI, NU, NC)
    r = O.train_step(params, icl, batch, 0.0, cfg, dtype=torch.float64)

    def loss_at(name, idx, eps):
        p = {k: np.array(v, np.float64) for k, v in params.items()}
        p[name][idx] += eps
        return O.train_step(p, icl, batch, 0.0, cfg, dtype=torch.float64)["loss"]

    for name in O.PARAM_NAMES:
        g = r["grads"][name]
        flat = np.abs(g).reshape(-1)
        idx = np.unravel_index(int(np.argmax(flat)), g.shape) if g.shape else ()
        fd = (loss_at(name, idx, 1e-6) - loss_at(name, idx, -1e-6)) / 2e-6
        assert abs(fd - g[idx]) <= 1e-5 * max(1e-4, abs(fd)), (name, fd, g[idx])


def test_softmax_bias_has_zero_gradient_and_masked_tokens_are_inert(dm):
    cfg, params = _setup(dm)
    batch = O.collate_train(dm.train_set[:64], 10)
    r = O.train_step(params, dm.icl, batch, 1.0, cfg, dtype=torch.float64)
    for k, v in r["grads"].items():
        if k.endswith("bn_dense_map2/linear_map/bias"):
            assert np.max(np.abs(v)) < 1e-12
    # padded ids are never used: changing them leaves the logits unchanged
    b2 = list(batch); b2[3] = batch[3].copy(); b2[4] = batch[4].copy()
    pad = np.arange(10)[None, :] >= np.asarray(batch[6])[:, None]
    b2[3][pad] = 7
    padn = np.arange(batch[4].shape[1])[None, :] >= np.asarray(batch[7])[:, None]
    b2[4][padn] = 9
    l1, _ = O.forward_logits(params, dm.icl, batch)
    l2, _ = O.forward_logits(params, dm.icl, tuple(b2))
    assert np.array_equal(l1, l2)


def test_clip_definitions_agree_when_inactive_and_differ_when_active(dm):
    cfg, params = _setup(dm)
    batch = O.collate_train(dm.train_set[:32], 10)
    a = O.train_step(params, dm.icl, batch, 1.0, cfg, clip_mode="tf")
    b = O.train_step(params, dm.icl, batch, 1.0, cfg, clip_mode="aggregated")
    assert a["scale"] == 1.0 and b["scale"] == 1.0
    assert a["norm_tf"] >= a["norm_agg"]                       # triangle inequality on duplicate rows
    for k in a["new_params"]:
        assert np.array_equal(a["new_params"][k], b["new_params"][k])
    tight = dict(cfg, max_gradient_norm=0.1)
    c = O.train_step(params, dm.icl, batch, 1.0, tight, clip_mode="tf")
    assert c["scale"] < 1.0 and abs(c["scale"] - 0.1 / c["norm_tf"]) < 1e-6


def test_streaming_topk_semantics():
    scores = np.array([[0.1, 0.9, 0.9, 0.3], [0.5, 0.5, 0.5, 0.5]])
    assert list(O.label_ranks(scores, [2, 0])) == [1, 0]       # ties -> lower index first
    st = O.StreamingTopK()
    p, r = st.update(scores, [2, 3])
    assert p[0] == 0.0 and r[0] == 0.0 and r[1] == 1.0 and abs(p[1] - 2 / 20) < 1e-12


@pytest.mark.parametrize("opt", ["adam", "rmsprop", "adadelta"])
def test_oracle_optimizers_follow_the_tf18_apply_ops(opt):
    """_opt_update against a scalar-loop restatement of TF 1.8's ApplyAdam / ApplyRMSProp / ApplyAdadelta functors
    (training_ops.cc), and the sparse row rule for item_b."""
    rng = np.random.default_rng(0)
    w = rng.standard_normal(7); g = rng.standard_normal(7); s1 = rng.random(7) + 0.5; s2 = rng.random(7) * 0.1
    t, lr = 3, 0.01
    tw, ts1, ts2 = O._opt_update(opt, torch.tensor(w), torch.tensor(g), lr, torch.tensor(s1), torch.tensor(s2), t)
    for i in range(7):
        if opt == "adam":
            m = 0.9 * s1[i] + 0.1 * g[i]; v = 0.999 * s2[i] + 0.001 * g[i] ** 2
            alpha = lr * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
            exp = (w[i] - alpha * m / (np.sqrt(v) + 1e-8), m, v)
        elif opt == "rmsprop":
            ms = s1[i] + (g[i] ** 2 - s1[i]) * (1 - 0.9); mom = 0.0 * s2[i] + lr * g[i] / np.sqrt(ms + 1e-10)
            exp = (w[i] - mom, ms, mom)
        else:
            acc = 0.95 * s1[i] + 0.05 * g[i] ** 2
            upd = np.sqrt(s2[i] + 1e-8) / np.sqrt(acc + 1e-8) * g[i]
            exp = (w[i] - lr * upd, acc, 0.95 * s2[i] + 0.05 * upd ** 2)
        assert np.allclose([float(tw[i]), float(ts1[i]), float(ts2[i])], exp, rtol=1e-12, atol=0)
    if opt != "adam":          # rows outside the IndexedSlices: untouched
        mask = torch.tensor([True, False] * 3 + [True])
        tw, ts1, ts2 = O._opt_update(opt, torch.tensor(w), torch.tensor(g), lr, torch.tensor(s1), torch.tensor(s2), t, mask)
        assert np.array_equal(tw.numpy()[1::2], w[1::2]) and np.array_equal(ts1.numpy()[1::2], s1[1::2])
