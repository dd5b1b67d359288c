"""The oracle against golden vectors produced by EXECUTING the reference's graph code (TLSAN/model.py, unmodified) on
the torch-backed TF-1.8 API shim (oracle/tf1_shim.py; generator oracle/make_model_golden.py).  CPU only; the vectors
are committed because /root/reference does not travel to the GPU box."""
import os

import numpy as np
import pytest
import torch

from oracle import tlsan_oracle as O
from tests.util import GOLD, load_digital_music


@pytest.fixture(scope="module")
def ctx():
    g = np.load(os.path.join(GOLD, "model_ref_graph.npz"))
    dm = load_digital_music()
    cfg = O.default_config(*dm.counts)
    params = O.randomize_params(O.init_params(cfg), seed=int(g["param_seed"]))
    return g, dm, cfg, params


def _close(a, b, tol=1e-10):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.max(np.abs(a - b)) <= tol * (np.max(np.abs(b)) + 1e-30), float(np.max(np.abs(a - b)))


def test_train_step_equals_the_reference_graph(ctx):
    """loss, every gradient of tf.gradients (model.py:198), global norm (:201) and the weights after
    GradientDescentOptimizer.apply_gradients (:204), float64."""
    g, dm, cfg, params = ctx
    lo, hi = (int(x) for x in g["train_rows"])
    batch = O.collate_train(dm.train_set[lo:hi], cfg["Ls"])
    ref = O.train_step(params, dm.icl, batch, float(g["lr"]), cfg, dtype=torch.float64, clip_mode="agg")
    _close(ref["loss"], g["train/loss"])
    _close(ref["norm_agg"], g["train/norm"])
    # the TF-1.8 reading of the global norm (un-aggregated IndexedSlices values), evaluated by the shim over the
    # lookups the reference graph really makes, equals the oracle's own bookkeeping of those slices
    _close(ref["norm_tf"], g["train/norm_tf"])
    assert abs(ref["norm_tf"] - ref["norm_agg"]) > 1e-2 * ref["norm_agg"]          # the two readings really differ
    assert ref["scale"] == 1.0                      # clip inactive on this batch
    for k in params:
        _close(ref["grads"][k], g["train/grad/" + k])
        _close(ref["new_params"][k], g["train/new/" + k])


def test_scoring_equals_the_reference_graph(ctx):
    """self.logits for both runs of Model.eval_auc (model.py:135-137, 239-261) and self.eval_logits (:140)."""
    g, dm, cfg, params = ctx
    lo, hi = (int(x) for x in g["test_rows"])
    tb = O.collate_test(dm.test_set[lo:hi], cfg["Ls"])
    r1, _ = O.forward_logits(params, dm.icl, tb, 1, dtype=torch.float64, config=cfg)
    r2, _ = O.forward_logits(params, dm.icl, tb, 2, dtype=torch.float64, config=cfg)
    _close(r1, g["test/logits_pos"])
    _close(r2, g["test/logits_neg"])
    sc = O.eval_logits_all(params, dm.icl, tb, dtype=torch.float64, config=cfg)
    _close(np.asarray(sc)[:8], g["test/eval_logits"])


@pytest.mark.parametrize("tag,workload,B,L", [("electronics", "electronics", 512, 10), ("movies", "movies", 512, 10),
                                              ("electronics_L90", "electronics", 64, 90)])
def test_baseline_shapes_equal_the_reference_graph(ctx, tag, workload, B, L):
    """synthetic batches of the BASELINE shapes (18-wide sessions, Ls = 90, 15 huge categories) through the reference
    graph: logits in full, loss, both readings of the global norm, sum |gradient| per variable"""
    from oracle import make_model_golden as G
    from tlsan_b200 import synth
    g = ctx[0]
    _, NU, NI, NC = synth.WORKLOADS[workload]
    rng = np.random.default_rng(G.SHAPE_SEED)
    cfg = O.default_config(NU, NI, NC, Ls=L)
    icl = rng.integers(0, NC, NI).astype(np.int32)
    batch = synth.synth_batches(rng, 1, B, L, NU, NI, NC)[0]
    params = O.randomize_params(O.init_params(cfg), seed=G.PARAM_SEED, scale=0.2)
    ref = O.train_step(params, icl, batch, G.LR, cfg, dtype=torch.float64)
    _close(ref["loss"], g[tag + "/loss"])
    _close(ref["norm_agg"], g[tag + "/norm"])
    _close(ref["norm_tf"], g[tag + "/norm_tf"])
    r1, _ = O.forward_logits(params, icl, batch, 1, dtype=torch.float64, config=cfg)
    _close(r1, g[tag + "/logits"])
    for k in params:
        _close(np.abs(np.asarray(ref["grads"][k], np.float64)).sum(), g[tag + "/gradabs/" + k], tol=1e-9)


def test_fp32_oracle_is_within_the_product_tolerance_of_the_reference_graph(ctx):
    """the float32 evaluation (what the CUDA kernels are compared with at 1e-4) against the float64 reference graph"""
    g, dm, cfg, params = ctx
    lo, hi = (int(x) for x in g["test_rows"])
    tb = O.collate_test(dm.test_set[lo:hi], cfg["Ls"])
    r1, _ = O.forward_logits(params, dm.icl, tb, 1, dtype=torch.float32, config=cfg)
    assert np.max(np.abs(r1 - g["test/logits_pos"]) / (np.abs(g["test/logits_pos"]) + 1e-2)) < 1e-5


@pytest.mark.skipif(not os.path.exists("/root/reference/TLSAN/model.py"), reason="the reference tree is only in the build container")
def test_generator_reproduces_the_committed_vectors(ctx, tmp_path, monkeypatch):
    """re-executes the unmodified reference file on the shim and compares with the committed fixture"""
    from oracle import make_model_golden as G
    monkeypatch.setattr(G, "GOLD", str(tmp_path))
    G.main()
    new = np.load(os.path.join(str(tmp_path), "model_ref_graph.npz"))
    g = ctx[0]
    assert sorted(new.files) == sorted(g.files)
    for k in g.files:
        assert np.array_equal(new[k], g[k]), k


def test_shim_ops_follow_the_documented_tf_semantics():
    from oracle import tf1_shim as tf
    x = torch.tensor([[-2.0, 0.5], [3.0, -0.1]], dtype=torch.float64)
    z = torch.tensor([[1.0, 0.0], [0.0, 1.0]], dtype=torch.float64)
    ce = tf.nn.sigmoid_cross_entropy_with_logits(logits=x, labels=z).numpy()
    want = -(z.numpy() * np.log(1 / (1 + np.exp(-x.numpy()))) + (1 - z.numpy()) * np.log(1 - 1 / (1 + np.exp(-x.numpy()))))
    assert np.allclose(ce, want, atol=1e-12)
    assert float(tf.nn.l2_loss(x)) == pytest.approx(float((x.numpy() ** 2).sum() / 2))
    m = tf.sequence_mask(torch.tensor([0, 2, 3]), 3).numpy()
    assert m.tolist() == [[False, False, False], [True, True, False], [True, True, True]]
    s = tf.nn.softmax(x, 1).numpy()
    assert np.allclose(s.sum(axis=1), 1.0) and np.allclose(s[0], np.exp(x.numpy()[0]) / np.exp(x.numpy()[0]).sum())
    parts = tf.split(torch.arange(24.).reshape(2, 3, 4), 2, axis=2)
    assert len(parts) == 2 and parts[0].shape == (2, 3, 2)
    assert torch.equal(tf.concat(parts, 2), torch.arange(24.).reshape(2, 3, 4))
    assert tf.tile(torch.ones(2, 3), [4, 1]).shape == (8, 3)
    clipped, norm = tf.clip_by_global_norm([torch.tensor([3.0, 4.0])], 2.5)
    assert float(norm) == 5.0 and np.allclose(clipped[0].numpy(), [1.5, 2.0])
