"""Golden vectors for the MODEL arithmetic, produced by executing the reference's own graph code.

TEST INFRASTRUCTURE.  Run once in the build container (``python oracle/make_model_golden.py``); the output
``tests/golden/model_ref_graph.npz`` is committed because /root/reference does not exist on the GPU box.

The reference model lives in ``/root/reference/TLSAN/model.py`` and needs tensorflow == 1.8, which cannot be installed
here.  This script imports that file UNMODIFIED with ``sys.modules['tensorflow']`` bound to ``oracle/tf1_shim.py`` -- an
eager, torch-backed restatement of the ~50 public TF-1.8 API entries the file calls -- and binds the placeholders to
fixed Digital-Music batches, so constructing ``Model(config, item_cate_list)`` executes ``build_model`` (model.py:56-183),
``attention_net`` / ``feature_wise_attention`` and helpers (:316-483) and ``init_optimizer`` (:185-205) exactly as the
reference wrote them.  Recorded, in float64:

  train/   loss, logits, every variable's gradient (tf.gradients), global norm (dense-gradient reading and the TF-1.8
           un-aggregated IndexedSlices reading over the lookups the graph really makes), the weights after
           apply_gradients (sgd)
  test/    logits of the positive and the negative candidate (Model.eval_auc's two runs), eval_logits of 8 rows
  <shape>/ the same train quantities as digests (logits, loss, norms, sum |gradient| per variable) on synthetic batches of
           the BASELINE shapes: Electronics (B 512), Movies-TV (B 512), Electronics with Ls 90 (B 64)

What this pins: the graph the reference builds (which ops, in which order, on which shapes, with which masks and
variable scopes).  What it cannot pin: TF's own kernels (summation order; irrelevant in float64) and TF-internal
gradient plumbing (IndexedSlices un-aggregated norm) -- DESIGN.md section 2.  No reference source is copied.
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/TLSAN"
sys.path.insert(0, ROOT)

TRAIN_ROWS = (0, 64)          # dm.train_set[0:64]
TEST_ROWS = (0, 64)           # dm.test_set[0:64]
LR = 1.0
PARAM_SEED = 7
SHAPE_SEED = 1234
# (tag, workload of tlsan_b200/synth.py, batch, Ls)
SHAPES = (("electronics", "electronics", 512, 10), ("movies", "movies", 512, 10), ("electronics_L90", "electronics", 64, 90))


def run_reference_graph(config, icl, params, feeds):
    """Execute the reference Model constructor on bound placeholders; returns (model object, shim state)."""
    from oracle import tf1_shim
    tf1_shim.install(feeds, params, dtype=torch.float64)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    sys.modules.pop("model", None)
    ref_model = importlib.import_module("model")            # /root/reference/TLSAN/model.py, unmodified
    assert os.path.realpath(ref_model.__file__).startswith("/root/reference/"), ref_model.__file__
    cfg = dict(config)
    cfg["model_dir"] = "/tmp/tlsan_ref_graph"
    try:
        m = ref_model.Model(cfg, [int(c) for c in icl])
    finally:                                                 # leave no fake `tensorflow` / reference `model` module behind
        sys.modules.pop("tensorflow", None)
        sys.modules.pop("model", None)
    return m, tf1_shim.state()


def feeds_of(batch, lr, cand_index=1, y=None):
    """placeholder creation order of model.py:27-53: u, u_cate, i, y, hist_i, hist_i_new, hist_t, sl, sl_new, lr, is_training"""
    n = len(batch[0])
    return [batch[0], batch[8], batch[cand_index], np.zeros(n, np.float32) if y is None else y, batch[3], batch[4],
            batch[5], batch[6], batch[7], lr, True]


def main():
    from oracle import tlsan_oracle as O
    from tests.util import load_digital_music
    dm = load_digital_music()
    cfg = O.default_config(*dm.counts)
    params = O.randomize_params(O.init_params(cfg), seed=PARAM_SEED)
    out = {"train_rows": np.array(TRAIN_ROWS), "test_rows": np.array(TEST_ROWS), "lr": np.float64(LR),
           "param_seed": np.int64(PARAM_SEED)}

    batch = O.collate_train(dm.train_set[TRAIN_ROWS[0]:TRAIN_ROWS[1]], cfg["Ls"])
    m, st = run_reference_graph(cfg, dm.icl, params, feeds_of(batch, LR, 1, np.asarray(batch[2], np.float32)))
    out["train/loss"] = np.float64(m.loss.detach())
    out["train/logits"] = m.logits.detach().numpy()
    out["train/norm"] = np.float64(st.last_norm)                    # norm of the dense (summed) gradient
    out["train/norm_tf"] = np.float64(st.last_norm_tf)              # TF-1.8 reading: un-aggregated IndexedSlices values
    assert set(st.last_grads) == set(params), sorted(set(st.last_grads) ^ set(params))
    for k, v in st.last_grads.items():
        out["train/grad/" + k] = v
    for k, v in st.applied.items():
        out["train/new/" + k] = v

    tb = O.collate_test(dm.test_set[TEST_ROWS[0]:TEST_ROWS[1]], cfg["Ls"])
    m1, _ = run_reference_graph(cfg, dm.icl, params, feeds_of(tb, 0.0, 1))      # Model.eval_auc, first run  (:239-249)
    out["test/logits_pos"] = m1.logits.detach().numpy()
    out["test/eval_logits"] = m1.eval_logits.detach().numpy()[:8]              # model.py:140  (eval_prec / eval_recall)
    m2, _ = run_reference_graph(cfg, dm.icl, params, feeds_of(tb, 0.0, 2))      # second run: self.i = batch[2]  (:251-261)
    out["test/logits_neg"] = m2.logits.detach().numpy()

    # ---- digests on the BASELINE shapes (synthetic generators of tlsan_b200/synth.py): logits in full, loss, both norms
    # and sum |gradient| per variable -- enough to pin the graph on 18-wide sessions, Ls = 90, 15 huge categories
    from tlsan_b200 import synth
    for tag, workload, B, L in SHAPES:
        _, NU, NI, NC = synth.WORKLOADS[workload]
        rng = np.random.default_rng(SHAPE_SEED)
        scfg = O.default_config(NU, NI, NC, Ls=L)
        sicl = rng.integers(0, NC, NI).astype(np.int32)
        sb = synth.synth_batches(rng, 1, B, L, NU, NI, NC)[0]
        sp = O.randomize_params(O.init_params(scfg), seed=PARAM_SEED, scale=0.2)
        ms, sts = run_reference_graph(scfg, sicl, sp, feeds_of(sb, LR, 1, np.asarray(sb[2], np.float32)))
        out[tag + "/loss"] = np.float64(ms.loss.detach())
        out[tag + "/logits"] = ms.logits.detach().numpy()
        out[tag + "/norm"] = np.float64(sts.last_norm)
        out[tag + "/norm_tf"] = np.float64(sts.last_norm_tf)
        for k, v in sts.last_grads.items():
            out[tag + "/gradabs/" + k] = np.float64(np.abs(v).sum())

    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, "model_ref_graph.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1e3), "loss %.9f norm %.6f" % (out["train/loss"], out["train/norm"]))


if __name__ == "__main__":
    main()
