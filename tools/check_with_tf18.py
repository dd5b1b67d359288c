#!/usr/bin/env python
"""For a maintainer who HAS tensorflow 1.8: run the real reference graph on the batches / weights behind
tests/golden/model_ref_graph.npz and compare with the committed vectors (which were produced by executing the same
reference file on oracle/tf1_shim.py, float64).  NOT runnable in the build image (no TensorFlow) -- written against the
TF-1.8 API, never executed here.

  python tools/check_with_tf18.py /path/to/TLSAN        # the directory holding the reference model.py

Prints the largest relative difference per quantity; float32 TF against the float64 vectors should agree to ~1e-6
on loss / logits and ~1e-5 on gradients.  A difference in `norm` alone (with equal gradients) is the TF-internal
un-aggregated IndexedSlices global norm, not a graph difference.
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def main(ref_dir):
    import tensorflow as tf                                    # the real one, 1.8
    sys.path.insert(0, ref_dir)
    from model import Model                                    # the reference class
    from oracle import tlsan_oracle as O
    from tests.util import GOLD, load_digital_music
    g = np.load(os.path.join(GOLD, "model_ref_graph.npz"))
    dm = load_digital_music()
    cfg = O.default_config(*dm.counts)
    cfg["model_dir"] = "/tmp/tlsan_tf18_check"
    params = O.randomize_params(O.init_params(cfg), seed=int(g["param_seed"]))
    model = Model(cfg, [int(c) for c in dm.icl])
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a, np.float64) - b)) / (np.max(np.abs(b)) + 1e-30))
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        sess.run(tf.local_variables_initializer())
        tvars = tf.trainable_variables()
        for v in tvars:                                        # "item_emb:0" -> "item_emb"
            sess.run(v.assign(params[v.name.split(":")[0]]))
        lo, hi = (int(x) for x in g["train_rows"])
        b = O.collate_train(dm.train_set[lo:hi], cfg["Ls"])
        feed = {model.u: b[0], model.u_cate: b[8], model.i: b[1], model.y: b[2], model.hist_i: b[3],
                model.hist_i_new: b[4], model.hist_t: b[5], model.sl: b[6], model.sl_new: b[7],
                model.lr: float(g["lr"]), model.is_training: True}
        grads = tf.gradients(model.loss, tvars)
        dense = [tf.convert_to_tensor(x) for x in grads]       # IndexedSlices -> dense (duplicates summed)
        norm = tf.global_norm(grads)                           # what clip_by_global_norm sees (model.py:201)
        loss, logits, gvals, nval = sess.run([model.loss, model.logits, dense, norm], feed)
        print("train loss   ", rel(loss, g["train/loss"]))
        print("train logits ", rel(logits, g["train/logits"]))
        for v, gv in zip(tvars, gvals):
            print("grad %-70s %.3e" % (v.name, rel(gv, g["train/grad/" + v.name.split(":")[0]])))
        print("global norm: TF %.9f   dense-gradient norm in the vectors %.9f" % (nval, float(g["train/norm"])))
        lo, hi = (int(x) for x in g["test_rows"])
        tb = O.collate_test(dm.test_set[lo:hi], cfg["Ls"])
        for idx, key in ((1, "test/logits_pos"), (2, "test/logits_neg")):
            feed = {model.u: tb[0], model.u_cate: tb[8], model.i: tb[idx], model.hist_i: tb[3], model.hist_i_new: tb[4],
                    model.hist_t: tb[5], model.sl: tb[6], model.sl_new: tb[7], model.is_training: False}
            print(key, rel(sess.run(model.logits, feed), g[key]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/TLSAN")
