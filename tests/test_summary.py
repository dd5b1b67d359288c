"""Scalar summaries: the reference's writer call sites keep working without TensorFlow."""
import json
from types import SimpleNamespace

from tlsan_b200.summary import JsonlSummaryWriter, attach_writers


def test_scalars_and_tf_style_summaries_land_in_jsonl(tmp_path):
    w = JsonlSummaryWriter(str(tmp_path / "eval"))
    w.add_scalar("Training Loss", 0.5, 100)                                   # Model.train (model.py:228-230)
    proto = SimpleNamespace(value=[SimpleNamespace(tag="AUC", simple_value=0.9687)])
    w.add_summary(summary=proto, global_step=1000)                            # train.py:91-95
    w.add_summary({"P@10": 0.0166}, 1000)
    w.close()
    rows = [json.loads(l) for l in open(w.path)]
    assert [(r["tag"], r["step"]) for r in rows] == [("Training Loss", 100), ("AUC", 1000), ("P@10", 1000)]
    assert abs(rows[1]["value"] - 0.9687) < 1e-12


def test_attach_writers_creates_train_and_eval_dirs(tmp_path):
    m = SimpleNamespace(config={"model_dir": str(tmp_path / "save_path")}, train_writer=None, eval_writer=None)
    attach_writers(m)
    m.train_writer.add_scalar("Training Loss", 1.0, 0)
    m.train_writer.flush()
    assert (tmp_path / "save_path" / "train" / "scalars.jsonl").exists()
    assert (tmp_path / "save_path" / "eval").is_dir()


def test_histogram_proto_follows_tf_buckets(tmp_path):
    """tf.summary.histogram (model.py:174-180): HistogramProto moments + counts over TF's exponential bucket limits."""
    import numpy as np
    import torch
    from tlsan_b200.summary import histogram_proto, tf_bucket_limits
    lim = tf_bucket_limits()
    assert lim[len(lim) // 2] == 0.0 and abs(lim[len(lim) // 2 + 1] - 1e-12) < 1e-24 and np.all(np.diff(lim) > 0)
    assert abs(lim[len(lim) // 2 + 2] / lim[len(lim) // 2 + 1] - 1.1) < 1e-12
    rng = np.random.default_rng(0)
    v = np.r_[rng.standard_normal(1000) * 0.3, 0.0, -1.0, 1.0]
    for h in (histogram_proto(v), histogram_proto(torch.as_tensor(v))):
        assert h["num"] == len(v) and sum(h["bucket"]) == len(v)
        assert abs(h["sum"] - v.sum()) < 1e-9 and abs(h["sum_squares"] - (v * v).sum()) < 1e-9
        assert h["min"] == v.min() and h["max"] == v.max()
        # every value lies in the bucket (previous limit, limit]... TF's rule: bucket = first limit > value
        edges = np.array(h["bucket_limit"])
        assert np.all(np.diff(edges) > 0)
        full = np.searchsorted(lim, v, side="right")
        for limit, count in zip(h["bucket_limit"], h["bucket"]):
            if count:
                assert count == int(np.sum(lim[full] == limit))
    w = JsonlSummaryWriter(str(tmp_path / "train"))
    w.add_histogram("gamma", np.array([1.0]), 7)
    w.close()
    rec = json.loads(open(w.hist_path).readline())
    assert rec["tag"] == "gamma" and rec["step"] == 7 and rec["num"] == 1 and rec["bucket"][-1] == 1
