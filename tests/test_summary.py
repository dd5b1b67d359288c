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
