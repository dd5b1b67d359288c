"""Scalar summaries without TensorFlow.

The reference attaches ``tf.summary.FileWriter`` objects to the model (``model.train_writer`` / ``model.eval_writer``,
TLSAN/model.py:174-183) and writes 'Training Loss' every ``display_freq`` steps (model.py:228-230) and 'AUC', 'P@k',
'R@k' at every evaluation (TLSAN/train.py:91-117).  ``JsonlSummaryWriter`` keeps those call sites working: it accepts
``add_scalar(tag, value, step)`` (what ``Model.train`` calls) and ``add_summary(summary, global_step)`` with either a
``{tag: value}`` dict or a TF ``Summary`` proto (duck-typed: ``summary.value[i].tag / .simple_value``), and appends
one JSON line per scalar to ``<logdir>/scalars.jsonl`` -- trivially convertible to TensorBoard event files by
whoever has TensorFlow installed."""
import json
import os
import time


class JsonlSummaryWriter(object):
    def __init__(self, logdir):
        os.makedirs(logdir, exist_ok=True)
        self.path = os.path.join(logdir, "scalars.jsonl")
        self._f = open(self.path, "a")

    def add_scalar(self, tag, value, step):
        self._f.write(json.dumps({"tag": str(tag), "value": float(value), "step": int(step), "wall": time.time()}) + "\n")

    def add_summary(self, summary, global_step=None):
        if isinstance(summary, dict):
            items = summary.items()
        else:                                            # tf.Summary(value=[tf.Summary.Value(tag=..., simple_value=...)])
            items = [(v.tag, v.simple_value) for v in summary.value]
        for tag, value in items:
            self.add_scalar(tag, value, 0 if global_step is None else global_step)

    def flush(self):
        self._f.flush()

    def close(self):
        self._f.close()


def attach_writers(model, model_dir=None):
    """``model.train_writer`` / ``model.eval_writer`` under <model_dir>/train and /eval (model.py:180-183)."""
    root = model_dir or model.config.get("model_dir", "save_path")
    model.train_writer = JsonlSummaryWriter(os.path.join(root, "train"))
    model.eval_writer = JsonlSummaryWriter(os.path.join(root, "eval"))
    return model
