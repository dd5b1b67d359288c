"""Scalar and histogram summaries without TensorFlow.

The reference attaches ``tf.summary.FileWriter`` objects to the model (``model.train_writer`` / ``model.eval_writer``,
TLSAN/model.py:174-183) and writes 'Training Loss' every ``display_freq`` steps (model.py:228-230) and 'AUC', 'P@k',
'R@k' at every evaluation (TLSAN/train.py:91-117).  ``JsonlSummaryWriter`` keeps those call sites working: it accepts
``add_scalar(tag, value, step)`` (what ``Model.train`` calls) and ``add_summary(summary, global_step)`` with either a
``{tag: value}`` dict or a TF ``Summary`` proto (duck-typed: ``summary.value[i].tag / .simple_value``), and appends
one JSON line per scalar to ``<logdir>/scalars.jsonl`` -- trivially convertible to TensorBoard event files by
whoever has TensorFlow installed.  ``add_histogram`` records what ``tf.summary.histogram`` stores (the HistogramProto
fields min / max / num / sum / sum_squares and the counts over TF's default exponential bucket limits,
+-1e-12 * 1.1^k) to ``<logdir>/histograms.jsonl``: the train summary of model.py:174-183 (gamma, the four embedding
tables, the attention output u_t)."""
import json
import os
import time

import numpy as np


def tf_bucket_limits():
    """Bucket limits of tensorflow/core/lib/histogram/histogram.cc (InitDefaultBucketsInner): 1e-12 * 1.1^k up to
    1e20, mirrored for negative values, 0 in the middle, DBL_MAX at the end."""
    pos = []
    v = 1e-12
    while v < 1e20:
        pos.append(v)
        v *= 1.1
    pos.append(np.finfo(np.float64).max)
    return np.array([-x for x in reversed(pos)] + [0.0] + pos)


_LIMITS = None


def histogram_proto(values):
    """HistogramProto of ``values`` (numpy or torch, any shape) as a dict; empty buckets are dropped the way
    Histogram::EncodeToProto does (runs of empty buckets collapse)."""
    global _LIMITS
    if _LIMITS is None:
        _LIMITS = tf_bucket_limits()
    try:
        import torch
        if isinstance(values, torch.Tensor):
            v = values.detach().reshape(-1).double()
            lim = torch.as_tensor(_LIMITS, device=v.device)
            idx = torch.bucketize(v, lim, right=True)                 # first limit > value (upper_bound)
            counts = torch.bincount(idx.clamp_(max=len(_LIMITS) - 1), minlength=len(_LIMITS)).cpu().numpy()
            stats = (float(v.min()), float(v.max()), int(v.numel()), float(v.sum()), float((v * v).sum()))
            return _encode(counts, stats)
    except ImportError:
        pass
    v = np.asarray(values, np.float64).reshape(-1)
    idx = np.minimum(np.searchsorted(_LIMITS, v, side="right"), len(_LIMITS) - 1)
    counts = np.bincount(idx, minlength=len(_LIMITS))
    return _encode(counts, (float(v.min()), float(v.max()), int(v.size), float(v.sum()), float((v * v).sum())))


def _encode(counts, stats):
    limit, bucket = [], []
    i, n = 0, len(counts)
    while i < n:
        j = i
        if counts[i] == 0:                        # a run of empty buckets is stored as ONE empty bucket (its last limit)
            while j + 1 < n and counts[j + 1] == 0:
                j += 1
            if j == n - 1:
                break
        limit.append(float(_LIMITS[j])); bucket.append(int(counts[j]))
        i = j + 1
    return dict(min=stats[0], max=stats[1], num=stats[2], sum=stats[3], sum_squares=stats[4], bucket_limit=limit,
                bucket=bucket)


class JsonlSummaryWriter(object):
    def __init__(self, logdir):
        os.makedirs(logdir, exist_ok=True)
        self.path = os.path.join(logdir, "scalars.jsonl")
        self._f = open(self.path, "a")
        self.hist_path = os.path.join(logdir, "histograms.jsonl")
        self._h = None

    def add_histogram(self, tag, values, step):
        if self._h is None:
            self._h = open(self.hist_path, "a")
        rec = histogram_proto(values)
        rec.update(tag=str(tag), step=int(step), wall=time.time())
        self._h.write(json.dumps(rec) + "\n")

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
        if self._h is not None:
            self._h.flush()

    def close(self):
        self._f.close()
        if self._h is not None:
            self._h.close()


def attach_writers(model, model_dir=None):
    """``model.train_writer`` / ``model.eval_writer`` under <model_dir>/train and /eval (model.py:180-183)."""
    root = model_dir or model.config.get("model_dir", "save_path")
    model.train_writer = JsonlSummaryWriter(os.path.join(root, "train"))
    model.eval_writer = JsonlSummaryWriter(os.path.join(root, "eval"))
    return model
