"""Generate tests/golden/* by EXECUTING the unmodified reference data path in this container.

TEST INFRASTRUCTURE.  Run once here (``python oracle/make_golden.py``); the outputs are
committed because /root/reference does not exist on the GPU box.

  1. runs /root/reference/TLSAN/build_dataset.py as-is (runpy, cwd = scratch dir whose
     ``../Data`` is a symlink to /root/reference/Data; one shim: ``pd.value_counts`` was
     removed in pandas 3) -> dataset.pkl, stored as CSR arrays in digital_music.npz; the raw
     review columns it read are stored next to it (digital_music_reviews.npz);
  2. imports /root/reference/TLSAN/input.py as-is and records DataInput / DataInputTest
     outputs for several (batch_size, k) settings -> input_batches.npz;
  3. evaluates the oracle restatement (fp64 and fp32) on fixed seeded weights for the first
     batches -> model_golden.npz (self-generated regression pins: the reference has no
     golden vectors for the model, see oracle/tlsan_oracle.py header).

No reference source is copied; only its outputs are stored.
"""
import importlib.util
import os
import pickle
import runpy
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
sys.path.insert(0, ROOT)


def run_reference_build_dataset():
    import pandas as pd
    if not hasattr(pd, "value_counts"):
        pd.value_counts = lambda x: pd.Series(x).value_counts()
    work = tempfile.mkdtemp(prefix="tlsan_golden_")
    os.symlink(os.path.join(REF, "Data"), os.path.join(work, "Data"))
    cwd = os.path.join(work, "TLSAN")
    os.makedirs(cwd)
    old = os.getcwd()
    os.chdir(cwd)
    try:
        runpy.run_path(os.path.join(REF, "TLSAN", "build_dataset.py"), run_name="__main__")
    finally:
        os.chdir(old)
    with open(os.path.join(cwd, "dataset.pkl"), "rb") as f:
        train_set = pickle.load(f)
        test_set = pickle.load(f)
        counts = pickle.load(f)
        icl = pickle.load(f)
    return train_set, test_set, counts, icl


def to_csr(samples, is_test):
    from oracle.tlsan_oracle import bucket_lut
    lut = bucket_lut()
    inv = {float(lut[n]): n for n in range(1, 13)}
    uid = np.array([t[0] for t in samples], np.int32)
    pre_off = np.zeros(len(samples) + 1, np.int64)
    new_off = np.zeros(len(samples) + 1, np.int64)
    pre_items, pre_bucket, new_items = [], [], []
    for n, t in enumerate(samples):
        assert len(t[1]) == len(t[3])
        pre_items.extend(t[1]); new_items.extend(t[2])
        for v in t[3]:
            # the value reaching the graph is float32(v) (input.py:36,45): store its bucket
            b = inv[float(np.float32(v))]
            assert np.float32(v) == lut[b]
            pre_bucket.append(b)
        pre_off[n + 1] = len(pre_items); new_off[n + 1] = len(new_items)
    out = dict(uid=uid, pre_off=pre_off, new_off=new_off,
               pre_items=np.array(pre_items, np.int32), pre_bucket=np.array(pre_bucket, np.uint8),
               new_items=np.array(new_items, np.int32))
    if is_test:
        out["pos"] = np.array([t[4][0] for t in samples], np.int32)
        out["neg"] = np.array([t[4][1] for t in samples], np.int32)
        out["ucate"] = np.array([t[5] for t in samples], np.int32)
    else:
        out["target"] = np.array([t[4] for t in samples], np.int32)
        out["label"] = np.array([t[5] for t in samples], np.int32)
        out["ucate"] = np.array([t[6] for t in samples], np.int32)
    return out


def from_csr(d, prefix, is_test):
    """Inverse of to_csr: rebuild the list-of-tuples layout of build_dataset.py:58-59,71."""
    from oracle.tlsan_oracle import bucket_lut
    lut = bucket_lut().astype(np.float64)
    cache = {k[len(prefix):]: np.asarray(d[k]) for k in d.keys() if k.startswith(prefix)}   # NpzFile re-reads per access
    g = lambda k: cache[k]
    out = []
    po, no = g("pre_off"), g("new_off")
    pi, pb, nw = g("pre_items"), g("pre_bucket"), g("new_items")
    for n in range(len(g("uid"))):
        pre = pi[po[n]:po[n + 1]].tolist()
        # 1/n as float64, like build_dataset.py:20 produces
        tim = [1 / np.float64(b) for b in pb[po[n]:po[n + 1]]]
        new = nw[no[n]:no[n + 1]].tolist()
        if is_test:
            out.append((int(g("uid")[n]), pre, new, tim, (int(g("pos")[n]), int(g("neg")[n])), int(g("ucate")[n])))
        else:
            out.append((int(g("uid")[n]), pre, new, tim, int(g("target")[n]), int(g("label")[n]), int(g("ucate")[n])))
    return out


def load_reference_input():
    spec = importlib.util.spec_from_file_location("ref_input", os.path.join(REF, "TLSAN", "input.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def record_batches(ref_input, train_set, test_set):
    out = {}
    cases = [("train", 32, 10, [0, 1, 2, 1186]), ("train", 1024, 10, [0, 37]), ("train", 128, 90, [0, 5]),
             ("train", 7, 3, [0, 11]), ("test", 128, 10, [0, 12]), ("test", 64, 90, [3]), ("test", 5, 1, [2])]
    for split, bs, k, which in cases:
        data = train_set if split == "train" else test_set
        it = (ref_input.DataInput if split == "train" else ref_input.DataInputTest)(data, bs, k)
        for step, batch in it:
            if step - 1 in which:
                tag = "%s_bs%d_k%d_b%d" % (split, bs, k, step - 1)
                for j, arr in enumerate(batch):
                    out["%s_f%d" % (tag, j)] = np.asarray(arr)
    return out


def model_goldens(train_set, test_set, counts, icl):
    import torch
    from oracle import tlsan_oracle as O
    cfg = O.default_config(*counts)
    params = O.randomize_params(O.init_params(cfg, seed=1234), seed=7)
    out = {}
    b_train = O.collate_train(train_set[:32], 10)
    b_test = O.collate_test(test_set[:128], 10)
    for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        r = O.train_step(params, icl, b_train, 1.0, cfg, dtype=dt)
        out["train_%s_loss" % name] = np.float64(r["loss"])
        out["train_%s_logits" % name] = r["logits"]
        out["train_%s_norm_tf" % name] = np.float64(r["norm_tf"])
        out["train_%s_norm_agg" % name] = np.float64(r["norm_agg"])
        for k, v in r["grads"].items():
            if k in ("item_emb", "user_emb", "usert_emb", "cate_emb", "item_b"):
                continue
            out["train_%s_grad/%s" % (name, k)] = v
        auc, r1, r2 = O.eval_auc(params, icl, b_test, dtype=dt, config=cfg)
        out["test_%s_auc" % name] = np.float64(auc)
        out["test_%s_pos" % name] = r1
        out["test_%s_neg" % name] = r2
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    train_set, test_set, counts, icl = run_reference_build_dataset()
    print("reference build_dataset.py:", len(train_set), "train /", len(test_set), "test", counts)
    d = {"counts": np.array(counts, np.int64), "icl": np.asarray(icl, np.int32)}
    for k, v in to_csr(train_set, False).items():
        d["train_" + k] = v
    for k, v in to_csr(test_set, True).items():
        d["test_" + k] = v
    np.savez_compressed(os.path.join(GOLD, "digital_music.npz"), **d)
    # raw inputs of build_dataset.py (the three columns of reviews_df; item -> category is `icl`):
    # lets tests rebuild the dataset with tlsan_b200.build_dataset and compare with the reference output
    with open(os.path.join(REF, "Data", "Digital_Music.pkl"), "rb") as f:
        reviews_df, meta_df = pickle.load(f)
    assert np.array_equal(meta_df["categories"].values, np.asarray(icl))
    np.savez_compressed(os.path.join(GOLD, "digital_music_reviews.npz"),
                        reviewer=reviews_df["reviewerID"].values.astype(np.int32),
                        asin=reviews_df["asin"].values.astype(np.int32),
                        day=reviews_df["unixReviewTime"].values.astype(np.int32))
    # round trip must reproduce the pickled samples exactly (floats compared as float32,
    # which is all that ever reaches the graph)
    back = from_csr(d, "train_", False)
    for a, b in zip(back[:2000], train_set[:2000]):
        assert a[:3] == (b[0], b[1], b[2]) and a[4:] == (b[4], b[5], int(b[6]))
        assert np.array_equal(np.float32(a[3]), np.float32(b[3]))
    ref_input = load_reference_input()
    np.savez_compressed(os.path.join(GOLD, "input_batches.npz"), **record_batches(ref_input, train_set, test_set))
    np.savez_compressed(os.path.join(GOLD, "model_golden.npz"), **model_goldens(train_set, test_set, counts, icl))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
