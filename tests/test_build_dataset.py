"""Session segmentation, negative sampling, u_cate, train/test split and time buckets
(TLSAN/build_dataset.py) -- our vectorised builder against the recorded output of the UNMODIFIED
reference script (tests/golden/digital_music.npz, see oracle/make_golden.py).  Bit-exact."""
import os

import numpy as np
import pytest

from oracle import tlsan_oracle as O
from tests.util import GOLD
from tlsan_b200.build_dataset import build_dataset


@pytest.fixture(scope="module")
def rebuilt(dm):
    r = np.load(os.path.join(GOLD, "digital_music_reviews.npz"))
    return build_dataset(r["reviewer"], r["asin"], r["day"], dm.icl, dm.counts[1], return_gaps=True)


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x[:3] == y[:3] and tuple(x[4:]) == tuple(y[4:])
        assert np.array_equal(np.float32(x[3]), np.float32(y[3]))


def test_builder_reproduces_reference_dataset_exactly(dm, rebuilt):
    train_set, test_set, _, _ = rebuilt
    assert len(train_set) == 37970 and len(test_set) == 1659 == dm.counts[0]
    _same(train_set, dm.train_set)          # same samples, same order: the random stream is preserved
    _same(test_set, dm.test_set)


def test_every_stored_weight_is_the_bucket_of_its_day_gap(dm, rebuilt):
    """float32(1/n(d)) == stored hist_t for EVERY history entry of every sample (SURVEY 8a note):
    pins bucketing and session segmentation together."""
    train_set, test_set, g_tr, g_te = rebuilt
    lut = O.bucket_lut()
    for samples, gaps, ref in ((train_set, g_tr, dm.train_set), (test_set, g_te, dm.test_set)):
        d = np.concatenate(gaps)
        assert d.min() >= 2                                  # sessions are distinct days
        n = np.minimum(12, np.floor(np.log2(d)).astype(np.int64))
        stored = np.concatenate([np.float32(t[3]) for t in ref])
        assert np.array_equal(lut[n], stored)


@pytest.mark.gpu
def test_gpu_bucket_kernel_reproduces_every_stored_weight(dm, rebuilt):
    import torch
    from tlsan_b200 import _lib
    lib = _lib.lib()
    _, _, g_tr, _ = rebuilt
    d = np.concatenate(g_tr).astype(np.int32)
    stored = np.concatenate([np.float32(t[3]) for t in dm.train_set])
    dd = torch.from_numpy(d).cuda(); lut = torch.from_numpy(O.bucket_lut()).cuda()
    out = torch.empty(len(d), device="cuda")
    _lib.check(lib.tlsan_time_bucket(dd.data_ptr(), lut.data_ptr(), out.data_ptr(), None, len(d), None))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), stored)


@pytest.mark.gpu
def test_gpu_builder_reproduces_reference_dataset_exactly(dm):
    """build_dataset_gpu (csrc/tlsan_builder.cu: segmentation, split rule, u_cate, time weights as kernels; the Python
    random stream stays on the host) must leave in HBM exactly the samples of the UNMODIFIED reference script, in its
    order: every CSR array against tests/golden/digital_music.npz, and the raw day gaps against the stored weights."""
    from tlsan_b200.build_dataset import build_dataset_gpu
    r = np.load(os.path.join(GOLD, "digital_music_reviews.npz"))
    train, test, gap_tr, gap_te = build_dataset_gpu(r["reviewer"], r["asin"], r["day"], dm.icl, dm.counts[1], with_gaps=True)
    g = dm.raw
    lut = O.bucket_lut()
    for ds, pre, second_key, second_ref, gaps in ((train, "train_", "second_f", g["train_label"].astype(np.float32), gap_tr),
                                                  (test, "test_", "second_i", g["test_neg"], gap_te)):
        c = ds.to_csr()
        assert len(ds) == len(g[pre + "uid"])
        assert np.array_equal(c.uid, g[pre + "uid"]) and np.array_equal(c.ucate, g[pre + "ucate"])
        assert np.array_equal(c.pre_off, g[pre + "pre_off"]) and np.array_equal(c.new_off, g[pre + "new_off"])
        assert np.array_equal(c.pre_items, g[pre + "pre_items"]) and np.array_equal(c.new_items, g[pre + "new_items"])
        assert np.array_equal(c.cand, g[pre + ("target" if pre == "train_" else "pos")])
        assert np.array_equal(np.asarray(c.second, second_ref.dtype), second_ref)
        assert np.array_equal(c.pre_time, lut[g[pre + "pre_bucket"]])                  # float32(1/n), bit for bit
        d = gaps.cpu().numpy()[:len(c.pre_items)]
        assert d.min() >= 2 and np.array_equal(np.minimum(12, np.floor(np.log2(d)).astype(np.int64)), g[pre + "pre_bucket"])
    # and it trains: a batch assembled from the GPU-built dataset equals the reference batcher on the reference samples
    batch = train.batch(np.arange(64), 10)
    ref = O.collate_train(dm.train_set[:64], 10)
    import torch
    got = batch.buf.cpu().numpy()
    offs = batch.offs
    assert np.array_equal(got[offs["hist_i"]:offs["hist_i"] + 640].reshape(64, 10), ref[3])
    assert np.array_equal(got[offs["u"]:offs["u"] + 64], np.asarray(ref[0]))
