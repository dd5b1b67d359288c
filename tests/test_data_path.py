"""Batch layout + time buckets: oracle and product batcher against outputs of the UNMODIFIED
reference TLSAN/input.py and TLSAN/build_dataset.py (tests/golden, see oracle/make_golden.py).
Integer / index work: bit-exact."""
import os

import numpy as np
import pytest

from oracle import tlsan_oracle as O
from tests.util import GOLD
from tlsan_b200.input import CsrDataset, DataInput, DataInputTest

CASES = [("train", 32, 10, [0, 1, 2, 1186]), ("train", 1024, 10, [0, 37]), ("train", 128, 90, [0, 5]),
         ("train", 7, 3, [0, 11]), ("test", 128, 10, [0, 12]), ("test", 64, 90, [3]), ("test", 5, 1, [2])]


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "input_batches.npz"))


def _same(batch, gold, tag):
    for j in range(9):
        ref = gold["%s_f%d" % (tag, j)]
        got = np.asarray(batch[j])
        assert got.shape == ref.shape, (tag, j)
        assert np.array_equal(got, ref), (tag, j)
        if j in (3, 4, 5):
            assert got.dtype == ref.dtype, (tag, j)          # int64 / float32 like input.py:35-37


@pytest.mark.parametrize("split,bs,k,which", CASES)
def test_oracle_collate_matches_reference_input(dm, gold, split, bs, k, which):
    data = dm.train_set if split == "train" else dm.test_set
    fn = O.collate_train if split == "train" else O.collate_test
    for w in which:
        _same(fn(data[w * bs:(w + 1) * bs], k), gold, "%s_bs%d_k%d_b%d" % (split, bs, k, w))


@pytest.mark.parametrize("split,bs,k,which", CASES)
def test_product_batcher_matches_reference_input(dm, gold, split, bs, k, which):
    data = dm.train_set if split == "train" else dm.test_set
    it = (DataInput if split == "train" else DataInputTest)(data, bs, k)
    seen = 0
    for step, batch in it:
        if step - 1 in which:
            _same(batch, gold, "%s_bs%d_k%d_b%d" % (split, bs, k, step - 1))
            seen += 1
    assert seen == len(which)
    assert it.epoch_size == -(-len(data) // bs)


def test_product_batcher_equals_oracle_everywhere(dm):
    """every batch of a full test epoch + ragged tail, several k"""
    for k in (1, 10, 90):
        it = DataInputTest(dm.test_set, 100, k)
        for step, batch in it:
            ref = O.collate_test(dm.test_set[(step - 1) * 100: step * 100], k)
            for a, b in zip(batch, ref):
                assert np.array_equal(np.asarray(a), np.asarray(b))


def test_csr_dataset_accepts_prebuilt(dm):
    csr = CsrDataset.from_samples(dm.train_set[:500], is_test=False)
    a = next(DataInput(csr, 64, 10))[1]
    b = next(DataInput(dm.train_set[:500], 64, 10))[1]
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_bucket_lut_matches_reference_values(dm):
    """every stored hist_t of the reference dataset is float32(1/n), n = sum(d >= gap)"""
    lut = O.bucket_lut()
    b = dm.raw["train_pre_bucket"]
    assert b.min() >= 1 and b.max() <= 12
    for d in (2, 3, 4, 7, 8, 100, 4095, 4096, 5000, 100000):
        n = O.time_bucket(d)
        assert n == min(12, int(np.floor(np.log2(d))))
        assert O.time_weight(d) == lut[n]
    # reference property: hist_t is non-decreasing along the history (older -> larger gap)
    po = dm.raw["train_pre_off"]
    for s in range(0, 2000):
        seg = b[po[s]:po[s + 1]]
        assert np.all(np.diff(seg.astype(np.int32)) <= 0)
