"""The packed feed (tlsan_b200.input.PackedBatch): DataInput(..., packed=True) must yield the reference batcher's
fields (TLSAN/input.py:17-54,70-107) -- checked against the recorded outputs of the unmodified reference batcher --
and training / scoring from it must equal training from the 9-tuple bit for bit."""
import os

import numpy as np
import pytest

from oracle import tlsan_oracle as O
from tests.util import GOLD, load_digital_music, model_from_params
from tlsan_b200 import DataInput, DataInputTest, PackedBatch
from tlsan_b200.input import staging_layout


@pytest.fixture(scope="module")
def dm():
    return load_digital_music()


def test_packed_batches_equal_reference_batcher_outputs(dm):
    g = np.load(os.path.join(GOLD, "input_batches.npz"))
    names = sorted({k.rsplit("_f", 1)[0] for k in g.files})
    checked = 0
    for name in names:                                  # e.g. train_bs32_k10_b0 : recorded reference batches
        kind, bs, k, bi = name.split("_")
        bs, k, bi = int(bs[2:]), int(k[1:]), int(bi[1:])
        cls, data = (DataInput, dm.train_set) if kind == "train" else (DataInputTest, dm.test_set)
        it = cls(data, bs, k, packed=True)
        for _ in range(bi + 1):
            _, pb = next(it)
        assert isinstance(pb, PackedBatch) and len(pb) == 9
        for f in range(9):
            ref = g["%s_f%d" % (name, f)]
            got = np.asarray(pb[f])
            assert got.shape == ref.shape and np.array_equal(got, ref), (name, f)
        checked += 1
    assert checked >= 7


def test_from_tuple_equals_native_packed_layout(dm):
    for cls, data in ((DataInput, dm.train_set[:700]), (DataInputTest, dm.test_set[:300])):
        for (_, t), (_, p) in zip(cls(data, 128, 10), cls(data, 128, 10, packed=True)):
            q = PackedBatch.from_tuple(t, cls.is_test)
            assert (q.B, q.L, q.S, q.n_new) == (p.B, p.L, p.S, p.n_new)
            offs, words = staging_layout(p.B, p.L, p.S)
            for name, n in (("u", p.B), ("i", p.B), ("second", p.B), ("c", p.B), ("sl", p.B), ("sl_new", p.B),
                            ("hist_i", p.B * p.L), ("hist_t", p.B * p.L), ("new_off", p.B), ("new_items", p.n_new)):
                assert np.array_equal(q.buf[offs[name]:offs[name] + n], p.buf[offs[name]:offs[name] + n]), name
            assert all(a <= b for a, b in zip(q.id_max, p.id_max))


def test_ring_buffers_are_recycled_after_depth_batches(dm):
    it = DataInput(dm.train_set[:64 * 6], 64, 10, packed=True)
    bufs = [pb.buf.ctypes.data for _, pb in it]
    assert len(set(bufs)) == 4 and bufs[0] == bufs[4]


@pytest.mark.gpu
def test_training_from_packed_batches_is_bit_identical(dm):
    import torch
    cfg = O.default_config(*dm.counts)
    params = O.randomize_params(O.init_params(cfg), 7)
    a, b = model_from_params(params, dm.icl, cfg), model_from_params(params, dm.icl, cfg)
    tup, pk = DataInput(dm.train_set[:96 * 5], 96, 10), DataInput(dm.train_set[:96 * 5], 96, 10, packed=True)
    nxt = next(pk)[1]
    for (_, t) in tup:
        p = nxt
        nxt = next(pk, (None, None))[1]
        la = a.train(None, t, 1.0)
        lb = b.train(None, p, 1.0, prefetch=nxt, lazy_loss=True)   # double-buffered packed feed: next batch staged on the
        assert la == float(lb) and la + 0.0 == lb + 0.0            # copy stream and presorted behind this step; lazy read-back
    for k, v in a.state_dict().items():
        assert torch.equal(v, b.state_dict()[k]), k
    # scoring: eval_auc from a packed test batch
    (_, tt), (_, tp) = next(DataInputTest(dm.test_set[:128], 128, 10)), next(DataInputTest(dm.test_set[:128], 128, 10, packed=True))
    assert a.eval_auc(None, tt) == b.eval_auc(None, tp)


@pytest.mark.gpu
def test_packed_feed_range_check_is_loud(dm):
    cfg = O.default_config(dm.counts[0], dm.counts[1] - 200, dm.counts[2])       # tables smaller than the data's ids
    m = model_from_params(O.init_params(cfg), dm.icl[:dm.counts[1] - 200], cfg)
    _, pb = next(DataInput(dm.train_set[:32], 32, 10, packed=True))
    with pytest.raises(IndexError):
        m.train(None, pb, 1.0)
