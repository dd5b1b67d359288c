"""Device-resident dataset: GPU batch assembly is bit-exact against the outputs of the
unmodified reference TLSAN/input.py (tests/golden/input_batches.npz) and against the host batcher."""
import os

import numpy as np
import pytest
import torch

from oracle import tlsan_oracle as O
from tests.util import GOLD, model_from_params

pytestmark = pytest.mark.gpu
CASES = [("train", 32, 10, [0, 1, 2, 1186]), ("train", 1024, 10, [0, 37]), ("train", 128, 90, [0, 5]),
         ("train", 7, 3, [0, 11]), ("test", 128, 10, [0, 12]), ("test", 64, 90, [3]), ("test", 5, 1, [2])]


def _fields(db):
    h = db.buf.cpu().numpy()
    o, B, L, S = db.offs, db.B, db.L, db.S
    g = lambda k, n: h[o[k]:o[k] + n]
    sec = g("second", B)
    return dict(u=g("u", B), i=g("i", B), second=sec if db.is_test else sec.view(np.float32), c=g("c", B),
                sl=g("sl", B), sl_new=g("sl_new", B), hist_i=g("hist_i", B * L).reshape(B, L),
                hist_i_new=g("hist_i_new", B * S).reshape(B, S), hist_t=g("hist_t", B * L).view(np.float32).reshape(B, L))


@pytest.mark.parametrize("split,bs,k,which", CASES)
def test_device_collate_matches_reference_input(dm, split, bs, k, which):
    from tlsan_b200.dataset import DeviceDataset
    gold = np.load(os.path.join(GOLD, "input_batches.npz"))
    data = dm.train_set if split == "train" else dm.test_set
    ds = DeviceDataset(data, is_test=split == "test")
    for w in which:
        tag = "%s_bs%d_k%d_b%d" % (split, bs, k, w)
        idx = np.arange(w * bs, min((w + 1) * bs, len(data)))
        f = _fields(ds.batch(idx, k))
        ref = [gold["%s_f%d" % (tag, j)] for j in range(9)]
        assert np.array_equal(f["u"], ref[0]) and np.array_equal(f["i"], ref[1]) and np.array_equal(f["second"], ref[2])
        assert np.array_equal(f["hist_i"], ref[3]) and np.array_equal(f["hist_i_new"], ref[4])
        assert np.array_equal(f["hist_t"], ref[5])                     # float32 bit patterns
        assert np.array_equal(f["sl"], ref[6]) and np.array_equal(f["sl_new"], ref[7]) and np.array_equal(f["c"], ref[8])


def test_device_collate_random_rows_and_max_width(dm):
    from tlsan_b200.dataset import DeviceDataset
    from tlsan_b200.input import CsrDataset
    csr = CsrDataset.from_samples(dm.train_set[:5000], False)
    ds = DeviceDataset(csr, is_test=False)
    rng = np.random.default_rng(1)
    idx = rng.permutation(5000)[:777]
    ref = csr.collate(idx, 10)
    f = _fields(ds.batch(idx, 10))
    for name, j in (("u", 0), ("i", 1), ("second", 2), ("hist_i", 3), ("hist_i_new", 4), ("hist_t", 5), ("sl", 6),
                    ("sl_new", 7), ("c", 8)):
        assert np.array_equal(f[name], np.asarray(ref[j]).astype(f[name].dtype)), name
    fm = _fields(ds.batch(torch.from_numpy(idx.astype(np.int32)).cuda(), 10, width="max"))
    S = ref[4].shape[1]
    assert fm["hist_i_new"].shape[1] == ds.max_new_len
    assert np.array_equal(fm["hist_i_new"][:, :S], ref[4]) and not fm["hist_i_new"][:, S:].any()
    with pytest.raises(IndexError):
        ds.batch(np.array([5000]), 10)


def test_training_from_device_dataset_equals_host_batches(dm):
    from tlsan_b200.dataset import DeviceDataset
    cfg = O.default_config(*dm.counts)
    params = O.randomize_params(O.init_params(cfg, seed=1234), seed=7)
    ds = DeviceDataset(dm.train_set[:640], is_test=False)
    ma, mb = model_from_params(params, dm.icl, cfg), model_from_params(params, dm.icl, cfg)
    for s in range(5):
        idx = np.arange(s * 128, (s + 1) * 128)
        la = float(ma.train_staged(ds.batch(idx, 10), 1.0)[0].item())
        lb = mb.train(None, O.collate_train(dm.train_set[s * 128:(s + 1) * 128], 10), 1.0)
        assert la == lb
    for k, v in ma.state_dict().items():
        assert torch.equal(v, mb.state_dict()[k]), k
