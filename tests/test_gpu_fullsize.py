"""BASELINE.json configs[1] at full size (Electronics-shape synthetic, per-GPU batch 65 536):
size-independent properties of the CUDA path + oracle spot checks on sampled rows."""
import numpy as np
import pytest
import torch

from oracle import tlsan_oracle as O
from tests.util import rel_err

pytestmark = pytest.mark.gpu
B, L = 65536, 10


@pytest.fixture(scope="module")
def setup():
    import bench
    from tlsan_b200.model import Model
    rng = np.random.default_rng(1234)
    cfg = O.default_config(bench.NU, bench.NI, bench.NC, Ls=L)
    icl = rng.integers(0, bench.NC, bench.NI).astype(np.int32)
    batch = bench.synth_batches(rng, 1, B, L)[0]
    params = O.randomize_params(O.init_params(cfg, seed=1234), seed=3, scale=0.2)

    def fresh():
        m = Model(cfg, icl, seed=1)
        m.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in params.items()})
        return m
    return cfg, icl, batch, params, fresh


def _as_test(batch, rng, NI):
    b = list(batch)
    b[2] = rng.integers(0, NI, len(batch[0])).astype(np.int64)
    return tuple(b)


def test_scoring_matches_oracle_on_sampled_rows_and_is_permutation_invariant(setup):
    import bench
    cfg, icl, batch, params, fresh = setup
    rng = np.random.default_rng(5)
    tb = _as_test(batch, rng, bench.NI)
    model = fresh()
    db = model.stage_batch(tb, is_test=True)
    lg, _ = model.score_staged(db, 2)
    lg = lg.cpu().numpy()
    rows = rng.choice(B, 384, replace=False)
    sub = tuple(np.asarray(f)[rows] for f in tb)
    r1, _ = O.forward_logits(params, icl, sub, 1, config=cfg)
    r2, _ = O.forward_logits(params, icl, sub, 2, config=cfg)
    assert rel_err(lg[rows, 0], r1) < 1e-4 and rel_err(lg[rows, 1], r2) < 1e-4
    assert round(float(np.mean(lg[rows, 0] - lg[rows, 1] > 0)), 4) == round(float(np.mean(r1 - r2 > 0)), 4)
    perm = rng.permutation(B)
    pb = tuple(np.asarray(f)[perm] for f in tb)
    lgp, _ = model.score_staged(model.stage_batch(pb, is_test=True), 2)
    assert np.array_equal(lgp.cpu().numpy(), lg[perm])            # rows are independent: bit-exact


def test_train_step_is_deterministic_and_consistent_with_scoring(setup):
    cfg, icl, batch, params, fresh = setup
    outs = []
    for _ in range(2):
        m = fresh()
        db = m.stage_batch(batch)
        # loss of the step must equal mean BCE of the forward logits + reg * l2 of the tables
        tb = list(batch); tb[2] = batch[1]
        lg, _ = m.score_staged(m.stage_batch(tuple(tb), is_test=True), 1)
        x = lg[:, 0].double().cpu()
        y = torch.as_tensor(np.asarray(batch[2], np.float64))
        bce = float(torch.mean(torch.clamp(x, min=0) - x * y + torch.log1p(torch.exp(-x.abs()))))
        l2 = 0.5 * sum(float(torch.sum(t.double() ** 2)) for t in (m.user_emb, m.item_emb, m.cate_emb, m.usert_emb))
        item_b0 = m.item_b.double().sum().item()
        g_sum = float(torch.sum(torch.sigmoid(x) - y)) / B
        stats = m.train_staged(db, 1.0).cpu().numpy()
        assert abs(stats[1] - bce) / bce < 1e-5
        assert abs(stats[0] - (bce + cfg["regulation_rate"] * l2)) / stats[0] < 1e-5
        assert stats[3] == 1.0                                        # clip inactive
        # conservation through sort + segmented reduce: sum of item_b updates = -lr * sum_b dL/dlogit_b
        d_item_b = m.item_b.double().sum().item() - item_b0
        # floor: g_sum is a cancelling sum (sigmoid - y over a balanced batch), while every one of the NI updated
        # biases is rounded to fp32 on its own (0.5 ulp of |item_b| ~ 0.4 each, random walk over NI rows)
        floor = 4 * 2.0 ** -24 * float(m.item_b.abs().max()) * cfg["item_count"] ** 0.5
        assert abs(d_item_b + g_sum) <= 1e-4 * abs(g_sum) + floor
        m.train_staged(db, 1.0)
        outs.append(({k: v.numpy().copy() for k, v in m.state_dict().items()}, stats.copy()))
    assert np.array_equal(outs[0][1], outs[1][1])
    for k in outs[0][0]:
        assert np.array_equal(outs[0][0][k], outs[1][0][k]), k


def test_train_step_updates_sampled_rows_like_the_oracle(setup):
    """Rows of user_emb touched by exactly one sample, and untouched rows (pure L2 decay), against
    closed forms; dense parameters against the oracle on a 2048-row sub-batch of the same data."""
    import bench
    cfg, icl, batch, params, fresh = setup
    m = fresh()
    before = m.state_dict()
    m.train_staged(m.stage_batch(batch), 1.0)
    after = m.state_dict()
    u = np.asarray(batch[0])
    untouched = np.setdiff1d(np.arange(bench.NU), u)[:500]
    reg = cfg["regulation_rate"]
    exp = before["user_emb"].numpy()[untouched] * np.float32(1.0 - reg)
    assert np.max(np.abs(after["user_emb"].numpy()[untouched] - exp)) < 1e-7
    hist_items = np.unique(np.concatenate([np.asarray(batch[3]).ravel(), np.asarray(batch[4]).ravel(), batch[1]]))
    cold = np.setdiff1d(np.arange(bench.NI), hist_items)
    if len(cold):
        exp = before["item_emb"].numpy()[cold] * np.float32(1.0 - reg)
        assert np.max(np.abs(after["item_emb"].numpy()[cold] - exp)) < 1e-7
    # small-batch oracle parity at this table shape (B = 2048)
    sub = tuple(np.asarray(f)[:2048] for f in batch)
    ref = O.train_step(params, icl, sub, 1.0, cfg, dtype=torch.float64)
    m2 = fresh()
    loss = m2.train(None, sub, 1.0)
    assert abs(loss - ref["loss"]) / ref["loss"] < 1e-4
    sd = m2.state_dict()
    for k, v in ref["new_params"].items():
        step = np.asarray(params[k], np.float64) - v
        assert np.max(np.abs(sd[k].numpy() - v)) <= 1e-4 * np.max(np.abs(step)) + 2e-7 * np.max(np.abs(v)), k
