"""init_optimizer's other branches (TLSAN/model.py:188-193): adam / rmsprop / adadelta, fused into the apply kernels
(tlsan_apply_flat_opt) against the fp64 oracle restatement of TF 1.8's apply ops, several steps so the slot variables
(m / v, rms / momentum, accum / accum_update) feed back into the weights."""
import numpy as np
import pytest
import torch

from oracle import tlsan_oracle as O
from tests.util import load_digital_music, model_from_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dm():
    return load_digital_music()


def _sync(model, params, state):
    """Put the oracle's weights, slot variables and step count into the GPU model."""
    model.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in params.items()})
    slots = model.slot_views()
    for k in params:
        slots[k][0].copy_(state["s1"][k].to(torch.float32)); slots[k][1].copy_(state["s2"][k].to(torch.float32))
    model._opt.step = int(state["t"])


# lr per optimizer: of the order the reference would be run with (tf defaults 1e-3; adadelta steps are tiny at lr 1)
@pytest.mark.parametrize("opt,lr", [("adam", 1e-3), ("rmsprop", 1e-3), ("adadelta", 1.0)])
def test_optimizer_steps_match_oracle(dm, opt, lr):
    """Each step starts from the oracle's state (weights, slots, step count) and must land inside the interval the
    ORACLE's apply op spans when its aggregated gradient moves by the fp32 tolerance the sgd tests grant
    (|dg| <= 1e-4 |g| + 1e-5 max|g|): adam's m / (sqrt(v) + eps) and adadelta's ratio of roots amplify a relative
    gradient error without bound where g ~ 0, so a fixed tolerance on the weights would test the conditioning of the
    update rule, not the kernel."""
    cfg = O.default_config(*dm.counts, optimizer=opt)
    params = O.randomize_params(O.init_params(cfg), 7)
    model = model_from_params(params, dm.icl, cfg)
    state = O.init_opt_state(params, opt)
    seen = np.zeros(cfg["item_count"], bool)
    for step in range(3):
        batch = O.collate_train(dm.train_set[step * 64:(step + 1) * 64], 10)
        seen[np.asarray(batch[1])] = True
        _sync(model, params, state)
        before = dict(s1={k: v.clone() for k, v in state["s1"].items()}, s2={k: v.clone() for k, v in state["s2"].items()})
        ref = O.train_step(params, dm.icl, batch, lr, cfg, dtype=torch.float64, opt_state=state)
        loss = model.train(None, batch, lr)
        assert abs(loss - ref["loss"]) <= 1e-4 * abs(ref["loss"])
        sd, slots = model.state_dict(), model.slot_views()
        touched = torch.zeros(cfg["item_count"], dtype=torch.bool); touched[torch.as_tensor(np.asarray(batch[1]))] = True
        gmax = {k: float(np.max(np.abs(v))) * ref["scale"] for k, v in ref["grads"].items()}
        for k in ref["new_params"]:
            g = torch.as_tensor(ref["grads"][k]) * ref["scale"]
            # the gradient of the second map's bias is identically 0 (softmax over the sequence is shift invariant):
            # the kernels return rounding noise of the size of the sibling kernel's gradient there
            sib = k.replace("/bias", "/W")
            gm = max(gmax[k], gmax.get(sib, 0.0)) if k.endswith("bn_dense_map2/linear_map/bias") else gmax[k]
            tol = 1e-4 * g.abs() + 1e-5 * gm
            w0 = torch.as_tensor(np.asarray(params[k], np.float64))
            args = (lr, before["s1"][k], before["s2"][k], state["t"], touched if k == "item_b" else None)
            mid = O._opt_update(opt, w0, g, *args)
            lo, hi = O._opt_update(opt, w0, g - tol, *args), O._opt_update(opt, w0, g + tol, *args)
            got = (sd[k].double(), slots[k][0].cpu().double(), slots[k][1].cpu().double())
            for what, m_, l_, h_, g_ in zip(("w", "slot1", "slot2"), mid, lo, hi, got):
                span = torch.maximum((l_ - m_).abs(), (h_ - m_).abs())
                moved = (m_ - w0).abs() if what == "w" else m_.abs()
                bound = 1.5 * span + 3e-7 * m_.abs() + 2e-6 * moved + 1e-30
                bad = (g_ - m_).abs() > bound
                if k == "item_b" and opt != "adam":
                    # a touched row whose aggregated gradient is below the tolerance may be read as untouched
                    bad &= ~((g.abs() <= tol) & touched)
                assert not bool(bad.any()), (opt, step, k, what, float(((g_ - m_).abs() / bound).max()))
        params = ref["new_params"]
    # item_b: rows outside every batch's IndexedSlices keep weight and slots under the sparse rmsprop / adadelta kernels
    if opt != "adam":
        p0 = O.randomize_params(O.init_params(cfg), 7)
        assert np.array_equal(model.state_dict()["item_b"].numpy()[~seen], np.asarray(p0["item_b"])[~seen])
        s1 = model.slot_views()["item_b"][0].cpu().numpy()
        assert np.all(s1[~seen] == (1.0 if opt == "rmsprop" else 0.0))


def test_optimizer_trajectory_stays_close(dm):
    """Free-running: 5 adam steps without re-synchronising; the loss trajectory follows the oracle's."""
    cfg = O.default_config(*dm.counts, optimizer="adam")
    params = O.randomize_params(O.init_params(cfg), 7)
    model = model_from_params(params, dm.icl, cfg)
    state = O.init_opt_state(params, "adam")
    for step in range(5):
        batch = O.collate_train(dm.train_set[step * 64:(step + 1) * 64], 10)
        ref = O.train_step(params, dm.icl, batch, 1e-3, cfg, dtype=torch.float64, opt_state=state)
        loss = model.train(None, batch, 1e-3)
        assert abs(loss - ref["loss"]) <= 2e-4 * abs(ref["loss"]), (step, loss, ref["loss"])
        params = ref["new_params"]


def test_optimizer_checkpoint_roundtrip(dm, tmp_path):
    cfg = O.default_config(*dm.counts, optimizer="adam", model_dir=str(tmp_path))
    params = O.randomize_params(O.init_params(cfg), 3)
    a = model_from_params(params, dm.icl, cfg)
    batch = O.collate_train(dm.train_set[:64], 10)
    a.train(None, batch, 1e-3)
    path = a.save(None)
    b = model_from_params(params, dm.icl, cfg)
    b.restore(None, path)
    a.train(None, batch, 1e-3); b.train(None, batch, 1e-3)
    for k, v in a.state_dict().items():
        assert torch.equal(v, b.state_dict()[k]), k


def test_sgd_is_the_else_branch(dm):
    """Any other string selects GradientDescentOptimizer (model.py:194-195)."""
    cfg = O.default_config(*dm.counts, optimizer="momentum")
    m = model_from_params(O.init_params(cfg), dm.icl, cfg)
    assert m.optimizer == "sgd"
