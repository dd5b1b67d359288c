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


# lr per optimizer: of the order the reference would be run with (tf defaults 1e-3; adadelta steps are tiny at lr 1)
@pytest.mark.parametrize("opt,lr", [("adam", 1e-3), ("rmsprop", 1e-3), ("adadelta", 1.0)])
def test_optimizer_steps_match_oracle(dm, opt, lr):
    cfg = O.default_config(*dm.counts, optimizer=opt)
    params = O.randomize_params(O.init_params(cfg), 7)
    model = model_from_params(params, dm.icl, cfg)
    state = O.init_opt_state(params, opt)
    ref_params = params
    for step in range(3):
        batch = O.collate_train(dm.train_set[step * 64:(step + 1) * 64], 10)
        ref = O.train_step(ref_params, dm.icl, batch, lr, cfg, dtype=torch.float64, opt_state=state)
        loss = model.train(None, batch, lr)
        assert abs(loss - ref["loss"]) <= 1e-4 * abs(ref["loss"])
        sd = model.state_dict()
        slots = model.slot_views()
        for k, v in ref["new_params"].items():
            got = sd[k].numpy().astype(np.float64)
            delta = np.asarray(params[k], np.float64) - v                      # total movement since step 0
            # element-wise on the accumulated update; floor = 1e-5 of the tensor's largest movement + fp32 rounding of w
            bound = 1e-4 * np.abs(delta) + 1e-5 * np.max(np.abs(delta)) + 3e-7 * np.abs(v) + 1e-12
            assert np.all(np.abs(got - v) <= bound), (opt, step, k, float(np.max(np.abs(got - v) / bound)))
            for s_, ref_slot in enumerate((state["s1"][k], state["s2"][k])):
                g, r = slots[k][s_].cpu().numpy().astype(np.float64), ref_slot.numpy()
                assert np.all(np.abs(g - r) <= 2e-4 * np.abs(r) + 1e-5 * np.max(np.abs(r)) + 1e-30), (opt, step, k, s_)
        ref_params = ref["new_params"]
    # item_b: rows outside the batch's IndexedSlices keep weight and slots under the sparse rmsprop / adadelta kernels
    if opt != "adam":
        seen = np.zeros(cfg["item_count"], bool)
        for step in range(3):
            seen[np.asarray(O.collate_train(dm.train_set[step * 64:(step + 1) * 64], 10)[1])] = True
        assert np.array_equal(model.state_dict()["item_b"].numpy()[~seen], np.asarray(params["item_b"])[~seen])
        s1 = model.slot_views()["item_b"][0].cpu().numpy()
        assert np.all(s1[~seen] == (1.0 if opt == "rmsprop" else 0.0))


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
