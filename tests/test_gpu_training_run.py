"""End-to-end anchor: the reference driver loop (TLSAN/train.py:190-199, batch 32, lr 1.0) on the recorded
Digital-Music dataset must learn -- loss falls and the test AUC climbs well above the untrained model's
(README anchor after 20 epochs: 0.9753; examples/train_digital_music.py reaches 0.9687, the CPU restatement 0.9693)."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_thousand_reference_steps_learn():
    from oracle import tlsan_oracle as O
    from tests.util import load_digital_music
    from tlsan_b200 import DataInput, DataInputTest
    from tlsan_b200.model import Model
    dm = load_digital_music()
    cfg = O.default_config(*dm.counts)
    model = Model(cfg, dm.icl, seed=1234)
    train_set, test_set = list(dm.train_set), list(dm.test_set)

    def auc():
        s = 0.0
        for _, b in DataInputTest(test_set, 128, 10):
            s += model.eval_auc(None, b) * len(b[0])
        return s / len(test_set)
    auc0 = auc()
    random.seed(1234)
    losses = []
    while model.global_step.eval() < 2000:
        random.shuffle(train_set)
        for _, batch in DataInput(train_set, 32, 10):
            losses.append(model.train(None, batch, 1.0))
            if model.global_step.eval() >= 2000:
                break
    auc1 = auc()
    assert np.isfinite(losses).all()
    assert np.mean(losses[-200:]) < np.mean(losses[:200]) - 0.05
    assert auc1 > 0.88 and auc1 > auc0 - 0.02, (auc0, auc1)
