"""The reference driver loop (TLSAN/train.py:180-245) on tlsan_b200 -- Digital-Music, reference defaults:
batch 32 / 128, Ls 10, SGD lr 1.0, L2 5e-5, clip 5, 20 epochs, eval every 1 000 steps, seed 1234.

    python examples/train_digital_music.py [--epochs 20] [--no-topk]

Only the TensorFlow session / flag / summary lines of train.py are gone; the loop, the batcher classes and the
Model calls are the reference's.  The dataset is the output of the unmodified reference build_dataset.py as
recorded in tests/golden/digital_music.npz (oracle/make_golden.py).  README anchor: best test AUC 0.9753
(README.md:35); the CPU restatement of the TF graph reached 0.9693 with its own glorot stream."""
import argparse
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tlsan_b200 import DataInput, DataInputTest          # noqa: E402  (reference: from input import ...)
from tlsan_b200.model import Model                         # noqa: E402  (reference: from model import Model)
from tests.util import load_digital_music                  # noqa: E402


def eval_auc(test_set, model, bs, Ls):                     # train.py:86-96
    auc_sum = 0.0
    for _, batch in DataInputTest(test_set, bs, Ls):
        auc_sum += model.eval_auc(None, batch) * len(batch[0])
    return auc_sum / len(test_set)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--train-batch-size", type=int, default=32)
    ap.add_argument("--test-batch-size", type=int, default=128)
    ap.add_argument("--eval-freq", type=int, default=1000)
    ap.add_argument("--no-topk", action="store_true", help="skip the P@k / R@k passes at every evaluation")
    ap.add_argument("--seed", type=int, default=1234, help="shuffle + initialisation seed (train.py:176-178 uses 1234)")
    args = ap.parse_args()
    random.seed(args.seed)
    np.random.seed(args.seed)
    dm = load_digital_music()
    user_count, item_count, cate_count = dm.counts
    train_set, test_set = list(dm.train_set), list(dm.test_set)
    config = {"hidden_units": 64, "num_blocks": 1, "num_heads": 8, "Ls": 10, "dropout": 0.0, "regulation_rate": 5e-5,
              "itemid_embedding_size": 32, "userid_embedding_size": 32, "cateid_embedding_size": 32,
              "optimizer": "sgd", "max_gradient_norm": 5.0, "model_dir": "save_path",
              "user_count": user_count, "item_count": item_count, "cate_count": cate_count}
    model = Model(config, dm.icl, seed=args.seed)
    print("Init finish.\tCost time: 0.00s\tInit AUC: %.4f" % eval_auc(test_set, model, args.test_batch_size, 10), flush=True)
    lr, best_auc, avg_loss, start = 1.0, 0.0, 0.0, time.time()
    best_prec, best_recall, curve = [0.0] * 6, [0.0] * 6, []
    for _ in range(args.epochs):
        random.shuffle(train_set)
        for _, batch in DataInput(train_set, args.train_batch_size, config["Ls"]):
            avg_loss += model.train(None, batch, lr, False)
            step = model.global_step.eval()
            if step % args.eval_freq == 0:
                test_auc = eval_auc(test_set, model, args.test_batch_size, config["Ls"])
                curve.append((step, round(time.time() - start, 2), round(float(test_auc), 4)))
                line = "Epoch %d Global_step %d\tTrain_loss: %.4f\tEval_auc: %.4f" % (
                    model.global_epoch_step.eval(), step, avg_loss / args.eval_freq, test_auc)
                if not args.no_topk:
                    for _, tb in DataInputTest(test_set, args.test_batch_size, config["Ls"]):
                        model.eval_prec(None, tb)
                    prec = [getattr(model, "prec_%d" % k).eval() for k in (1, 10, 20, 30, 40, 50)]
                    for _, tb in DataInputTest(test_set, args.test_batch_size, config["Ls"]):
                        model.eval_recall(None, tb)
                    recall = [getattr(model, "recall_%d" % k).eval() for k in (1, 10, 20, 30, 40, 50)]
                    line += "\tP@10 %.4f R@10 %.4f R@50 %.4f" % (prec[1], recall[1], recall[5])
                    if step > 20000:
                        best_prec = [max(a, b) for a, b in zip(best_prec, prec)]
                        best_recall = [max(a, b) for a, b in zip(best_recall, recall)]
                print(line, flush=True)
                avg_loss = 0.0
                if test_auc > 0.8 and test_auc > best_auc:
                    best_auc = test_auc
            if step == 150000:
                lr = 0.1
        print("Epoch %d DONE\tCost time: %.2f" % (model.global_epoch_step.eval(), time.time() - start), flush=True)
        model.global_epoch_step_op.eval()
    wall = time.time() - start
    print(json.dumps({"seed": args.seed, "best_test_auc": round(float(best_auc), 4), "readme_auc": 0.9753, "steps": model.global_step.eval(),
                      "wall_s": round(wall, 1), "train_samples_per_s_incl_eval": round(model.global_step.eval() * args.train_batch_size / wall),
                      "best_recall_at_1_10_20_30_40_50": [round(float(x), 4) for x in best_recall], "curve": curve}))


if __name__ == "__main__":
    main()
