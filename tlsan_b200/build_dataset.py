"""Dataset builder of the reference (TLSAN/build_dataset.py), vectorised  (SURVEY.md section 8f-3).

Per user: sessions = runs of equal review day (build_dataset.py:38-46); every non-first session
yields two train samples (positive / sampled negative) while ``i + count < min(len, 90) - 1``
(:55-62), else ONE test sample and the user is done (:63-72); long-term time weights are
``1 / n`` with ``n = sum_j [d >= 2^j]``, ``d = cur_day - day + 1`` (:16-21); ``u_cate`` is the most
frequent category so far (:54).  Negative sampling, the test-item choice and the final shuffles use
the Python ``random`` stream exactly like the reference (seed 1234, same call order), so the output
is IDENTICAL to the reference script's ``dataset.pkl`` -- ``tests/test_build_dataset.py`` checks it
against the recorded output of the unmodified script.

What changes: the O(N * NI) ``meta_df[meta_df['asin'] == item]`` scan per item (:47; 22 s on
Digital-Music, hours on Movies-TV) becomes an array look-up, and the day gaps can be bucketed on the
GPU (``tlsan_time_bucket``) -- ``return_gaps=True`` exposes the raw integer gaps ``d``.
"""
import random

import numpy as np

MAX_LENGTH = 90                                                    # build_dataset.py:7
GAP = np.array([2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096])   # build_dataset.py:16


def time_weights(days, cur):
    """proc_time_emb (build_dataset.py:18-21) for one history: float64 1/n per entry + the gaps d."""
    d = cur - np.asarray(days, np.int64) + 1
    n = (d[:, None] >= GAP[None, :]).sum(1)
    return [1 / np.sum(x >= GAP) for x in d], d, n


def _dominant(cates):
    """pd.value_counts(pre_cates).index[0] (build_dataset.py:54): highest count, ties -> the
    category that reached the list first (pandas keeps first-appearance order among equal counts)."""
    best, best_n, seen = None, 0, {}
    for c in cates:
        seen[c] = seen.get(c, 0) + 1
    for c, n in seen.items():                                       # dict preserves first appearance
        if n > best_n:
            best, best_n = c, n
    return best


def build_dataset(reviewer, asin, day, item_cate, item_count, seed=1234, return_gaps=False):
    """reviewer / asin / day: the three columns of ``reviews_df`` (sorted by reviewer, then time, as
    utils/2_remap_id.py:91 leaves them).  Returns (train_set, test_set) in the reference tuple layout
    (build_dataset.py:58-59,71); with return_gaps also the list of integer day-gap arrays per sample."""
    rnd = random.Random(seed)                                       # random.seed(1234), build_dataset.py:8
    reviewer = np.asarray(reviewer); asin = np.asarray(asin); day = np.asarray(day)
    item_cate = np.asarray(item_cate)
    starts = np.flatnonzero(np.r_[True, reviewer[1:] != reviewer[:-1]])
    ends = np.r_[starts[1:], len(reviewer)]
    train_set, test_set, train_gaps, test_gaps = [], [], [], []
    for s, e in zip(starts, ends):                                  # groupby('reviewerID'), ascending ids
        uid = int(reviewer[s])
        pos_list = asin[s:e].tolist()
        tim = day[s:e]
        pos_set = set(pos_list)
        neg_list = []
        for _ in pos_list:                                          # gen_neg, build_dataset.py:28-33
            neg = pos_list[0]
            while neg in pos_set:
                neg = rnd.randint(0, item_count - 1)
            neg_list.append(neg)
        valid_length = min(len(pos_list), MAX_LENGTH)
        # sessions: the reference walks sorted unique days and counts them; tim is sorted per user
        cuts = np.flatnonzero(np.r_[True, tim[1:] != tim[:-1]])
        bounds = np.r_[cuts, len(tim)]
        cates = item_cate[asin[s:e]].tolist()
        i = int(bounds[1]) if len(bounds) > 1 else len(pos_list)   # first session goes to the history
        for k in range(1, len(bounds) - 1):
            count = int(bounds[k + 1] - bounds[k])
            now_cate = _dominant(cates[:i])
            new_session = pos_list[i:i + count]
            if i + count < valid_length - 1:
                emb, d, _ = time_weights(tim[:i], int(tim[i]))
                pre = pos_list[:i]
                train_set.append((uid, pre, new_session, emb, pos_list[i + count], 1, now_cate))
                train_set.append((uid, pre, new_session, emb, neg_list[i + count], 0, now_cate))
                train_gaps.append(d); train_gaps.append(d)
                i += count
            else:
                pos_item = pos_list[i]
                if count > 1:
                    pos_item = rnd.choice(new_session)
                    new_session.remove(pos_item)
                neg = neg_list[pos_list.index(pos_item)]
                emb, d, _ = time_weights(tim[:i], int(tim[i]))
                test_set.append((uid, pos_list[:i], new_session, emb, (pos_item, neg), now_cate))
                test_gaps.append(d)
                break
    # random.shuffle(train_set); random.shuffle(test_set)  (build_dataset.py:75-76) -- same stream
    perm_tr = list(range(len(train_set))); rnd.shuffle(perm_tr)
    perm_te = list(range(len(test_set))); rnd.shuffle(perm_te)
    train_set = [train_set[j] for j in perm_tr]; test_set = [test_set[j] for j in perm_te]
    if return_gaps:
        return train_set, test_set, [train_gaps[j] for j in perm_tr], [test_gaps[j] for j in perm_te]
    return train_set, test_set
