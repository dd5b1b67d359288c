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


def build_dataset_gpu(reviewer, asin, day, item_cate, item_count, seed=1234, device=None, with_gaps=False):
    """The same dataset, built ON THE GPU and left there: returns (train, test) ``DeviceDataset`` objects whose rows are
    the reference's samples in the reference's (shuffled) order -- ready for ``DeviceDataset.batch`` / ``tlsan_collate``.

    Session segmentation, the split rule (build_dataset.py:55-72), u_cate (:54), the time-gap weights (:16-21) and
    the scatter of every sample into the CSR image run as kernels (csrc/tlsan_builder.cu).  The host keeps only what
    consumes the Python ``random`` stream, in the reference's call order: negative sampling (:28-33), the choice of
    the test item (:66) and the two shuffles (:75-76).  ``with_gaps=True`` also returns the raw integer day gaps of
    every history entry (device int32, same CSR offsets) for the fused-bucketing input path."""
    import ctypes as C
    import torch
    from . import _lib
    from .dataset import DeviceDataset
    lib = _lib.lib()
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    reviewer = np.ascontiguousarray(reviewer, np.int32); asin = np.ascontiguousarray(asin, np.int32)
    day = np.ascontiguousarray(day, np.int32); item_cate = np.ascontiguousarray(item_cate, np.int32)
    N = len(reviewer)
    starts = np.flatnonzero(np.r_[True, reviewer[1:] != reviewer[:-1]])
    user_off = np.r_[starts, N].astype(np.int64)
    nu = len(starts)
    up = lambda a: torch.from_numpy(a).to(dev)
    d_rev, d_asin, d_day, d_cate, d_off = up(reviewer), up(asin), up(day), up(item_cate), up(user_off)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    counts = torch.empty(nu, 4, dtype=torch.int32, device=dev)
    test2 = torch.empty(nu, 2, dtype=torch.int32, device=dev)
    _lib.check(lib.tlsan_ds_plan(d_day.data_ptr(), d_off.data_ptr(), nu, counts.data_ptr(), test2.data_ptr(), st))
    c, t2 = counts.cpu().numpy(), test2.cpu().numpy()            # per USER, a few bytes each
    pairs, has_test = c[:, 0].astype(np.int64), c[:, 1].astype(np.int64)
    first_tr = np.r_[0, np.cumsum(2 * pairs)].astype(np.int64)
    first_te = np.r_[0, np.cumsum(has_test)].astype(np.int64)
    ntr, nte = int(first_tr[-1]), int(first_te[-1])
    # ---- the Python random stream, reference call order: per user gen_neg for every position, then the test choice
    rnd = random.Random(seed)
    neg = np.empty(N, np.int32)
    pick = np.zeros(nu, np.int32)
    for u in range(nu):
        s, e = int(user_off[u]), int(user_off[u + 1])
        pos_list = asin[s:e].tolist()
        pos_set = set(pos_list)
        for k in range(e - s):
            n_ = pos_list[0]
            while n_ in pos_set:
                n_ = rnd.randint(0, item_count - 1)
            neg[s + k] = n_
        if has_test[u] and t2[u, 1] > 1:
            pick[u] = rnd.choice(range(int(t2[u, 1])))          # same draw as random.choice(new_session)
    perm_tr = list(range(ntr)); rnd.shuffle(perm_tr)
    perm_te = list(range(nte)); rnd.shuffle(perm_te)
    perm_tr = np.asarray(perm_tr, np.int64); perm_te = np.asarray(perm_te, np.int64)
    pos_tr = np.empty(ntr, np.int64); pos_tr[perm_tr] = np.arange(ntr)
    pos_te = np.empty(nte, np.int64); pos_te[perm_te] = np.arange(nte)
    # ---- sizes of every sample -> CSR offsets in the final order
    d_ftr, d_fte = up(first_tr), up(first_te)
    lp_tr = torch.zeros(max(ntr, 1), dtype=torch.int32, device=dev); ln_tr = torch.zeros_like(lp_tr)
    lp_te = torch.zeros(max(nte, 1), dtype=torch.int32, device=dev); ln_te = torch.zeros_like(lp_te)
    _lib.check(lib.tlsan_ds_lengths(d_day.data_ptr(), d_off.data_ptr(), nu, d_ftr.data_ptr(), d_fte.data_ptr(),
                                    lp_tr.data_ptr(), ln_tr.data_ptr(), lp_te.data_ptr(), ln_te.data_ptr(), st))

    def offsets(lens, perm):
        off = torch.zeros(len(perm) + 1, dtype=torch.int64, device=dev)
        if len(perm):
            off[1:] = torch.cumsum(lens[up(perm)].to(torch.int64), 0)
        return off

    def alloc(n, pre_off, new_off, is_test):
        npre, nnew = int(pre_off[-1].item()), int(new_off[-1].item())
        t = dict(uid=torch.empty(n, dtype=torch.int32, device=dev), pre_off=pre_off,
                 pre_items=torch.empty(max(npre, 1), dtype=torch.int32, device=dev),
                 pre_time=torch.empty(max(npre, 1), dtype=torch.float32, device=dev), new_off=new_off,
                 new_items=torch.empty(max(nnew, 1), dtype=torch.int32, device=dev),
                 cand=torch.empty(n, dtype=torch.int32, device=dev), ucate=torch.empty(n, dtype=torch.int32, device=dev))
        t["second_i" if is_test else "second_f"] = torch.empty(n, dtype=torch.int32 if is_test else torch.float32, device=dev)
        gap = torch.empty(max(npre, 1), dtype=torch.int32, device=dev) if with_gaps else None
        return t, gap
    tr, gap_tr = alloc(ntr, offsets(lp_tr, perm_tr), offsets(ln_tr, perm_tr), False)
    te, gap_te = alloc(nte, offsets(lp_te, perm_te), offsets(ln_te, perm_te), True)

    def cstruct(t, n):
        p = lambda k: t[k].data_ptr() if k in t else None
        return _lib.Dataset(uid=p("uid"), pre_off=p("pre_off"), pre_items=p("pre_items"), pre_time=p("pre_time"),
                            new_off=p("new_off"), new_items=p("new_items"), cand=p("cand"), second_i=p("second_i"),
                            second_f=p("second_f"), ucate=p("ucate"), n=n)
    lut = np.zeros(13, np.float32)
    for n_ in range(1, 13):
        lut[n_] = np.float32(1 / np.float64(n_))                  # float64 1/n stored as float32 (input.py:36,45)
    d_lut, d_neg, d_pick, d_ptr, d_pte = up(lut), up(neg), up(pick), up(pos_tr), up(pos_te)
    ctr, cte = cstruct(tr, ntr), cstruct(te, nte)
    _lib.check(lib.tlsan_ds_emit(d_rev.data_ptr(), d_asin.data_ptr(), d_day.data_ptr(), d_cate.data_ptr(), d_off.data_ptr(),
                                 nu, d_ftr.data_ptr(), d_fte.data_ptr(), d_ptr.data_ptr(), d_pte.data_ptr(),
                                 d_neg.data_ptr(), d_pick.data_ptr(), d_lut.data_ptr(), C.byref(ctr), C.byref(cte),
                                 gap_tr.data_ptr() if with_gaps else None, gap_te.data_ptr() if with_gaps else None, st))
    torch.cuda.synchronize(dev)
    train, test = DeviceDataset.from_device(tr, False, dev), DeviceDataset.from_device(te, True, dev)
    if with_gaps:
        return train, test, gap_tr, gap_te
    return train, test
