"""Synthetic workloads of SURVEY.md section 8d (BASELINE.json configs[1..4]) in the TLSAN/input.py batch layout.

Shapes: Electronics (NU 39 991, NI 22 048, NC 673), Movies-TV (NU 35 896, NI 28 589, NC 15), 10 M-item table
(NU 40 000, NI 10 000 000, NC 673).  Lengths follow the Digital-Music empirical laws; ids are uniform (the
worst case for caches and for duplicate-row reduction).  Shared by bench.py, tools/ and tests/."""
import numpy as np

WORKLOADS = {"electronics": ("TLSAN Electronics-shape synthetic", 39991, 22048, 673),
             "movies": ("TLSAN Movies-TV-shape synthetic", 35896, 28589, 15),
             "items10m": ("TLSAN 10M-item synthetic", 40000, 10000000, 673)}
# Digital-Music empirical laws: P(min(len,10) = k), k = 1..10 ; short length pmf
P_LONG = np.array([7.0, 6.9, 6.7, 6.5, 6.3, 5.9, 5.2, 4.5, 3.9, 47.1]) / 100.0
P_SHORT_HEAD = np.array([.8724, .0886, .0223, .0085, .0040])
S_MAX = 18


def synth_batches(rng, n_batches, B, L, NU, NI, NC, full=False, is_test=False, raw_gaps=False):
    """`n_batches` batches (9-tuples of TLSAN/input.py:54 / :107).  full=True: every history has L entries (the
    roofline variant of the scoring sweep).  raw_gaps=True: element 5 is the int32 day-gap matrix d[B,L]
    (0 = padding) instead of the float32 weights 1/n(d) (build_dataset.py:16-21)."""
    p_long = P_LONG / P_LONG.sum()
    tail = np.full(S_MAX - 5, (1.0 - P_SHORT_HEAD.sum()) / (S_MAX - 5))
    p_short = np.concatenate([P_SHORT_HEAD, tail])
    p_short /= p_short.sum()
    out = []
    for _ in range(n_batches):
        if full:
            sl = np.full(B, L, np.int64)
        else:
            frac = rng.choice(10, B, p=p_long) + 1                           # law of min(len, 10)
            sl = np.maximum(1, np.round(frac * (L / 10.0))).astype(np.int64) if L != 10 else frac.astype(np.int64)
        new_sl = (rng.choice(S_MAX, B, p=p_short) + 1).astype(np.int64)
        S = int(new_sl.max())
        hist_i = rng.integers(0, NI, (B, L)).astype(np.int64)
        hist_i_new = rng.integers(0, NI, (B, S)).astype(np.int64)
        col = np.arange(L)[None, :]
        if raw_gaps:
            # day gaps >= 2, non-increasing in t (older entries first), log-uniform over the 12 buckets
            e = np.sort(rng.uniform(1.0, 12.9, (B, L)), axis=1)[:, ::-1]
            hist_t = np.maximum(2, np.floor(2.0 ** e)).astype(np.int32)
        else:
            n = np.sort(rng.integers(1, 13, (B, L)), axis=1)[:, ::-1]        # bucket non-increasing in t
            hist_t = (1.0 / n).astype(np.float32)
        hist_i[col >= sl[:, None]] = 0
        hist_t[col >= sl[:, None]] = 0
        hist_i_new[np.arange(S)[None, :] >= new_sl[:, None]] = 0
        third = rng.integers(0, NI, B).astype(np.int64) if is_test else rng.integers(0, 2, B).astype(np.int64)
        out.append((rng.integers(0, NU, B).astype(np.int64), rng.integers(0, NI, B).astype(np.int64),
                    third, hist_i, hist_i_new, hist_t, sl, new_sl, rng.integers(0, NC, B).astype(np.int64)))
    return out


def algorithmic_bytes(batch, L):
    """SURVEY.md 8d byte model, evaluated on the actual lengths of `batch` (totals per batch).
    Per sample: R = 2(l+s)+4 embedding rows of 128 B.  The per-kernel figures split the train-step formula by
    which kernel touches what; scratch traffic, sort traffic and the backward's re-read of
    the token rows are never credited."""
    sl = np.asarray(batch[6], np.int64); s = np.asarray(batch[7], np.int64)
    S = np.shape(batch[4])[1]
    R = 2 * (sl + s) + 4
    scoring1 = 4 * (2 * L + S + 6) + 4 * (sl + s + 1) + 128 * R + (4 * L + 4) + 4
    scoring2 = scoring1 + 272
    train = scoring1 + 2 * (128 * R + 4 * sl + 4) + 4
    long_fwd = 4 * (2 * L + 2) + 4 * sl + 128 * 2 * sl + 4 * L                 # ids, hist_t, icl, rows, usert row
    short = 4 * (S + 6) + 4 * (s + 1) + 128 * (2 * s + 4) + 4 + 128 * (2 * s + 4) + 8   # reads + gradient rows written
    bwd_long = 128 * 2 * sl + 4 * sl                                            # gradient rows + d usert written
    reduce_ = 128 * R + 4 * sl + 4                                              # every gradient row read once
    return {"scoring1": int(scoring1.sum()), "scoring2": int(scoring2.sum()), "train": int(train.sum()),
            "long_fwd": int(long_fwd.sum()), "short": int(short.sum()), "bwd_long": int(bwd_long.sum()),
            "reduce": int(reduce_.sum())}


def table_bytes(L, NU, NI, NC):
    return 4 * (33 * NI + 32 * NU + L * NU + 32 * NC)
