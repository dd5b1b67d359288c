"""Host-side model of the balanced segmented reduce (csrc/tlsan_update.cu: k_row_reduce_bal + k_row_fix): the range
split, the head / tail rule and the fix-up sum must reproduce plain per-row sums for any sorted key list, and write
every row exactly once.  (The CUDA kernels are checked against the oracle by the -m gpu parity tests; this pins the
plan they implement, including the cases a random batch rarely hits: a segment ending exactly on a range boundary,
one segment covering several whole ranges, empty ranges.)"""
import numpy as np
import pytest


def rr_range(T, NW):
    return max(16, ((T + NW - 1) // NW + 15) & ~15)


def plan_reduce(keys, vals, NW, NR):
    T = len(keys)
    R = rr_range(T, NW)
    g = np.zeros(NR)
    written = np.zeros(NR, int)
    head = np.full(NW, np.nan)
    tail = np.full(NW, np.nan)
    key_of = lambda q: keys[q] if 0 <= q < T else -1
    for gw in range(NW):
        p0 = min(gw * R, T)
        p1 = min(p0 + R, T)
        if p0 >= p1:
            continue
        acc, first = 0.0, True
        head_shared = key_of(p0 - 1) == key_of(p0)
        for p in range(p0, p1, 16):
            cnt = min(16, p1 - p)
            last_batch = p + 16 >= p1
            tail_shared = last_batch and key_of(p + cnt - 1) == key_of(p + cnt)
            for i in range(cnt):
                acc += vals[p + i]
                if key_of(p + i) != key_of(p + i + 1) or (last_batch and i + 1 == cnt):
                    to_head = first and head_shared
                    to_tail = (not to_head) and tail_shared and i == cnt - 1
                    if to_head:
                        head[gw] = acc
                    elif to_tail:
                        tail[gw] = acc
                    else:
                        g[keys[p + i]] = acc
                        written[keys[p + i]] += 1
                    acc, first = 0.0, False
    seg = np.searchsorted(keys, np.arange(NR + 1))
    for r in range(NR):                                   # k_row_fix
        lo, hi = seg[r], seg[r + 1]
        if lo >= hi:
            g[r] = 0.0
            written[r] += 1
            continue
        wa, wb = lo // R, (hi - 1) // R
        if wa == wb:
            continue
        a = tail[wa]
        for w in range(wa + 1, wb + 1):
            a += head[w]
        g[r] = a
        written[r] += 1
    return g, written


@pytest.mark.parametrize("T,NR,NW", [(5000, 50, 16), (20000, 300, 64), (1000, 500, 8), (100, 5, 4), (64, 3, 4), (7, 9, 8)])
def test_balanced_reduce_plan_equals_row_sums(T, NR, NW):
    rng = np.random.default_rng(T + NW)
    keys = np.sort(np.minimum(rng.zipf(1.3, T) - 1, NR - 1))
    vals = rng.standard_normal(T)
    g, written = plan_reduce(keys, vals, NW, NR)
    ref = np.zeros(NR)
    np.add.at(ref, keys, vals)
    assert np.all(written == 1)
    assert np.allclose(g, ref, rtol=0, atol=1e-9)


def test_balanced_reduce_plan_boundary_cases():
    # segments that end exactly on range boundaries (R = 16), one segment covering three whole ranges, a single key
    for keys in (np.repeat([0, 1, 2, 3], 16), np.r_[np.zeros(8, int), np.ones(48, int), np.full(8, 2)], np.zeros(100, int)):
        vals = np.arange(len(keys), dtype=float) + 1
        g, written = plan_reduce(keys, vals, 4, 5)
        ref = np.zeros(5)
        np.add.at(ref, keys, vals)
        assert np.all(written == 1) and np.array_equal(g, ref)
