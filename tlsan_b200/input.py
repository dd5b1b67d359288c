"""Batch layout of the reference batcher (TLSAN/input.py), vectorised.

``DataInput`` / ``DataInputTest`` keep the reference constructor, iterator protocol and the
9-field tuple (TLSAN/input.py:54,107):

    train: (u, i, y,     hist_i, hist_i_new, hist_t, sl, new_sl, c)
    test : (u, i_pos, j, hist_i, hist_i_new, hist_t, sl, new_sl, c)

``hist_i`` / ``hist_t`` hold the LAST ``k`` long-term entries left-aligned and zero padded
(input.py:39-49); ``hist_i_new`` is zero padded to the longest session of the batch
(input.py:33,37,50-51).  dtypes match the reference (int64 / float32); the per-sample
scalars come back as numpy arrays instead of Python lists (they feed identically).

The per-sample Python loops of the reference are replaced by a one-off CSR conversion
(``CsrDataset``) and fancy indexing, so a batch costs O(1) numpy calls.
"""
import numpy as np


class CsrDataset:
    """Samples of TLSAN/build_dataset.py:58-59 (train) / :71 (test) in CSR form."""

    def __init__(self, uid, pre_off, pre_items, pre_time, new_off, new_items, cand, second, ucate, is_test):
        self.uid = np.asarray(uid, np.int64)
        self.pre_off = np.asarray(pre_off, np.int64)
        self.pre_items = np.asarray(pre_items, np.int64)
        self.pre_time = np.asarray(pre_time, np.float32)      # cast of input.py:36,45
        self.new_off = np.asarray(new_off, np.int64)
        self.new_items = np.asarray(new_items, np.int64)
        self.cand = np.asarray(cand, np.int64)                 # t[4] (train) / t[4][0] (test)
        self.second = np.asarray(second)                       # label t[5] (train) / neg item t[4][1] (test)
        self.ucate = np.asarray(ucate, np.int64)
        self.is_test = bool(is_test)

    def __len__(self):
        return len(self.uid)

    @classmethod
    def from_samples(cls, samples, is_test):
        n = len(samples)
        pre_off = np.zeros(n + 1, np.int64)
        new_off = np.zeros(n + 1, np.int64)
        np.cumsum([len(t[1]) for t in samples], out=pre_off[1:])
        np.cumsum([len(t[2]) for t in samples], out=new_off[1:])
        pre_items = np.fromiter((x for t in samples for x in t[1]), np.int64, count=int(pre_off[-1]))
        pre_time = np.fromiter((x for t in samples for x in t[3]), np.float64, count=int(pre_off[-1]))
        new_items = np.fromiter((x for t in samples for x in t[2]), np.int64, count=int(new_off[-1]))
        uid = [t[0] for t in samples]
        if is_test:
            cand = [t[4][0] for t in samples]; second = np.array([t[4][1] for t in samples], np.int64)
            ucate = [t[5] for t in samples]
        else:
            cand = [t[4] for t in samples]; second = np.array([t[5] for t in samples])
            ucate = [t[6] for t in samples]
        return cls(uid, pre_off, pre_items, pre_time.astype(np.float32), new_off, new_items, cand, second, ucate,
                   is_test)

    def collate(self, idx, k, index_dtype=np.int64):
        """Rows ``idx`` in the layout of DataInput.__next__ / DataInputTest.__next__.  ``index_dtype=np.int32``
        emits every integer field as int32 (same values): Model.stage_batch then skips the 64 -> 32 bit narrowing
        pass of the feed (model.py:210-222 casts to int32 anyway)."""
        out = self._collate(idx, k)
        if np.dtype(index_dtype) == np.int64:
            return out
        return tuple(np.ascontiguousarray(f, dtype=index_dtype) if n != 5 and np.issubdtype(np.asarray(f).dtype, np.integer)
                     else f for n, f in enumerate(out))

    def max_new_len(self):
        if not hasattr(self, "_max_new"):
            self._max_new = max(int(np.max(np.diff(self.new_off))) if len(self) else 1, 1)
        return self._max_new

    def _i32(self):
        """int32 mirrors of the id arrays + the id maxima of the whole dataset (one pass, cached)."""
        if not hasattr(self, "_m32"):
            m = {k: np.ascontiguousarray(getattr(self, k), np.int32) for k in ("uid", "pre_items", "new_items", "cand", "ucate")}
            sec = np.asarray(self.second)
            m["second"] = (np.ascontiguousarray(sec, np.int32) if self.is_test
                           else np.ascontiguousarray(sec, np.float32).view(np.int32))
            for k in ("uid", "pre_items", "new_items", "cand", "ucate"):
                if len(m[k]) and int(np.min(getattr(self, k))) < 0:
                    raise IndexError("negative id in dataset field %s" % k)
            items = max(int(self.pre_items.max(initial=0)), int(self.new_items.max(initial=0)), int(self.cand.max(initial=0)),
                        int(np.max(sec, initial=0)) if self.is_test else 0)
            m["id_max"] = (int(self.uid.max(initial=0)), items, int(self.ucate.max(initial=0)))
            self._m32 = m
        return self._m32

    def collate_packed(self, idx, k, slot=None):
        """Rows ``idx`` as a PackedBatch: the same fields as ``collate``, written straight into a page-locked int32
        buffer in the device's staging layout (no second pass over the batch on the way to the GPU)."""
        m = self._i32()
        idx = np.asarray(idx, np.int64)
        B = len(idx)
        start, stop = self.pre_off[idx], self.pre_off[idx + 1]
        sl = np.minimum(stop - start, k)
        first = stop - sl
        nstart = self.new_off[idx]
        new_sl = self.new_off[idx + 1] - nstart
        S = max(int(new_sl.max()) if B else 0, 1)
        offs, words = staging_layout(B, k, S)
        if slot is None:
            buf, owner = _pinned_words(words)
        else:
            buf, owner = slot[0]
        seg = lambda name, n: buf[offs[name]:offs[name] + n]
        np.take(m["uid"], idx, out=seg("u", B)); np.take(m["cand"], idx, out=seg("i", B))
        np.take(m["second"], idx, out=seg("second", B)); np.take(m["ucate"], idx, out=seg("c", B))
        seg("sl", B)[:] = sl; seg("sl_new", B)[:] = new_sl
        col = np.arange(k, dtype=np.int64)[None, :]
        mask = col < sl[:, None]
        pos = np.where(mask, first[:, None] + col, 0)
        hi = seg("hist_i", B * k).reshape(B, k)
        ht = seg("hist_t", B * k).view(np.float32).reshape(B, k)
        if len(self.pre_items):
            np.take(m["pre_items"], pos, out=hi); hi *= mask
            np.take(self.pre_time, pos, out=ht); ht *= mask
        else:
            hi[:] = 0; ht[:] = 0
        new_off = np.zeros(B + 1, np.int64)
        np.cumsum(new_sl, out=new_off[1:])
        n_new = int(new_off[-1])
        seg("new_off", B)[:] = new_off[:-1]
        if n_new:
            src = np.repeat(nstart - new_off[:-1], new_sl) + np.arange(n_new, dtype=np.int64)
            np.take(m["new_items"], src, out=seg("new_items", n_new))
        pb = PackedBatch(buf, owner, B, k, S, n_new, self.is_test, m["id_max"])
        if slot is not None:
            slot[1] = pb
        return pb

    def _collate(self, idx, k):
        idx = np.asarray(idx, np.int64)
        start, stop = self.pre_off[idx], self.pre_off[idx + 1]
        length = stop - start
        sl = np.minimum(length, k)                                         # input.py:30
        first = stop - sl                                                   # keep the last k (:40-44)
        col = np.arange(k, dtype=np.int64)[None, :]
        mask = col < sl[:, None]
        pos = np.where(mask, first[:, None] + col, 0)
        if len(self.pre_items):
            hist_i = np.where(mask, self.pre_items[pos], 0).astype(np.int64)
            hist_t = np.where(mask, self.pre_time[pos], np.float32(0)).astype(np.float32)
        else:
            hist_i = np.zeros(mask.shape, np.int64); hist_t = np.zeros(mask.shape, np.float32)
        nstart = self.new_off[idx]
        new_sl = self.new_off[idx + 1] - nstart                            # :31
        width = int(new_sl.max()) if len(idx) else 0                       # :32
        ncol = np.arange(width, dtype=np.int64)[None, :]
        nmask = ncol < new_sl[:, None]
        npos = np.where(nmask, nstart[:, None] + ncol, 0)
        if len(self.new_items):
            hist_i_new = np.where(nmask, self.new_items[npos], 0).astype(np.int64)
        else:
            hist_i_new = np.zeros(nmask.shape, np.int64)
        return (self.uid[idx], self.cand[idx], self.second[idx], hist_i, hist_i_new, hist_t, sl, new_sl,
                self.ucate[idx])


def staging_layout(B, L, S):
    """Word offsets of the int32 staging buffer (include/tlsan_b200.h, tlsan_stage_batch_host): the packed batch
    layout u | i | second | c | sl | sl_new | hist_i[B*L] | hist_i_new[B*S] | hist_t[B*L] followed by the ragged form of
    the session matrix, new_off[B] | new_items[<= B*S]; every segment rounded up to 4 words."""
    offs, o = {}, 0
    for name, n in (("u", B), ("i", B), ("second", B), ("c", B), ("sl", B), ("sl_new", B), ("hist_i", B * L),
                    ("hist_i_new", B * S), ("hist_t", B * L), ("new_off", B), ("new_items", B * S)):
        offs[name] = o
        o += (n + 3) // 4 * 4
    return offs, o


def _pinned_words(words):
    """int32 host buffer, page-locked when a CUDA device is present (the H2D copy then runs at DMA speed and
    asynchronously); returns (numpy view, owner)."""
    try:
        import torch
        if torch.cuda.is_available():
            t = torch.empty(int(words), dtype=torch.int32, pin_memory=True)
            return t.numpy(), t
    except ImportError:
        pass
    a = np.empty(int(words), np.int32)
    return a, a


class PackedBatch(object):
    """A batch in the device's staging layout, in ONE page-locked int32 buffer: what ``DataInput(..., packed=True)``
    yields.  ``Model.train`` / ``eval_auc`` copy it to the GPU as is (tlsan_stage_packed: two DMA copies, no host
    pass), instead of casting and packing the nine arrays of the reference tuple every step (model.py:210-222).

    It still reads like the reference 9-tuple: ``len(b) == 9`` and ``b[k]`` is field k of TLSAN/input.py:54,107 with
    the same values (integer fields int32 instead of int64; ``b[4]``, the zero-padded session matrix, is rebuilt from
    its ragged form on demand).  ``id_max = (max user id, max item id, max category id)`` bounds every id the
    producer can emit (its whole dataset); the model compares it with its table sizes in place of per-batch checks."""

    def __init__(self, buf, owner, B, L, S, n_new, is_test, id_max):
        self.buf, self._owner = buf, owner
        self.B, self.L, self.S, self.n_new, self.is_test = int(B), int(L), int(S), int(n_new), bool(is_test)
        self.id_max = tuple(int(x) for x in id_max)
        self.offs, self.words = staging_layout(self.B, self.L, self.S)
        self.copied = None                      # CUDA event of the last H2D copy out of this buffer (set by Model)

    def _seg(self, name, n, dtype=np.int32):
        return self.buf[self.offs[name]:self.offs[name] + n].view(dtype)

    def __len__(self):
        return 9

    def __iter__(self):
        return (self[k] for k in range(9))

    def __getitem__(self, k):
        B, L, S = self.B, self.L, self.S
        if k < 0:
            k += 9
        if k == 0: return self._seg("u", B)
        if k == 1: return self._seg("i", B)
        if k == 2: return self._seg("second", B) if self.is_test else self._seg("second", B, np.float32)
        if k == 3: return self._seg("hist_i", B * L).reshape(B, L)
        if k == 4:
            off, n = self._seg("new_off", B).astype(np.int64), self._seg("sl_new", B).astype(np.int64)
            out = np.zeros((B, S), np.int32)
            col = np.arange(S)[None, :]
            mask = col < n[:, None]
            out[mask] = self._seg("new_items", self.n_new)
            return out
        if k == 5: return self._seg("hist_t", B * L, np.float32).reshape(B, L)
        if k == 6: return self._seg("sl", B)
        if k == 7: return self._seg("sl_new", B)
        if k == 8: return self._seg("c", B)
        raise IndexError(k)

    @classmethod
    def from_tuple(cls, batch, is_test=False):
        """Pack a reference 9-tuple once (multi-threaded C pass, tlsan_stage_batch_host with no device)."""
        import ctypes as C
        from . import _lib
        lib = _lib.lib()
        B, L = len(batch[0]), int(np.shape(batch[3])[1])
        S = max(int(np.shape(batch[4])[1]), 1)
        i64 = lambda x: np.ascontiguousarray(x, dtype=np.int64)
        u, i, c, sl, sn, hi = i64(batch[0]), i64(batch[1]), i64(batch[8]), i64(batch[6]), i64(batch[7]), i64(batch[3])
        hn = i64(batch[4]) if np.shape(batch[4])[1] else np.zeros((B, 1), np.int64)
        ht = np.ascontiguousarray(batch[5], dtype=np.float32)
        i2 = i64(batch[2]) if is_test else None
        y = None if is_test else np.ascontiguousarray(batch[2], dtype=np.float32)
        offs, words = staging_layout(B, L, S)
        buf, owner = _pinned_words(words)
        big = 2 ** 31 - 2                       # ids are bounded by id_max below, not by a model's tables
        dims = _lib.Dims(B=B, L=L, S=S, NI=big, NU=big, NC=big, B_global=B, reserved=0)
        p = lambda a: None if a is None else a.ctypes.data
        rc = lib.tlsan_stage_batch_host(C.byref(dims), p(u), p(i), p(i2), p(y), p(hi), p(hn), p(ht), p(sl), p(sn), p(c),
                                        buf.ctypes.data, None, words, 1, 0, None)
        if rc == -1:
            raise IndexError(lib.tlsan_last_error().decode())
        _lib.check(rc)
        n_new = int(np.clip(sn, 0, S).sum())
        items = max(int(hi.max(initial=0)), int(hn.max(initial=0)), int(i.max(initial=0)),
                    int(i2.max(initial=0)) if is_test else 0)
        return cls(buf, owner, B, L, S, n_new, is_test, (int(u.max(initial=0)), items, int(c.max(initial=0))))


class _Input:
    is_test = False

    def __init__(self, data, batch_size, k, index_dtype=np.int64, packed=False):
        self.k = k
        self.index_dtype = index_dtype
        self.packed = bool(packed)
        self._ring, self._ring_i = [], 0
        self.batch_size = batch_size
        self.data = data
        self.csr = data if isinstance(data, CsrDataset) else CsrDataset.from_samples(data, self.is_test)
        n = len(self.csr)
        self.epoch_size = n // batch_size + (1 if n % batch_size else 0)   # input.py:9-11
        self.i = 0

    def __iter__(self):
        return self

    def __next__(self):
        if self.i == self.epoch_size:
            raise StopIteration
        lo = self.i * self.batch_size
        hi = min(lo + self.batch_size, len(self.csr))
        self.i += 1
        if self.packed:
            return self.i, self.csr.collate_packed(np.arange(lo, hi), self.k, self._staging(hi - lo))
        return self.i, self.csr.collate(np.arange(lo, hi), self.k, self.index_dtype)

    def _staging(self, B, depth=4):
        """Ring of page-locked buffers: a PackedBatch stays valid until `depth` more batches have been drawn (its
        host->device copy, recorded in ``copied``, is waited for before the buffer is written again)."""
        _, words = staging_layout(B, self.k, self.csr.max_new_len())
        if len(self._ring) < depth:
            self._ring.append([_pinned_words(words), None])
        slot = self._ring[self._ring_i % len(self._ring)]
        self._ring_i += 1
        if len(slot[0][0]) < words:
            slot[0] = _pinned_words(words)
        if slot[1] is not None and slot[1].copied is not None:
            slot[1].copied.synchronize()
        return slot


class DataInput(_Input):
    """Training batches, reference TLSAN/input.py:4-54."""
    is_test = False


class DataInputTest(_Input):
    """Evaluation batches (pos / neg item per row), reference TLSAN/input.py:57-107."""
    is_test = True
