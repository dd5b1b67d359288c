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


class _Input:
    is_test = False

    def __init__(self, data, batch_size, k, index_dtype=np.int64):
        self.k = k
        self.index_dtype = index_dtype
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
        return self.i, self.csr.collate(np.arange(lo, hi), self.k, self.index_dtype)


class DataInput(_Input):
    """Training batches, reference TLSAN/input.py:4-54."""
    is_test = False


class DataInputTest(_Input):
    """Evaluation batches (pos / neg item per row), reference TLSAN/input.py:57-107."""
    is_test = True
