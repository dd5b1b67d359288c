"""Device-resident dataset + GPU batch assembly.

The samples built by the reference ``TLSAN/build_dataset.py`` are uploaded ONCE as a CSR image
(`tlsan_dataset_t`); every batch is then assembled in HBM by ``tlsan_collate`` in the exact
layout of ``DataInput.__next__`` / ``DataInputTest.__next__`` (TLSAN/input.py:17-54,70-107),
so an epoch loop never touches the host batcher (10^3-10^4 x slower than the kernels)::

    ds = DeviceDataset(train_set, is_test=False)
    for lo in range(0, len(ds), bs):
        model.train_staged(ds.batch(perm[lo:lo + bs], model.L), lr)
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .input import CsrDataset
from .model import DeviceBatch, _pack_offsets


class DeviceDataset:
    def __init__(self, data, is_test=False, device=None):
        csr = data if isinstance(data, CsrDataset) else CsrDataset.from_samples(data, is_test)
        if csr.is_test != bool(is_test):
            raise ValueError("dataset was built with is_test=%s" % csr.is_test)
        self._lib = _lib.lib()
        if not torch.cuda.is_available():
            raise _lib.TlsanError("DeviceDataset needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.is_test = bool(is_test)
        self.n = len(csr)
        self.new_len = (csr.new_off[1:] - csr.new_off[:-1]).astype(np.int64)       # host copy: batch width
        self.max_new_len = int(self.new_len.max()) if self.n else 1
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.device)
        self._t = dict(
            uid=up(csr.uid, np.int32), pre_off=up(csr.pre_off, np.int64), pre_items=up(csr.pre_items, np.int32),
            pre_time=up(csr.pre_time, np.float32), new_off=up(csr.new_off, np.int64),
            new_items=up(csr.new_items, np.int32), cand=up(csr.cand, np.int32), ucate=up(csr.ucate, np.int32))
        if self.is_test:
            self._t["second_i"] = up(csr.second, np.int32)
        else:
            self._t["second_f"] = up(csr.second, np.float32)
        p = lambda k: self._t[k].data_ptr() if k in self._t else None
        self.c = _lib.Dataset(uid=p("uid"), pre_off=p("pre_off"), pre_items=p("pre_items"), pre_time=p("pre_time"),
                              new_off=p("new_off"), new_items=p("new_items"), cand=p("cand"),
                              second_i=p("second_i"), second_f=p("second_f"), ucate=p("ucate"), n=self.n)

    @classmethod
    def from_device(cls, tensors, is_test, device=None):
        """Wrap CSR arrays that are ALREADY in HBM (the GPU dataset builder's output): keys uid, pre_off, pre_items,
        pre_time, new_off, new_items, cand, ucate and second_i (test) / second_f (train)."""
        self = cls.__new__(cls)
        self._lib = _lib.lib()
        self.device = torch.device(device if device is not None else tensors["uid"].device)
        self.is_test = bool(is_test)
        self.n = int(tensors["uid"].numel())
        self.new_len = (tensors["new_off"][1:] - tensors["new_off"][:-1]).cpu().numpy().astype(np.int64)
        self.max_new_len = int(self.new_len.max()) if self.n else 1
        self._t = dict(tensors)
        p = lambda k: self._t[k].data_ptr() if k in self._t else None
        self.c = _lib.Dataset(uid=p("uid"), pre_off=p("pre_off"), pre_items=p("pre_items"), pre_time=p("pre_time"),
                              new_off=p("new_off"), new_items=p("new_items"), cand=p("cand"),
                              second_i=p("second_i"), second_f=p("second_f"), ucate=p("ucate"), n=self.n)
        return self

    def to_csr(self):
        """Host copy as a CsrDataset (tests, checkpoints)."""
        from .input import CsrDataset
        g = lambda k: self._t[k].cpu().numpy()
        return CsrDataset(g("uid"), g("pre_off"), g("pre_items"), g("pre_time"), g("new_off"), g("new_items"), g("cand"),
                          g("second_i") if self.is_test else g("second_f"), g("ucate"), self.is_test)

    def __len__(self):
        return self.n

    def batch(self, idx, L, width="batch"):
        """Rows ``idx`` (numpy ints, or a device int32 tensor with width='max') as a DeviceBatch.
        width='batch' pads the session to the longest one among the rows like input.py:32-37;
        width='max' pads to the dataset maximum (no host look-up; the extra columns are never read)."""
        if isinstance(idx, torch.Tensor):
            if width != "max":
                raise ValueError("a device index tensor needs width='max'")
            didx = idx.to(device=self.device, dtype=torch.int32).contiguous()
            S = self.max_new_len
        else:
            idx = np.asarray(idx, np.int64)
            if idx.size and (idx.min() < 0 or idx.max() >= self.n):
                raise IndexError("row index out of range")
            S = self.max_new_len if width == "max" else int(self.new_len[idx].max())
            didx = torch.from_numpy(idx.astype(np.int32)).to(self.device)
        S = max(S, 1)
        B = int(didx.numel())
        offs, total = _pack_offsets(B, L, S)
        buf = torch.empty(total, dtype=torch.int32, device=self.device)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._lib.tlsan_collate(C.byref(self.c), didx.data_ptr(), B, L, S, 1 if self.is_test else 0,
                                           buf.data_ptr(), total, st))
        db = DeviceBatch(buf, B, L, S, offs, self.is_test)
        db._keep = didx
        return db
