"""`Model`: the reference TLSAN/model.py surface on top of the sm_100a C-ABI library.

Same constructor and methods as the reference class (TLSAN/model.py:13-313):

    Model(config, item_cate_list)
    train(sess, batch, lr, add_summary=False) -> loss      model.py:208-234
    eval_auc(sess, batch) -> float                         model.py:237-263
    eval_prec(sess, batch) / eval_recall(sess, batch)      model.py:265-299
    save(sess) / restore(sess, path)                       model.py:302-313
    global_step / global_epoch_step / global_epoch_step_op (``.eval()``)

``sess`` is accepted and ignored (there is no TF session).  ``batch`` is the 9-tuple of
TLSAN/input.py:54,107 (lists / numpy, any integer dtype; element 5 may also be the RAW integer day gaps
d[B,L] -- then proc_time_emb of build_dataset.py:16-21 runs inside the gather kernels): it is packed into ONE pinned host
buffer, copied host->device once and handed to the CUDA kernels through the C ABI of
include/tlsan_b200.h.  There is no CPU path: without the CUDA library every compute method
raises.  PyTorch is used for device memory, streams and torch.distributed only.
"""
import ctypes as C
import json
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import OFF, STAT, Batch, Dims, Next, Params, check
from .input import PackedBatch

_L = "all/long_term/num_blocks0_0/long_term_layer/feature_wise_attention1/"
_S = "all/short_term/num_blocks1_0/short_term_layer/feature_wise_attention2/"
_D = "all/long_term/num_blocks0_0/dense/"
# TF variable name -> (offset, shape) inside the packed `dense` vector (include/tlsan_b200.h)
DENSE_LAYOUT = OrderedDict([
    ("gamma_parameter", (OFF["GAMMA"], ())),
    (_L + "bn_dense_map1/linear_map/W", (OFF["W1L"], (8, 8))),
    (_L + "bn_dense_map1/linear_map/bias", (OFF["B1L"], (8,))),
    (_L + "bn_dense_map2/linear_map/W", (OFF["W2L"], (8, 8))),
    (_L + "bn_dense_map2/linear_map/bias", (OFF["B2L"], (8,))),
    (_D + "kernel", (OFF["WD"], (64, 64))),
    (_D + "bias", (OFF["BD"], (64,))),
    (_S + "bn_dense_map1/linear_map/W", (OFF["W1S"], (8, 8))),
    (_S + "bn_dense_map1/linear_map/bias", (OFF["B1S"], (8,))),
    (_S + "bn_dense_map2/linear_map/W", (OFF["W2S"], (8, 8))),
    (_S + "bn_dense_map2/linear_map/bias", (OFF["B2S"], (8,))),
])
TABLES = ("item_emb", "item_b", "user_emb", "usert_emb", "cate_emb")
KS = (1, 10, 20, 30, 40, 50)                                  # model.py:144-156


class _Var:
    """Stands in for a non-trainable tf.Variable read with ``.eval()`` (train.py:94,194,...)."""

    def __init__(self, value=0):
        self.value = value

    def eval(self, session=None):
        return self.value


class _Metric:
    def __init__(self, fn):
        self._fn = fn

    def eval(self, session=None):
        return self._fn()


class _IncrOp:
    def __init__(self, var):
        self._var = var

    def eval(self, session=None):
        self._var.value += 1
        return self._var.value


class LazyLoss(object):
    """Loss of a step whose read-back is still in flight; ``float(x)`` waits for it.  (At most 8 may be pending.)"""

    def __init__(self, host, ev):
        self._host, self._ev, self._v = host, ev, None

    def __float__(self):
        if self._v is None:
            self._ev.synchronize()
            self._v = float(self._host[0])
        return self._v

    def __radd__(self, other):
        return other + float(self)

    def __add__(self, other):
        return float(self) + other


class DeviceBatch:
    """A batch resident in HBM: one packed int32 buffer + the C struct pointing into it."""

    def __init__(self, buf, B, L, S, offs, is_test, raw_gaps=False):
        self.buf, self.B, self.L, self.S, self.is_test, self.raw_gaps = buf, B, L, S, is_test, raw_gaps
        base = buf.data_ptr()
        p = lambda k: base + 4 * offs[k]
        self.offs = offs
        self.c = Batch(u=p("u"), i=p("i"), i2=p("second") if is_test else None,
                       y=None if is_test else p("second"), hist_i=p("hist_i"), hist_i_new=p("hist_i_new"),
                       hist_t=p("hist_t"), sl=p("sl"), sl_new=p("sl_new"), c=p("c"),
                       hist_d=p("hist_t") if raw_gaps else None)     # raw int32 day gaps travel in the hist_t words
        self.nbytes = buf.numel() * 4


def _pack_offsets(B, L, S):
    offs, o = {}, 0
    for name, n in (("u", B), ("i", B), ("second", B), ("c", B), ("sl", B), ("sl_new", B),
                    ("hist_i", B * L), ("hist_i_new", B * S), ("hist_t", B * L)):
        offs[name] = o
        o += (n + 3) // 4 * 4
    return offs, o


def pack_batch(lib, batch, dims, is_test, out, validate=True, dev_ptr=None, stream=None):
    """Host-side feed (reference model.py:210-222): the 9-tuple -> packed int32 words in ``out``.
    Out-of-range ids raise IndexError, like tf.gather on CPU (InvalidArgumentError).
    With ``dev_ptr`` the words are also copied to the device on ``stream``, phase by phase while the
    rest is still being packed (tlsan_stage_batch_host; ``out`` must then be pinned memory)."""
    # integer fields that are ALL int32 already (tlsan_b200.input emits them) skip the 64 -> 32 bit narrowing pass
    ints = [batch[k] for k in (0, 1, 3, 4, 6, 7, 8)] + ([batch[2]] if is_test else [])
    narrow = all(isinstance(x, np.ndarray) and x.dtype == np.int32 for x in ints)
    idt = np.int32 if narrow else np.int64

    def i64(x):
        return np.ascontiguousarray(x, dtype=idt)
    B, S = dims.B, dims.S
    u, i, c, sl, sl_new = i64(batch[0]), i64(batch[1]), i64(batch[8]), i64(batch[6]), i64(batch[7])
    hist_i = i64(batch[3])
    hist_i_new = i64(batch[4])
    if hist_i_new.shape[1] == 0:
        hist_i_new = np.zeros((B, 1), idt)
    if np.issubdtype(np.asarray(batch[5]).dtype, np.integer):
        # raw day gaps d = cur_day - day + 1 (0 = padding): same 4-byte words, bucketed on the GPU while gathering
        hist_t = np.ascontiguousarray(batch[5], dtype=np.int32).view(np.float32)
    else:
        hist_t = np.ascontiguousarray(batch[5], dtype=np.float32)
    i2 = i64(batch[2]) if is_test else None
    y = None if is_test else np.ascontiguousarray(batch[2], dtype=np.float32)
    if hist_i.shape != (B, dims.L) or hist_t.shape != (B, dims.L) or hist_i_new.shape != (B, S):
        raise ValueError("batch arrays have inconsistent shapes")
    p = lambda a: None if a is None else a.ctypes.data
    f_pack = lib.tlsan_pack_batch_host_i32 if narrow else lib.tlsan_pack_batch_host
    f_stage = lib.tlsan_stage_batch_host_i32 if narrow else lib.tlsan_stage_batch_host
    if dev_ptr is None:
        rc = f_pack(C.byref(dims), p(u), p(i), p(i2), p(y), p(hist_i), p(hist_i_new), p(hist_t),
                                       p(sl), p(sl_new), p(c), out.ctypes.data, out.size, 1 if validate else 0, 0)
    else:
        rc = f_stage(C.byref(dims), p(u), p(i), p(i2), p(y), p(hist_i), p(hist_i_new), p(hist_t),
                                        p(sl), p(sl_new), p(c), out.ctypes.data, dev_ptr, out.size,
                                        1 if validate else 0, 0, stream)
    if rc == -1:
        raise IndexError(lib.tlsan_last_error().decode())
    check(rc)


class Model(object):
    @staticmethod
    def _check_config(config):
        """The kernels are built for the reference defaults that define the path (train.py:26-49)."""
        if config.get("num_blocks", 1) != 1:
            raise ValueError("num_blocks != 1 is ill-formed in the reference (model.py:331-364); unsupported")
        if (config.get("hidden_units", 64), config.get("num_heads", 8)) != (64, 8) or any(
                config.get(k, 32) != 32 for k in ("itemid_embedding_size", "userid_embedding_size",
                                                  "cateid_embedding_size")):
            raise ValueError("kernels are built for hidden_units=64, num_heads=8, embedding sizes 32")
        if config.get("dropout", 0.0) != 0.0:
            raise ValueError("dropout > 0 is not supported (reference default 0.0, train.py:30)")

    def __init__(self, config, item_cate_list, device=None, seed=1234, process_group=None, validate=True,
                 dp_mode=None):
        self.config = config
        self._check_config(config)
        self._lib = _lib.lib()                                  # raises if the CUDA library is missing
        if not torch.cuda.is_available():
            raise _lib.TlsanError("tlsan_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.NI, self.NU, self.NC = int(config["item_count"]), int(config["user_count"]), int(config["cate_count"])
        self.L = int(config["Ls"])
        if not 1 <= self.L <= _lib.MAX_L:
            raise ValueError("Ls must be in [1, %d]" % _lib.MAX_L)
        self.reg = float(config.get("regulation_rate", 0.00005))
        self.clip = float(config.get("max_gradient_norm", 5.0))
        self.validate = validate
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.rank = torch.distributed.get_rank(process_group) if process_group is not None else 0
        # data-parallel exchange: "nccl" = all_reduce (NVLS in-switch reduction on NVSwitch) + tlsan_apply_flat
        # [default: measured faster at 2 and 8 GPUs]; "p2p" = fused reduce-scatter / update / all-gather over
        # NVLink peer memory (tlsan_dp_exchange)
        self.dp_mode = dp_mode or os.environ.get("TLSAN_DP_MODE", "nccl")
        if self.dp_mode not in ("p2p", "nccl"):
            raise ValueError("dp_mode must be 'p2p' or 'nccl'")
        self._arenas = None
        self._dp_epoch = 0
        self._dp_steps_unchecked = 0
        # the peer-memory exchange spins (bounded) on its peers' flags; a timed-out step skips its update and raises
        # TLSAN_STAT_DP_ERR, which is read back (one 4-byte D2H) every `dp_check_every` steps
        self.dp_check_every = int(os.environ.get("TLSAN_DP_CHECK_EVERY", "64"))

        icl = np.ascontiguousarray(np.asarray(item_cate_list, dtype=np.int32))
        if icl.shape != (self.NI,) or icl.min() < 0 or icl.max() >= self.NC:
            raise ValueError("item_cate_list must be int[item_count] with values in [0, cate_count)")
        order = np.argsort(icl, kind="stable").astype(np.int32)
        cate_off = np.zeros(self.NC + 1, np.int32)
        np.cumsum(np.bincount(icl, minlength=self.NC), out=cate_off[1:])
        dev = self.device
        self.icl = torch.from_numpy(icl).to(dev)
        self.cate_off = torch.from_numpy(cate_off).to(dev)
        self.cate_items = torch.from_numpy(order).to(dev)

        # ---- variables (model.py:56-81, 347, 443-454); glorot_uniform = tf.get_variable default
        g = torch.Generator().manual_seed(seed)

        def glorot(rows, cols):
            lim = (6.0 / (rows + cols)) ** 0.5
            return (torch.rand(rows, cols, generator=g) * 2 - 1) * lim

        NR = self.NI + self.NC + self.NU
        emb = torch.empty(NR, 32)
        emb[:self.NI] = glorot(self.NI, 32)
        emb[self.NI + self.NC:] = glorot(self.NU, 32)
        emb[self.NI:self.NI + self.NC] = glorot(self.NC, 32)
        # all trainable state in ONE device buffer, emb | usert | item_b | dense (each padded to 16 B): the layout
        # the NVLink data-parallel exchange sums and updates slice by slice (tlsan_dp_exchange)
        up4 = lambda n: (n + 3) // 4 * 4
        o_usert = NR * 32
        o_itemb = o_usert + up4(self.NU * self.L)
        o_dense = o_itemb + up4(self.NI)
        self._wflat = torch.zeros(o_dense + _lib.DENSE_PAD, device=dev)
        self.emb = self._wflat[:o_usert].view(NR, 32)
        self.emb.copy_(emb)
        self.item_emb = self.emb[:self.NI]
        self.cate_emb = self.emb[self.NI:self.NI + self.NC]
        self.user_emb = self.emb[self.NI + self.NC:]
        self.item_b = self._wflat[o_itemb:o_itemb + self.NI]
        self.usert_emb = self._wflat[o_usert:o_usert + self.NU * self.L].view(self.NU, self.L)
        self.usert_emb.fill_(-1.0)
        dense = torch.zeros(_lib.DENSE_PAD)
        for name, (off, shape) in DENSE_LAYOUT.items():
            if name == "gamma_parameter":
                dense[off] = 1.0
            elif len(shape) == 2:
                dense[off:off + shape[0] * shape[1]] = glorot(*shape).reshape(-1)
        self.dense = self._wflat[o_dense:]
        self.dense.copy_(dense)
        self._params = Params(emb=self.emb.data_ptr(), usert=self.usert_emb.data_ptr(),
                              item_b=self.item_b.data_ptr(), dense=self.dense.data_ptr(),
                              icl=self.icl.data_ptr(), cate_off=self.cate_off.data_ptr(),
                              cate_items=self.cate_items.data_ptr())

        # ---- optimizer (model.py:188-195): anything but adadelta / adam / rmsprop is GradientDescentOptimizer
        opt = config.get("optimizer", "sgd")
        self.optimizer = opt if opt in ("adadelta", "adam", "rmsprop") else "sgd"
        self._slots = None
        if self.optimizer != "sgd":
            if self.world > 1 and self.dp_mode == "p2p":
                raise ValueError("the peer-memory exchange (dp_mode='p2p') implements sgd only; use dp_mode='nccl'")
            # slot variables mirror the weight buffer element for element; TF 1.8 initial values: adam m, v = 0;
            # rmsprop rms = 1, momentum = 0; adadelta accum, accum_update = 0
            self._slots = (torch.ones_like(self._wflat) if self.optimizer == "rmsprop" else torch.zeros_like(self._wflat),
                           torch.zeros_like(self._wflat))
            d = dict(adam=(0.9, 0.999, 0.0, 0.0, 1e-8), rmsprop=(0.0, 0.0, 0.9, 0.0, 1e-10),
                     adadelta=(0.0, 0.0, 0.95, 0.0, 1e-8))[self.optimizer]
            self._opt = _lib.Opt(kind=_lib.OPT_KIND[self.optimizer], step=0, beta1=d[0], beta2=d[1], rho=d[2],
                                 momentum=d[3], epsilon=d[4], slot1=self._slots[0].data_ptr(),
                                 slot2=self._slots[1].data_ptr())

        self._stats = torch.zeros(_lib.STAT_COUNT, device=dev)
        self._ws = [None, None]
        self._ws_slot = 0
        self._presorted = None      # (batch buffer address, workspace slot) sorted ahead by the previous step
        self._score_ws = None
        self._rank_ws = None
        self._flat = None
        self._stage_cache = {}
        self._copy_stream = None
        self._prefetched = None
        self.last_h2d_bytes = 0
        self.last_d2h_bytes = 0

        # ---- step variables and streaming metrics (model.py:142-161)
        self.global_step = _Var(0)
        self.global_epoch_step = _Var(0)
        self.global_epoch_step_op = _IncrOp(self.global_epoch_step)
        self.reset_metrics()
        for n, k in enumerate(KS):
            setattr(self, "prec_%d" % k, _Metric(lambda n=n: self._metric(n, "p")))
            setattr(self, "recall_%d" % k, _Metric(lambda n=n: self._metric(n, "r")))
        self.train_writer = None
        self.eval_writer = None

    # ------------------------------------------------------------------ plumbing
    def _dims(self, B, S, B_global=None, flags=0):
        return Dims(B=B, L=self.L, S=S, NI=self.NI, NU=self.NU, NC=self.NC,
                    B_global=int(B_global if B_global is not None else B), reserved=flags)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _workspace(self, dims, slot=0):
        """Step workspace; two slots so that a pipelined step can presort the next batch into the other one."""
        need = C.c_size_t()
        # sized for the session width rounded up to 8 columns: S is the longest session of the batch (input.py:33)
        # and wobbles from batch to batch; a reallocation costs a device synchronisation and a cudaMalloc
        roomy = Dims(B=dims.B, L=dims.L, S=(dims.S + 7) // 8 * 8, NI=dims.NI, NU=dims.NU, NC=dims.NC,
                     B_global=dims.B_global, reserved=0)
        check(self._lib.tlsan_workspace_bytes(C.byref(roomy), C.byref(need)))
        if self._ws[slot] is None or self._ws[slot].numel() < need.value:
            if self._ws[slot] is not None:
                torch.cuda.synchronize(self.device)      # the library's side streams may still be using the old one
            self._ws[slot] = torch.empty(int(need.value), dtype=torch.uint8, device=self.device)
            if self._presorted is not None and self._presorted[2] == slot:
                self._presorted = None
        return self._ws[slot]

    def stage_batch(self, batch, is_test=False, out=None):
        """Pack the input.py 9-tuple into pinned memory (tlsan_pack_batch_host: multi-threaded
        int64->int32 cast + id range checks) and copy it to the device (one H2D).  `out` (optional): a device
        int32 tensor to stage into instead of a fresh allocation (the double-buffered feed of `prefetch`)."""
        if isinstance(batch, PackedBatch):
            return self._stage_packed(batch, is_test, out)
        B = len(batch[0])
        S = max(int(np.shape(batch[4])[1]), 1)
        if np.shape(batch[3])[1] != self.L:
            raise ValueError("hist_i has %d columns but the model was built with Ls=%d" % (np.shape(batch[3])[1], self.L))
        offs, total = _pack_offsets(B, self.L, S)
        dims = self._dims(B, S)
        words = C.c_int64()
        check(self._lib.tlsan_stage_words(C.byref(dims), C.byref(words)))    # packed layout + ragged session tail
        words = int(words.value)
        key = (B, S)
        if key not in self._stage_cache:
            self._stage_cache[key] = [0, [(torch.empty(words, dtype=torch.int32).pin_memory(), torch.cuda.Event())
                                          for _ in range(2)]]
        slot = self._stage_cache[key]
        slot[0] ^= 1                          # two pinned buffers per shape: one may still feed a copy in flight
        host, ev = slot[1][slot[0]]
        ev.synchronize()                      # the previous copy out of this pinned buffer has finished
        dev = out[:words] if out is not None else torch.empty(words, dtype=torch.int32, device=self.device)
        pack_batch(self._lib, batch, dims, is_test, host.numpy(), self.validate, dev.data_ptr(), self._stream())
        ev.record(torch.cuda.current_stream(self.device))
        # bytes that actually crossed PCIe: everything but the padded session matrix, plus its ragged form
        n_new = int(np.clip(np.asarray(batch[7], dtype=np.int64), 0, S).sum())
        self.last_h2d_bytes = 4 * (total - (B * S + 3) // 4 * 4 + B + n_new)
        return DeviceBatch(dev, B, self.L, S, offs, is_test,
                           raw_gaps=np.issubdtype(np.asarray(batch[5]).dtype, np.integer))

    def _stage_packed(self, pb, is_test, out=None):
        """The packed feed (tlsan_b200.input.PackedBatch): the batch already sits in page-locked memory in the staging
        layout, so the feed is two DMA copies and the session expansion (tlsan_stage_packed) -- no cast, no pack.  The
        range checks of the feed (tf.gather's InvalidArgumentError) compare the id maxima the producer vouches for with
        this model's tables."""
        if pb.L != self.L:
            raise ValueError("batch was built with Ls=%d but the model with Ls=%d" % (pb.L, self.L))
        if pb.is_test != bool(is_test):
            raise ValueError("PackedBatch.is_test does not match the call (train batches carry labels, test batches a second item)")
        if self.validate:
            for name, mx, n in (("u", pb.id_max[0], self.NU), ("item ids", pb.id_max[1], self.NI), ("c", pb.id_max[2], self.NC)):
                if mx >= n:
                    raise IndexError("batch field %s out of range [0, %d): the producer's largest id is %d" % (name, n, mx))
        dims = self._dims(pb.B, pb.S)
        words = pb.words
        dev = out[:words] if out is not None else torch.empty(words, dtype=torch.int32, device=self.device)
        check(self._lib.tlsan_stage_packed(C.byref(dims), pb.buf.ctypes.data, dev.data_ptr(), words, pb.n_new, self._stream()))
        pb.copied = torch.cuda.Event()
        pb.copied.record(torch.cuda.current_stream(self.device))
        o = pb.offs
        self.last_h2d_bytes = 4 * (o["hist_i_new"] + o["new_items"] + pb.n_new - o["hist_t"])
        offs, _ = _pack_offsets(pb.B, self.L, pb.S)
        return DeviceBatch(dev, pb.B, self.L, pb.S, offs, is_test)

    # ------------------------------------------------------------------ training
    def train_staged(self, db, lr, global_batch=None, next_db=None, next_ready=None):
        """One optimiser step on a device-resident batch; returns the device stats tensor
        (index with ``tlsan_b200._lib.STAT``) without synchronising.  ``next_db`` (optional) is the batch the
        NEXT call will train on: its occurrence sort is then enqueued behind this step's backward kernels
        (tlsan_*_pipelined), off the next step's critical path.  Results do not depend on it."""
        Bg = self._global_rows(db.B, global_batch)
        slot = self._ws_slot
        flags = 0
        if self._presorted is not None:
            # the announced batch is matched by IDENTITY (the model holds a reference, so its buffer can neither be
            # freed under the presort kernels nor be recycled for a different batch at the same address)
            pdb, pBg, pslot = self._presorted
            if pdb is db and pBg == Bg:
                slot, flags = pslot, 2
            self._presorted = None
        dims = self._dims(db.B, db.S, Bg, flags)
        ws = self._workspace(dims, slot)
        st = self._stream()
        nxt = None
        if next_db is not None:
            nBg = next_db.B * self.world if global_batch is None else global_batch
            if self.world > 1 and global_batch is None and next_db.B != db.B:
                next_db = None          # its global row count is unknown until its own step: no presort
        if next_db is not None:
            ndims = self._dims(next_db.B, next_db.S, nBg)
            nws = self._workspace(ndims, 1 - slot)
            # next_ready: torch.cuda.Event after which next_db's buffer is complete (its H2D copy runs on another stream)
            nxt = Next(dims=C.pointer(ndims), batch=C.pointer(next_db.c), workspace=nws.data_ptr(),
                       workspace_bytes=nws.numel(),
                       ready_event=next_ready.cuda_event if next_ready is not None else None)
        nref = C.byref(nxt) if nxt is not None else None
        if self.optimizer != "sgd":
            # gradients into the flat buffer (summed over the ranks when data parallel), then the fused
            # L2 + clip + adam / rmsprop / adadelta update of every table row and small parameter
            n = C.c_int64()
            check(self._lib.tlsan_flat_count(C.byref(dims), C.byref(n)))
            if self._flat is None or self._flat.numel() != n.value:
                self._flat = torch.empty(int(n.value), dtype=torch.float32, device=self.device)
            check(self._lib.tlsan_step_grads_pipelined(C.byref(dims), C.byref(self._params), C.byref(db.c), nref,
                                                       ws.data_ptr(), ws.numel(), self._flat.data_ptr(), st))
            if self.world > 1:
                torch.distributed.all_reduce(self._flat, group=self.pg)
            self._opt.step += 1
            check(self._lib.tlsan_apply_flat_opt(C.byref(dims), C.byref(self._params), self._flat.data_ptr(), lr,
                                                 self.reg, self.clip, C.byref(self._opt), ws.data_ptr(), ws.numel(),
                                                 self._stats.data_ptr(), st))
        elif self.world == 1:
            check(self._lib.tlsan_train_step_pipelined(C.byref(dims), C.byref(self._params), C.byref(db.c), nref, lr,
                                                       self.reg, self.clip, ws.data_ptr(), ws.numel(),
                                                       self._stats.data_ptr(), st))
        else:
            if self.dp_mode == "p2p" and self.world <= 16:
                arenas = self._dp_arenas()                     # this rank's arena doubles as the flat gradient buffer
                check(self._lib.tlsan_step_grads_pipelined(C.byref(dims), C.byref(self._params), C.byref(db.c), nref,
                                                           ws.data_ptr(), ws.numel(), arenas[self.rank], st))
                self._dp_epoch += 1
                check(self._lib.tlsan_dp_exchange(C.byref(dims), C.byref(self._params), arenas, self.rank, self.world,
                                                  self._dp_epoch, lr, self.reg, self.clip, ws.data_ptr(), ws.numel(),
                                                  self._stats.data_ptr(), st))
            else:
                n = C.c_int64()
                check(self._lib.tlsan_flat_count(C.byref(dims), C.byref(n)))
                if self._flat is None or self._flat.numel() != n.value:
                    self._flat = torch.empty(int(n.value), dtype=torch.float32, device=self.device)
                check(self._lib.tlsan_step_grads_pipelined(C.byref(dims), C.byref(self._params), C.byref(db.c), nref,
                                                           ws.data_ptr(), ws.numel(), self._flat.data_ptr(), st))
                torch.distributed.all_reduce(self._flat, group=self.pg)
                check(self._lib.tlsan_apply_flat(C.byref(dims), C.byref(self._params), self._flat.data_ptr(), lr,
                                                 self.reg, self.clip, ws.data_ptr(), ws.numel(),
                                                 self._stats.data_ptr(), st))
        if self.world > 1 and self.dp_mode == "p2p" and self.world <= 16:
            self._dp_steps_unchecked += 1
            if self._dp_steps_unchecked >= self.dp_check_every:
                self.check_dp_health()
        if next_db is not None:
            self._presorted = (next_db, nBg, 1 - slot)
            self._ws_slot = 1 - slot
        self.global_step.value += 1
        return self._stats

    def _global_rows(self, B, global_batch):
        """Denominator of reduce_mean (model.py:171) = rows of the GLOBAL batch.  One process: B.  Data parallel:
        the caller's `global_batch`, else the sum of the ranks' row counts (one 8-byte all_reduce per step -- ranks
        may hold uneven blocks, parallel.shard_rows, and the last batch of an epoch is ragged)."""
        if global_batch is not None:
            return int(global_batch)
        if self.world == 1:
            return B
        t = torch.tensor([B], dtype=torch.int64, device=self.device)
        torch.distributed.all_reduce(t, group=self.pg)
        return int(t.item())

    def check_dp_health(self):
        """Raise if a peer-memory exchange step timed out waiting for a peer (that step left the weights
        untouched on this rank; the replicas may have diverged)."""
        self._dp_steps_unchecked = 0
        if float(self._stats[STAT["dp_err"]].item()) != 0.0:
            raise _lib.TlsanError("tlsan_dp_exchange: timed out waiting for a peer rank; the update of that step "
                                  "was skipped on this rank -- replicas may have diverged, restore from a checkpoint")

    def __del__(self):
        # kernels on the library's side streams (sort, table norms, a presort of the next batch) may still be reading
        # this model's buffers; the caching allocator only orders frees against the current stream
        try:
            torch.cuda.synchronize(self.device)
        except Exception:
            pass

    def _dp_arenas(self):
        """IPC-shared exchange arenas of all ranks (created once): every rank cudaMallocs one, the 64-byte IPC
        handles travel through torch.distributed, peers map them (NVLink peer memory)."""
        if self._arenas is None:
            dims = self._dims(1, 1)
            need = C.c_size_t()
            check(self._lib.tlsan_dp_arena_bytes(C.byref(dims), self.world, C.byref(need)))
            mine, handle = C.c_void_p(), C.create_string_buffer(64)
            check(self._lib.tlsan_dp_arena_create(need.value, C.byref(mine), handle))
            handles = [None] * self.world
            torch.distributed.all_gather_object(handles, bytes(handle.raw), group=self.pg)
            ptrs = (C.c_void_p * self.world)()
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs[r] = mine.value
                else:
                    peer = C.c_void_p()
                    check(self._lib.tlsan_dp_arena_open(h, C.byref(peer)))
                    ptrs[r] = peer.value
            torch.distributed.barrier(group=self.pg)
            self._arenas = ptrs
        return self._arenas

    def prefetch(self, batch, is_test=False):
        """Stage `batch` NOW on a side copy stream (pack on the host, H2D, session expansion); the next train /
        eval call that is given the same tuple object picks the staged copy up instead of staging again."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._feed_ring, self._feed_i = [], 0
        if isinstance(batch, PackedBatch):
            B, S = batch.B, batch.S
        else:
            B, S = len(batch[0]), max(int(np.shape(batch[4])[1]), 1)
        words = C.c_int64()
        check(self._lib.tlsan_stage_words(C.byref(self._dims(B, S)), C.byref(words)))
        words = int(words.value)
        if len(self._feed_ring) < 3:          # three persistent device buffers: staged / in use / being filled
            self._feed_ring.append([torch.empty(words + words // 8, dtype=torch.int32, device=self.device), None])
        slot = self._feed_ring[self._feed_i % len(self._feed_ring)]
        self._feed_i += 1
        if slot[0].numel() < words:
            torch.cuda.synchronize(self.device)
            slot[0] = torch.empty(words + words // 8, dtype=torch.int32, device=self.device)
            slot[1] = None
        if slot[1] is not None:
            self._copy_stream.wait_event(slot[1])      # the step that last read this buffer has finished
        with torch.cuda.stream(self._copy_stream):
            db = self.stage_batch(batch, is_test=is_test, out=slot[0])
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._prefetched = (batch, is_test, db, ev, slot)

    def drop_prefetch(self):
        self._prefetched = None

    def _staged(self, batch, is_test):
        """(DeviceBatch, ring slot or None) for `batch`: the prefetched copy if it is this very tuple."""
        pf, self._prefetched = self._prefetched, None
        if pf is not None and pf[0] is batch and pf[1] == is_test:
            torch.cuda.current_stream(self.device).wait_event(pf[3])
            return pf[2], pf[4]
        return self.stage_batch(batch, is_test=is_test), None

    def train(self, sess, batch, lr, add_summary=False, global_batch=None, prefetch=None, lazy_loss=False):
        """Reference Model.train (model.py:208-234): feed the 9-tuple, run [loss, train_op].
        Data parallel: `batch` is this rank's block of the global batch; `global_batch` = its total row count
        (found with one small all_reduce when omitted).  `prefetch` (optional) = the batch of the NEXT call: it is
        staged on the copy stream while this step's kernels run (double-buffered feed) and its occurrence sort is
        enqueued behind this step's backward kernels.  `lazy_loss=True` returns a ``LazyLoss`` (the 4-byte read-back is
        enqueued, ``float()`` waits for it): a loop that reads loss k after calling train for batch k+1 keeps the GPU
        busy across the read-back."""
        db, slot = self._staged(batch, False)
        nxt_db = nxt_ev = None
        # a PackedBatch costs the host two memcpy calls: stage it FIRST, so that this step can name it and its
        # occurrence sort runs behind this step's backward kernels.  A 9-tuple costs a ~1 ms host pass: enqueue this
        # step first and pack while the GPU works.
        early = prefetch is not None and isinstance(prefetch, PackedBatch)
        if early:
            self.prefetch(prefetch)
            if not (self.world > 1 and global_batch is None):   # (the next step's global row count must be known)
                nxt_db, nxt_ev = self._prefetched[2], self._prefetched[3]
        stats = self.train_staged(db, float(lr), global_batch=global_batch, next_db=nxt_db, next_ready=nxt_ev)
        if slot is not None:                  # its staging buffer may be refilled once this step has run
            slot[1] = torch.cuda.Event()
            slot[1].record(torch.cuda.current_stream(self.device))
        if prefetch is not None and not early:
            self.prefetch(prefetch)
        self.last_d2h_bytes = 4
        if lazy_loss and not (add_summary and self.train_writer is not None):
            return self._lazy(stats)
        loss = float(stats[STAT["loss"]].item())
        if add_summary and self.train_writer is not None:
            self._write_train_summary(db, loss, stats)
        return loss

    def _lazy(self, stats):
        if not hasattr(self, "_loss_ring"):
            self._loss_ring = [torch.empty(1, dtype=torch.float32, pin_memory=True) for _ in range(8)]
            self._loss_i = 0
        host = self._loss_ring[self._loss_i % 8]
        self._loss_i += 1
        host.copy_(stats[STAT["loss"]:STAT["loss"] + 1], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return LazyLoss(host, ev)

    def _write_train_summary(self, db, loss, stats):
        """self.train_summary of model.py:174-183,228-230: five variable histograms, the histogram of the batch's
        attention output u_t, 'L2_norm_user_item' (the l2_norm sum, :164-169) and 'Training Loss'."""
        w, step = self.train_writer, self.global_step.eval()
        w.add_scalar("Training Loss", loss, step)
        bce = float(stats[STAT["bce"]].item())
        if self.reg > 0:
            w.add_scalar("L2_norm_user_item", (loss - bce) / self.reg, step)
        if hasattr(w, "add_histogram"):
            w.add_histogram("gamma", self.dense[OFF["GAMMA"]:OFF["GAMMA"] + 1], step)
            w.add_histogram("embedding/1_item_emb", self.item_emb, step)
            w.add_histogram("embedding/2_user_emb", self.user_emb, step)
            w.add_histogram("embedding/3_cate_emb", self.cate_emb, step)
            w.add_histogram("embedding/4_usert_emb", self.usert_emb, step)
            _, ut = self.score_staged(db, 1, want_ut=True)
            w.add_histogram("attention_output", ut, step)

    # ------------------------------------------------------------------ scoring
    def score_staged(self, db, ncand=1, want_ut=False):
        dims = self._dims(db.B, db.S)
        logits = torch.empty(db.B, ncand, dtype=torch.float32, device=self.device)
        ut = torch.empty(db.B, 64, dtype=torch.float32, device=self.device) if want_ut else None
        if db.B >= 2048:                                       # batched-dense path needs a scratch workspace
            need = C.c_size_t()
            check(self._lib.tlsan_score_workspace_bytes(C.byref(dims), C.byref(need)))
            if self._score_ws is None or self._score_ws.numel() < need.value:
                self._score_ws = torch.empty(int(need.value), dtype=torch.uint8, device=self.device)
            check(self._lib.tlsan_score_ws(C.byref(dims), C.byref(self._params), C.byref(db.c), ncand,
                                           logits.data_ptr(), ut.data_ptr() if want_ut else None,
                                           self._score_ws.data_ptr(), self._score_ws.numel(), self._stream()))
        else:
            check(self._lib.tlsan_score(C.byref(dims), C.byref(self._params), C.byref(db.c), ncand,
                                        logits.data_ptr(), ut.data_ptr() if want_ut else None, self._stream()))
        return logits, ut

    def logits(self, batch, cand_index=1):
        """sess.run(self.logits, ...) with ``self.i = batch[cand_index]`` (model.py:239-261)."""
        b = list(batch)
        b[1] = batch[cand_index]
        b[2] = batch[cand_index]
        db = self.stage_batch(tuple(b), is_test=True)
        lg, _ = self.score_staged(db, 1)
        return lg[:, 0].cpu().numpy()

    def eval_auc(self, sess, batch):
        """Reference Model.eval_auc (model.py:237-263): mean(logit(pos) - logit(neg) > 0)."""
        db = self.stage_batch(batch, is_test=True)
        lg, _ = self.score_staged(db, 2)
        res = lg.cpu().numpy()
        self.last_d2h_bytes = res.nbytes
        return np.mean(res[:, 0] - res[:, 1] > 0)

    def _update_topk(self, batch):
        db = self.stage_batch(batch, is_test=True)
        _, ut = self.score_staged(db, 1, want_ut=True)
        dims = self._dims(db.B, db.S)
        rank = torch.empty(db.B, dtype=torch.int32, device=self.device)
        label = db.buf[db.offs["i"]:db.offs["i"] + db.B]
        if os.environ.get("TLSAN_RANK_IMPL", "tc") == "ffma":     # CUDA-core formulation, kept for A/B timing
            check(self._lib.tlsan_label_rank(C.byref(dims), C.byref(self._params), ut.data_ptr(), label.data_ptr(),
                                             rank.data_ptr(), self._stream()))
        else:
            need = C.c_size_t()
            check(self._lib.tlsan_rank_workspace_bytes(C.byref(dims), C.byref(need)))
            if self._rank_ws is None or self._rank_ws.numel() < need.value:
                self._rank_ws = torch.empty(int(need.value), dtype=torch.uint8, device=self.device)
            check(self._lib.tlsan_label_rank_ws(C.byref(dims), C.byref(self._params), ut.data_ptr(), label.data_ptr(),
                                                rank.data_ptr(), self._rank_ws.data_ptr(), self._rank_ws.numel(),
                                                self._stream()))
        return rank.cpu().numpy()

    def _metric(self, n, which):
        tp, other = (self._ptp[n], self._pfp[n]) if which == "p" else (self._rtp[n], self._rfn[n])
        return float(tp / (tp + other)) if tp + other > 0 else float("nan")

    def eval_prec(self, sess, batch):
        """Reference Model.eval_prec (model.py:265-281): run the six precision_at_k update ops and
        return their values.  The accumulators are never reset by the reference driver
        (train.py:75-76,82), so values are cumulative; ``reset_metrics`` re-initialises them."""
        rank = self._update_topk(batch)
        for n, k in enumerate(KS):
            hit = float(np.sum(rank < k))
            self._ptp[n] += hit
            self._pfp[n] += len(rank) * k - hit
        return [self._metric(n, "p") for n in range(len(KS))]

    def eval_recall(self, sess, batch):
        """Reference Model.eval_recall (model.py:283-299)."""
        rank = self._update_topk(batch)
        for n, k in enumerate(KS):
            hit = float(np.sum(rank < k))
            self._rtp[n] += hit
            self._rfn[n] += len(rank) - hit
        return [self._metric(n, "r") for n in range(len(KS))]

    def reset_metrics(self):
        """sess.run(tf.initialize_variables(metric_ops)) of train.py:75-76."""
        self._ptp = np.zeros(len(KS)); self._pfp = np.zeros(len(KS))
        self._rtp = np.zeros(len(KS)); self._rfn = np.zeros(len(KS))

    # ------------------------------------------------------------------ state
    def slot_views(self):
        """Optimizer slot variables per TF variable name: {name: (slot1, slot2)} (adam m / v, rmsprop rms / momentum,
        adadelta accum / accum_update); None for sgd."""
        if self._slots is None:
            return None
        NR, up4 = self.NI + self.NC + self.NU, lambda n: (n + 3) // 4 * 4
        o_usert = NR * 32; o_itemb = o_usert + up4(self.NU * self.L); o_dense = o_itemb + up4(self.NI)
        out = OrderedDict()
        for s_ in (0, 1):
            f = self._slots[s_]
            e = f[:o_usert].view(NR, 32)
            t = dict(item_emb=e[:self.NI], cate_emb=e[self.NI:self.NI + self.NC], user_emb=e[self.NI + self.NC:],
                     usert_emb=f[o_usert:o_usert + self.NU * self.L].view(self.NU, self.L),
                     item_b=f[o_itemb:o_itemb + self.NI])
            for name, (off, shape) in DENSE_LAYOUT.items():
                n = int(np.prod(shape)) if shape else 1
                t[name] = f[o_dense + off:o_dense + off + n].reshape(shape)
            for k, v in t.items():
                out.setdefault(k, [None, None])[s_] = v
        return out

    def state_dict(self):
        """All trainable variables keyed by their TF names (model.py:56-81, scopes :328-364)."""
        sd = OrderedDict()
        sd["item_emb"] = self.item_emb.detach().cpu().clone()
        sd["item_b"] = self.item_b.detach().cpu().clone()
        sd["user_emb"] = self.user_emb.detach().cpu().clone()
        sd["usert_emb"] = self.usert_emb.detach().cpu().clone()
        sd["cate_emb"] = self.cate_emb.detach().cpu().clone()
        dense = self.dense.detach().cpu()
        for name, (off, shape) in DENSE_LAYOUT.items():
            n = int(np.prod(shape)) if shape else 1
            sd[name] = dense[off:off + n].reshape(shape).clone()
        return sd

    def load_state_dict(self, sd):
        for name in TABLES:
            getattr(self, name).copy_(torch.as_tensor(np.asarray(sd[name]), dtype=torch.float32))
        dense = self.dense.detach().cpu().clone()
        for name, (off, shape) in DENSE_LAYOUT.items():
            n = int(np.prod(shape)) if shape else 1
            dense[off:off + n] = torch.as_tensor(np.asarray(sd[name]), dtype=torch.float32).reshape(-1)
        self.dense.copy_(dense)

    def save(self, sess=None):
        """Reference Model.save (model.py:302-308): <model_dir>/TLSAN-<step> + config JSON."""
        os.makedirs(self.config["model_dir"], exist_ok=True)
        checkpoint_path = os.path.join(self.config["model_dir"], "TLSAN")
        step = self.global_step.eval()
        save_path = "%s-%d" % (checkpoint_path, step)
        ck = {"variables": OrderedDict(self.state_dict()), "global_step": int(step),
              "global_epoch_step": int(self.global_epoch_step.eval())}
        if self._slots is not None:           # tf.train.Saver stores the optimizer's slot variables too
            ck["optimizer"] = self.optimizer
            ck["opt_step"] = int(self._opt.step)
            ck["slot1"], ck["slot2"] = self._slots[0].cpu(), self._slots[1].cpu()
        torch.save(ck, save_path)
        json.dump(dict(self.config), open("%s-%d.json" % (checkpoint_path, step), "w"), indent=2)
        print("model saved at %s" % save_path, flush=True)
        return save_path

    def restore(self, sess, path):
        """Reference Model.restore (model.py:310-313)."""
        ck = torch.load(path, map_location="cpu", weights_only=True)     # tensors + ints only: no unpickling of objects
        self.load_state_dict(ck["variables"])
        self.global_step.value = int(ck["global_step"])
        self.global_epoch_step.value = int(ck["global_epoch_step"])
        if self._slots is not None and ck.get("optimizer") == self.optimizer:
            self._slots[0].copy_(ck["slot1"]); self._slots[1].copy_(ck["slot2"])
            self._opt.step = int(ck["opt_step"])
        print("model restored from %s" % path, flush=True)
