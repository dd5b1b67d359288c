"""Row-sharded item tables (BASELINE config 5: a 10 M-item catalogue on 8 x B200).

``item_emb`` (NI x 32), ``item_b`` and ``item_cate_list`` are split by row over the ranks of a process
group; ``cate_emb``, ``user_emb``, ``usert_emb`` and the 4 449 small parameters stay replicated.  The reference
has no counterpart (its tables live in one TF process, TLSAN/model.py:56-81); the arithmetic of a step is the
same as ``Model.train`` -- the weights after a step equal the replicated model's up to fp32 summation order
(tests/test_gpu_sharded.py).

One step on every rank:

1. distinct item ids of the local batch (``hist_i``, ``hist_i_new``, ``i`` [, ``i2``]) ON THE DEVICE: a presence
   bitmap over owner-major positions, its popcount prefix, per-owner request lists of fixed capacity ``cap``
   (tlsan_route_ids; no torch.unique, no host round trip -- the capacity is calibrated once on the first batch)
2. NCCL all-to-all (equal splits) of the request lists to the owners, owners answer with 144-B exchange rows
   (item_emb row | item_b | icl), all-to-all back                        (tlsan_shard_pack_rows / _unpack_rows)
3. the rows land in a COMPACT table (row owner * cap + rank of the id inside the owner's group; category and user
   rows follow at a fixed offset), the batch ids are rewritten to compact rows, and the unchanged fused kernels
   run on it (``tlsan_step_grads``: sort, forward, backward, deterministic segmented reduce)
4. all-reduce (sum) of [user/usert gradients | small-parameter gradients, loss and norm partials | category
   gradients | partial sums of squares of the item shards]                                    -- ONE collective
5. per-id gradient rows travel back to the owners (same equal splits); an owner first applies the L2 decay
   W <- W (1 - lr*scale*reg) to every row of its shard (the L2 term touches all rows, model.py:164-169), then
   subtracts lr*scale*g for the received rows, source ranks in rank order (deterministic)
                                                                          (tlsan_sgd_dense, tlsan_shard_apply_grads)
6. the replicated tables are updated identically on every rank                 (tlsan_shard_apply_replicated)

PyTorch supplies device memory, one ``cumsum`` (the popcount prefix), one stable ``sort`` (category CSR of the
compact table) and the NCCL collectives; every gather, scatter, reduction and update of table data is a kernel of
libtlsan_b200.so.
There is no CPU path.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import PART, SHARD_ROW, STAT, Batch, Dims, Params, check
from .model import DENSE_LAYOUT, KS, DeviceBatch, Model, _pack_offsets, pack_batch

NSQ = 256          # partial sums of squares per item shard


def owner_of(ids, world, NI, partition):
    """Owner rank of each global item id.  ``mod``: id % world (cyclic); ``block``: contiguous blocks of
    ceil(NI / world) rows."""
    if partition == "mod":
        return ids % world
    return ids // (-(-NI // world))


def local_of(ids, world, NI, partition):
    """Row index inside the owner's shard."""
    if partition == "mod":
        return ids // world
    return ids % (-(-NI // world))


def shard_ids(rank, world, NI, partition):
    """Global ids of the rows of ``rank``'s shard, in shard order (numpy int64)."""
    if partition == "mod":
        return np.arange(rank, NI, world, dtype=np.int64)
    blk = -(-NI // world)
    return np.arange(min(rank * blk, NI), min((rank + 1) * blk, NI), dtype=np.int64)


def exchange(send, in_splits, out_splits, group, world):
    """all_to_all_single with explicit row splits (rows of ``send`` grouped by destination rank)."""
    if world == 1:
        return send
    recv = torch.empty((int(sum(out_splits)),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send.contiguous(), list(out_splits), list(in_splits), group=group)
    return recv


def route_ids(uniq, world, NI, partition, group):
    """Step (2a): send every distinct id to its owner.  Returns (order, recv_ids, in_splits, out_splits):
    ``uniq[order]`` is the send buffer (grouped by owner, ids ascending inside a group), ``recv_ids`` the ids
    this rank must serve, grouped by requesting rank."""
    owner = owner_of(uniq, world, NI, partition)
    order = torch.sort(owner, stable=True).indices
    counts = torch.bincount(owner, minlength=world)
    if world == 1:
        n = [int(uniq.numel())]
        return order, uniq[order], n, n
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    in_splits = [int(x) for x in counts.tolist()]
    out_splits = [int(x) for x in recv_counts.tolist()]
    recv_ids = exchange(uniq[order], in_splits, out_splits, group, world)
    return order, recv_ids, in_splits, out_splits


class ShardedModel(object):
    """TLSAN with row-sharded item tables.  Same train / eval_auc surface as ``Model``."""

    def __init__(self, config, item_cate_list, process_group=None, partition="mod", seed=1234, device=None,
                 validate=True, route_capacity=None):
        if partition not in ("mod", "block"):
            raise ValueError("partition must be 'mod' or 'block'")
        Model._check_config(config)
        if config.get("optimizer", "sgd") in ("adadelta", "adam", "rmsprop"):
            raise ValueError("row-sharded tables implement the reference default optimizer 'sgd' only (train.py:40)")
        self._lib = _lib.lib()
        if not torch.cuda.is_available():
            raise _lib.TlsanError("tlsan_b200 needs a CUDA device (no CPU fallback)")
        self.config = config
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if process_group is not None else 1
        self.rank = dist.get_rank(process_group) if process_group is not None else 0
        self.partition = partition
        self.NI, self.NU, self.NC = int(config["item_count"]), int(config["user_count"]), int(config["cate_count"])
        self.L = int(config["Ls"])
        self.PU = (32 + self.L + 3) // 4 * 4
        self.reg = float(config.get("regulation_rate", 0.00005))
        self.clip = float(config.get("max_gradient_norm", 5.0))
        self.validate = validate
        dev = self.device

        mine = shard_ids(self.rank, self.world, self.NI, partition)
        self.n_local = int(mine.size)
        icl = np.asarray(item_cate_list)
        if icl.shape != (self.NI,) or icl.min() < 0 or icl.max() >= self.NC:
            raise ValueError("item_cate_list must be int[item_count] with values in [0, cate_count)")
        self.icl_shard = torch.from_numpy(np.ascontiguousarray(icl[mine].astype(np.int32))).to(dev)
        # glorot_uniform like tf.get_variable (model.py:62-81); one stream per rank for the shard
        g = torch.Generator(device=dev).manual_seed(seed + 7919 * self.rank)
        lim = (6.0 / (self.NI + 32)) ** 0.5
        self.item_emb_shard = (torch.rand(max(self.n_local, 1), 32, generator=g, device=dev) * 2 - 1) * lim
        self.item_b_shard = torch.zeros(max(self.n_local, 1), device=dev)
        gr = torch.Generator().manual_seed(seed)              # replicated part: same on every rank

        def glorot(rows, cols):
            return (torch.rand(rows, cols, generator=gr) * 2 - 1) * (6.0 / (rows + cols)) ** 0.5
        self._tail = torch.cat([glorot(self.NC, 32), glorot(self.NU, 32)]).to(dev)   # cate rows, then user rows
        self.usert_emb = torch.full((self.NU, self.L), -1.0, device=dev)
        dense = torch.zeros(_lib.DENSE_PAD)
        for name, (off, shape) in DENSE_LAYOUT.items():
            if name == "gamma_parameter":
                dense[off] = 1.0
            elif len(shape) == 2:
                dense[off:off + shape[0] * shape[1]] = glorot(*shape).reshape(-1)
        self.dense = dense.to(dev)

        self.cap = 0                         # rows of the compact table = world * per-owner capacity
        self.route_cap = int(route_capacity) if route_capacity else 0      # 0: calibrated on the first batch
        self._grow(1024)
        nw = C.c_int64()
        check(self._lib.tlsan_route_bitmap_words(self.NI, self.world, C.byref(nw)))
        self._bits = torch.zeros(int(nw.value), dtype=torch.int32, device=dev)
        self._wpref = torch.zeros(int(nw.value), dtype=torch.int32, device=dev)
        self._counts = torch.zeros(self.world, dtype=torch.int32, device=dev)
        self._overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self._ovf_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._xbuf = {}
        self._bad = torch.zeros(1, dtype=torch.int32, device=dev)
        self._stats = torch.zeros(_lib.STAT_COUNT, device=dev)
        self._ws = None
        self._score_ws = None
        self._rank_ws = None
        self._flat = None
        self._stage_cache = {}
        self.reset_metrics()
        self.global_step = 0
        self.last_exchange_bytes = 0
        self.profile = None          # set to [] to collect (label, cuda event) marks per step (tools/bench_sharded.py)

    # ------------------------------------------------------------------ compact table
    def _grow(self, need):
        """Compact table = [cap item slots | NC cate rows | NU user rows]; the tail holds the master copy of the
        replicated cate_emb / user_emb, so growing the item capacity moves it."""
        cap = max(1024, -(-int(need) // 1024) * 1024)
        if cap <= self.cap:
            return
        dev = self.device
        tail = self._tail if self.cap == 0 else self.emb_c[self.cap:]
        emb_c = torch.zeros(cap + self.NC + self.NU, 32, device=dev)
        emb_c[cap:] = tail
        self.emb_c, self.cap, self._tail = emb_c, cap, None
        self.item_b_c = torch.zeros(cap, device=dev)
        self.icl_c = torch.zeros(cap, dtype=torch.int32, device=dev)
        self.cate_off = torch.zeros(self.NC + 1, dtype=torch.int32, device=dev)
        self.cate_items = torch.zeros(cap, dtype=torch.int32, device=dev)
        self._params = Params(emb=self.emb_c.data_ptr(), usert=self.usert_emb.data_ptr(),
                              item_b=self.item_b_c.data_ptr(), dense=self.dense.data_ptr(),
                              icl=self.icl_c.data_ptr(), cate_off=self.cate_off.data_ptr(),
                              cate_items=self.cate_items.data_ptr())
        self._ws = None
        self._flat = None

    @property
    def cate_emb(self):
        return self.emb_c[self.cap:self.cap + self.NC]

    @property
    def user_emb(self):
        return self.emb_c[self.cap + self.NC:]

    def _dims(self, B, S, B_global=None):
        return Dims(B=B, L=self.L, S=S, NI=self.cap, NU=self.NU, NC=self.NC,
                    B_global=int(B_global if B_global is not None else B), reserved=1)   # norms: see train_staged

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _mark(self, label):
        if self.profile is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.profile.append((label, ev))

    # ------------------------------------------------------------------ staging
    def stage_batch(self, batch, is_test=False):
        """Host 9-tuple (GLOBAL item ids) -> device, like Model.stage_batch; ids are range-checked against
        the global item_count."""
        B = len(batch[0])
        S = max(int(np.shape(batch[4])[1]), 1)
        if np.shape(batch[3])[1] != self.L:
            raise ValueError("hist_i has %d columns but the model was built with Ls=%d" % (np.shape(batch[3])[1], self.L))
        offs, total = _pack_offsets(B, self.L, S)
        gd = Dims(B=B, L=self.L, S=S, NI=self.NI, NU=self.NU, NC=self.NC, B_global=B, reserved=0)
        words = C.c_int64()
        check(self._lib.tlsan_stage_words(C.byref(gd), C.byref(words)))
        words = int(words.value)
        key = (B, S)
        if key not in self._stage_cache:
            self._stage_cache[key] = (torch.empty(words, dtype=torch.int32).pin_memory(), torch.cuda.Event())
        host, ev = self._stage_cache[key]
        ev.synchronize()
        dev = torch.empty(words, dtype=torch.int32, device=self.device)
        pack_batch(self._lib, batch, gd, is_test, host.numpy(), self.validate, dev.data_ptr(), self._stream())
        ev.record(torch.cuda.current_stream(self.device))
        return DeviceBatch(dev, B, self.L, S, offs, is_test)

    def _id_fields(self, db):
        f = [("hist_i", db.B * db.L), ("hist_i_new", db.B * db.S), ("i", db.B)]
        if db.is_test:
            f.append(("second", db.B))
        return f

    def _calibrate(self, db):
        """Per-owner capacity of the request lists, once: 1.25 x the largest group of the first batch (max over
        ranks), rounded up to 1024.  The only place a batch is inspected on the host."""
        ids = torch.cat([db.buf[db.offs[k]:db.offs[k] + n] for k, n in self._id_fields(db)])
        uniq = torch.unique(ids.to(torch.int64))
        cnt = torch.bincount(owner_of(uniq, self.world, self.NI, self.partition), minlength=self.world).max().reshape(1)
        if self.world > 1:
            dist.all_reduce(cnt, op=dist.ReduceOp.MAX, group=self.pg)
        nloc = -(-self.NI // self.world)
        self.route_cap = min(max(1024, -(-int(1.25 * int(cnt.item())) // 1024) * 1024), -(-nloc // 1024) * 1024)

    def _buf(self, name, shape, dtype=torch.float32):
        t = self._xbuf.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._xbuf[name] = t
        return t

    def _a2a(self, name, send):
        if self.world == 1:
            return send
        recv = self._buf(name, send.shape, send.dtype)
        dist.all_to_all_single(recv, send, group=self.pg)
        return recv

    def _fetch(self, db):
        """Steps (1)-(3a): returns (compact DeviceBatch, owner-side request lists [world * cap])."""
        lib, st = self._lib, self._stream()
        if int(self._ovf_host[0]):
            raise _lib.TlsanError("a rank's request list exceeded route_capacity=%d in an earlier step; construct "
                                  "ShardedModel(..., route_capacity=N) with a larger N" % self.route_cap)
        if not self.route_cap:
            self._calibrate(db)
        W, cap = self.world, self.route_cap
        self._grow(W * cap)
        fields = self._id_fields(db)
        cbuf = db.buf.clone()
        nf = len(fields)
        src = (C.c_void_p * 4)(*([db.buf.data_ptr() + 4 * db.offs[k] for k, _ in fields] + [None] * (4 - nf)))
        dst = (C.c_void_p * 4)(*([cbuf.data_ptr() + 4 * db.offs[k] for k, _ in fields] + [None] * (4 - nf)))
        cnt = (C.c_int64 * 4)(*([n for _, n in fields] + [0] * (4 - nf)))
        send_ids = self._buf("send_ids", (W * cap,), torch.int32)
        slot_row = self._buf("slot_row", (W * cap,), torch.int32)
        args = (src, dst, cnt, nf, self.NI, W, 1 if self.partition == "mod" else 0, cap, self._bits.data_ptr(),
                self._wpref.data_ptr(), send_ids.data_ptr(), slot_row.data_ptr(), self._counts.data_ptr(),
                self._overflow.data_ptr())
        check(lib.tlsan_route_ids(*args, 0, st))
        self._wpref.cumsum_(0)                                       # inclusive prefix of the word popcounts
        check(lib.tlsan_route_ids(*args, 1, st))
        self._ovf_host.copy_(self._overflow, non_blocking=True)      # looked at by a LATER call: never waited for
        recv_ids = self._a2a("recv_ids", send_ids)
        rows_out = self._buf("rows_out", (W * cap, SHARD_ROW))
        check(lib.tlsan_shard_pack_rows(self.item_emb_shard.data_ptr(), self.item_b_shard.data_ptr(),
                                        self.icl_shard.data_ptr(), recv_ids.data_ptr(), W * cap, self.n_local,
                                        rows_out.data_ptr(), self._bad.data_ptr(), st))
        rows_in = self._a2a("rows_in", rows_out)
        check(lib.tlsan_shard_unpack_rows(rows_in.data_ptr(), slot_row.data_ptr(), W * cap, self.emb_c.data_ptr(),
                                          self.item_b_c.data_ptr(), self.icl_c.data_ptr(), st))
        cb = DeviceBatch(cbuf, db.B, db.L, db.S, db.offs, db.is_test)
        self.last_exchange_bytes = W * cap * (4 + 4 * SHARD_ROW)
        self._n_u = self._wpref[-1:]                               # distinct ids of this batch (device scalar)
        return cb, recv_ids

    @property
    def last_unique(self):
        """Distinct item ids of the last batch on this rank (reads the device counters: synchronises)."""
        return int(self._counts.sum().item())

    def _compact_csr(self):
        """Items of the compact table grouped by category (stable), for the hierarchical category reduce."""
        # rows >= n_u hold stale data of earlier steps: they go to a dummy category NC behind the real ones
        rows = torch.arange(self.cap, device=self.device, dtype=torch.int32)
        icl = torch.where(rows < self._n_u, self.icl_c[:self.cap], self.NC)
        self.cate_items[:self.cap] = torch.sort(icl, stable=True).indices.to(torch.int32)
        self.cate_off[1:] = torch.cumsum(torch.bincount(icl, minlength=self.NC + 1)[:self.NC], 0).to(torch.int32)

    # ------------------------------------------------------------------ training
    def train_staged(self, db, lr, global_batch=None):
        lib, st = self._lib, self._stream()
        self._mark("start")
        cb, recv_ids = self._fetch(db)
        W, cap = self.world, self.route_cap
        self._compact_csr()
        self._mark("fetch_rows")
        Bg = global_batch if global_batch is not None else db.B * self.world
        dims = self._dims(db.B, db.S, Bg)
        need = C.c_size_t()
        check(lib.tlsan_workspace_bytes(C.byref(dims), C.byref(need)))
        if self._ws is None or self._ws.numel() < need.value:
            self._ws = torch.empty(int(need.value), dtype=torch.uint8, device=self.device)
        f_gb = (self.cap + self.NC) * 64
        f_gu = f_gb + (self.cap + 3) // 4 * 4
        f_dgrad = f_gu + self.NU * self.PU
        n_flat = C.c_int64()
        check(lib.tlsan_flat_count(C.byref(dims), C.byref(n_flat)))
        assert n_flat.value == f_dgrad + PART
        f_cate, f_sq = f_dgrad + PART, f_dgrad + PART + self.NC * 32
        total = f_sq + NSQ
        if self._flat is None or self._flat.numel() != total:
            self._flat = torch.empty(total, dtype=torch.float32, device=self.device)
        flat = self._flat
        fp = lambda off: flat.data_ptr() + 4 * off
        check(lib.tlsan_step_grads(C.byref(dims), C.byref(self._params), C.byref(cb.c), self._ws.data_ptr(),
                                   self._ws.numel(), flat.data_ptr(), st))
        check(lib.tlsan_reduce_cate(C.byref(dims), C.byref(self._params), flat.data_ptr(), fp(f_cate), st))
        check(lib.tlsan_sumsq(self.item_emb_shard.data_ptr(), self.n_local * 32, fp(f_sq), NSQ, st))
        self._mark("step_grads")
        if self.world > 1:
            dist.all_reduce(flat[f_gu:], group=self.pg)
        self._mark("allreduce")
        # gradient rows of the requested ids -> owners (reverse of the row exchange, same equal splits)
        grads_out = self._buf("grads_out", (W * cap, SHARD_ROW))
        check(lib.tlsan_shard_pack_grads(flat.data_ptr(), fp(f_gb), self._xbuf["slot_row"].data_ptr(), W * cap,
                                         grads_out.data_ptr(), st))
        grads_in = self._a2a("grads_in", grads_out)
        self._mark("return_grads")
        check(lib.tlsan_shard_apply_replicated(C.byref(dims), C.byref(self._params), fp(f_cate), fp(f_gu), fp(f_dgrad),
                                               fp(f_sq), NSQ, lr, self.reg, self.clip, self._ws.data_ptr(),
                                               self._ws.numel(), self._stats.data_ptr(), st))
        scale = self._stats.data_ptr() + 4 * STAT["scale"]
        # L2 decay of every shard row, then the received gradient rows, source ranks in rank order
        check(lib.tlsan_sgd_dense(self.item_emb_shard.data_ptr(), None, self.n_local * 32, lr, self.reg, scale, st))
        for o in range(W):
            check(lib.tlsan_shard_apply_grads(grads_in.data_ptr() + 4 * SHARD_ROW * o * cap,
                                              recv_ids.data_ptr() + 4 * o * cap, cap, self.item_emb_shard.data_ptr(),
                                              self.item_b_shard.data_ptr(), lr, scale, st))
        self._mark("apply")
        self.global_step += 1
        return self._stats

    def train(self, sess, batch, lr, add_summary=False):
        """Model.train (model.py:208-234) on this rank's rows of the global batch."""
        stats = self.train_staged(self.stage_batch(batch), float(lr))
        if int(self._bad.item()):
            raise _lib.TlsanError("an item id was routed to a rank that does not own it")
        if int(self._overflow.item()):
            raise _lib.TlsanError("a request list exceeded route_capacity=%d" % self.route_cap)
        return float(stats[STAT["loss"]].item())

    # ------------------------------------------------------------------ scoring
    def score_staged(self, db, ncand=1, want_ut=False, want_compact=False):
        cb, _ = self._fetch(db)
        dims = self._dims(db.B, db.S)
        logits = torch.empty(db.B, ncand, dtype=torch.float32, device=self.device)
        ut = torch.empty(db.B, 64, dtype=torch.float32, device=self.device) if want_ut else None
        utp = ut.data_ptr() if want_ut else None
        if db.B >= 2048:                                       # same kernel selection as Model.score_staged
            need = C.c_size_t()
            check(self._lib.tlsan_score_workspace_bytes(C.byref(dims), C.byref(need)))
            if self._score_ws is None or self._score_ws.numel() < need.value:
                self._score_ws = torch.empty(int(need.value), dtype=torch.uint8, device=self.device)
            check(self._lib.tlsan_score_ws(C.byref(dims), C.byref(self._params), C.byref(cb.c), ncand,
                                           logits.data_ptr(), utp, self._score_ws.data_ptr(),
                                           self._score_ws.numel(), self._stream()))
        else:
            check(self._lib.tlsan_score(C.byref(dims), C.byref(self._params), C.byref(cb.c), ncand,
                                        logits.data_ptr(), utp, self._stream()))
        if want_compact:
            return logits, ut, cb
        return (logits, ut) if want_ut else logits

    # ------------------------------------------------------------------ full-catalogue P@k / R@k (model.py:140-156)
    def _label_ranks(self, batch):
        """rank of batch[1] among ALL items for this rank's rows: every rank scores its rows (u_t), the rows / labels
        / label rows are all-gathered, every shard counts the items of ITS rows that beat each label on the tensor
        cores (tlsan_label_rank_shard), and one integer all-reduce adds the shards up (last row)."""
        db = self.stage_batch(batch, is_test=True)
        _, ut, cb = self.score_staged(db, 1, want_ut=True, want_compact=True)
        B = db.B
        o = db.offs["i"]
        lab = db.buf[o:o + B].contiguous()                                # global ids
        ci = cb.buf[o:o + B].long()                                       # their compact rows
        lab_rows = torch.zeros(B, 68, device=self.device)
        lab_rows[:, :32] = self.emb_c[ci]
        lab_rows[:, 32:64] = self.cate_emb[self.icl_c[ci].long()]
        lab_rows[:, 64] = self.item_b_c[ci]
        W = self.world
        if W > 1:
            nb = torch.tensor([B], device=self.device)
            nbs = [torch.empty_like(nb) for _ in range(W)]
            dist.all_gather(nbs, nb, group=self.pg)
            nbs = [int(x.item()) for x in nbs]
            Bm = max(nbs)

            def gather(t):
                pad = torch.zeros((Bm,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
                pad[:B] = t
                out = [torch.empty_like(pad) for _ in range(W)]
                dist.all_gather(out, pad, group=self.pg)
                return torch.cat(out)
            ut_all, lab_all, rows_all = gather(ut), gather(lab), gather(lab_rows)
        else:
            nbs, Bm, ut_all, lab_all, rows_all = [B], B, ut, lab, lab_rows
        n_all = W * Bm
        part = torch.zeros(n_all, dtype=torch.int32, device=self.device)
        if self.n_local:
            need = (-(-self.n_local // 128)) * 73728 + 512
            if self._rank_ws is None or self._rank_ws.numel() < need:
                self._rank_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            mul, add = (W, self.rank) if self.partition == "mod" else (1, self.rank * (-(-self.NI // W)))
            check(self._lib.tlsan_label_rank_shard(n_all, self.n_local, self.item_emb_shard.data_ptr(),
                                                   self.item_b_shard.data_ptr(), self.icl_shard.data_ptr(),
                                                   self.cate_emb.data_ptr(), ut_all.data_ptr(), lab_all.data_ptr(),
                                                   rows_all.data_ptr(), mul, add, part.data_ptr(),
                                                   self._rank_ws.data_ptr(), self._rank_ws.numel(), self._stream()))
        if W > 1:
            dist.all_reduce(part, group=self.pg)
        return part[self.rank * Bm:self.rank * Bm + B].cpu().numpy()

    def eval_prec(self, sess, batch):
        """Model.eval_prec (model.py:265-281) on this rank's rows; cumulative like the reference (reset_metrics)."""
        rank = self._label_ranks(batch)
        for n, k in enumerate(KS):
            hit = float(np.sum(rank < k))
            self._ptp[n] += hit
            self._pfp[n] += len(rank) * k - hit
        return [float(self._ptp[n] / (self._ptp[n] + self._pfp[n])) for n in range(len(KS))]

    def eval_recall(self, sess, batch):
        """Model.eval_recall (model.py:283-299) on this rank's rows."""
        rank = self._label_ranks(batch)
        for n, k in enumerate(KS):
            hit = float(np.sum(rank < k))
            self._rtp[n] += hit
            self._rfn[n] += len(rank) - hit
        return [float(self._rtp[n] / (self._rtp[n] + self._rfn[n])) for n in range(len(KS))]

    def reset_metrics(self):
        self._ptp = np.zeros(len(KS)); self._pfp = np.zeros(len(KS))
        self._rtp = np.zeros(len(KS)); self._rfn = np.zeros(len(KS))

    def eval_auc(self, sess, batch):
        """Model.eval_auc (model.py:237-263) on this rank's rows."""
        res = self.score_staged(self.stage_batch(batch, is_test=True), 2).cpu().numpy()
        return np.mean(res[:, 0] - res[:, 1] > 0)

    # ------------------------------------------------------------------ state
    def load_full_state(self, sd):
        """Take this rank's rows of a full (replicated-model) state dict -- tests and checkpoints."""
        mine = torch.from_numpy(shard_ids(self.rank, self.world, self.NI, self.partition))
        if self.n_local:
            self.item_emb_shard[:self.n_local] = torch.as_tensor(np.asarray(sd["item_emb"]), dtype=torch.float32)[mine].to(self.device)
            self.item_b_shard[:self.n_local] = torch.as_tensor(np.asarray(sd["item_b"]), dtype=torch.float32)[mine].to(self.device)
        self.cate_emb.copy_(torch.as_tensor(np.asarray(sd["cate_emb"]), dtype=torch.float32))
        self.user_emb.copy_(torch.as_tensor(np.asarray(sd["user_emb"]), dtype=torch.float32))
        self.usert_emb.copy_(torch.as_tensor(np.asarray(sd["usert_emb"]), dtype=torch.float32))
        dense = self.dense.detach().cpu().clone()
        for name, (off, shape) in DENSE_LAYOUT.items():
            n = int(np.prod(shape)) if shape else 1
            dense[off:off + n] = torch.as_tensor(np.asarray(sd[name]), dtype=torch.float32).reshape(-1)
        self.dense.copy_(dense)

    def gather_full_state(self):
        """Reassemble the full state dict on every rank (all_gather of the shards)."""
        sd = {}
        mine = torch.from_numpy(shard_ids(self.rank, self.world, self.NI, self.partition)).to(self.device)
        pad = -(-self.NI // self.world)
        emb = torch.zeros(pad, 33, device=self.device)
        ids = torch.full((pad,), -1, dtype=torch.int64, device=self.device)
        emb[:self.n_local, :32] = self.item_emb_shard[:self.n_local]
        emb[:self.n_local, 32] = self.item_b_shard[:self.n_local]
        ids[:self.n_local] = mine
        if self.world > 1:
            embs = [torch.empty_like(emb) for _ in range(self.world)]
            idss = [torch.empty_like(ids) for _ in range(self.world)]
            dist.all_gather(embs, emb, group=self.pg)
            dist.all_gather(idss, ids, group=self.pg)
            emb, ids = torch.cat(embs), torch.cat(idss)
        keep = ids >= 0
        full = torch.zeros(self.NI, 33, device=self.device)
        full[ids[keep]] = emb[keep]
        sd["item_emb"] = full[:, :32].cpu()
        sd["item_b"] = full[:, 32].cpu()
        sd["user_emb"] = self.user_emb.detach().cpu().clone()
        sd["usert_emb"] = self.usert_emb.detach().cpu().clone()
        sd["cate_emb"] = self.cate_emb.detach().cpu().clone()
        dense = self.dense.detach().cpu()
        for name, (off, shape) in DENSE_LAYOUT.items():
            n = int(np.prod(shape)) if shape else 1
            sd[name] = dense[off:off + n].reshape(shape).clone()
        return sd
