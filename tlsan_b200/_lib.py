"""ctypes binding of include/tlsan_b200.h.  No CPU fallback: if the CUDA library is missing
or fails to load, importing the compute entry points raises."""
import atexit
import ctypes as C
import os

from .build import LIB

DENSE_COUNT, DENSE_PAD, STAT_COUNT, MAX_L = 4449, 4452, 8, 96
PART, SHARD_ROW = 4456, 36
OFF = dict(W1L=0, B1L=64, W2L=72, B2L=136, W1S=144, B1S=208, W2S=216, B2S=280, WD=288, BD=4384, GAMMA=4448)
STAT = dict(loss=0, bce=1, norm=2, scale=3, l2=4, dp_err=5)


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "L", "S", "NI", "NU", "NC", "B_global", "reserved")]


class Params(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("emb", "usert", "item_b", "dense", "icl", "cate_off", "cate_items")]


class Batch(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("u", "i", "i2", "y", "hist_i", "hist_i_new", "hist_t", "sl", "sl_new", "c", "hist_d")]


class Next(C.Structure):
    _fields_ = [("dims", C.POINTER(Dims)), ("batch", C.POINTER(Batch)), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t), ("ready_event", C.c_void_p)]


class Opt(C.Structure):
    _fields_ = [("kind", C.c_int32), ("step", C.c_int32), ("beta1", C.c_float), ("beta2", C.c_float), ("rho", C.c_float),
                ("momentum", C.c_float), ("epsilon", C.c_float), ("slot1", C.c_void_p), ("slot2", C.c_void_p)]


OPT_KIND = dict(sgd=0, adam=1, rmsprop=2, adadelta=3)


class Dataset(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("uid", "pre_off", "pre_items", "pre_time", "new_off", "new_items", "cand",
                                          "second_i", "second_f", "ucate")] + [("n", C.c_int64)]


class TlsanError(RuntimeError):
    pass


_lib = None

_SIGS = {
    "tlsan_abi_version": (C.c_int, []),
    "tlsan_last_error": (C.c_char_p, []),
    "tlsan_time_bucket": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tlsan_gather_concat": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_void_p]),
    "tlsan_score": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.POINTER(Batch), C.c_int32, C.c_void_p,
                              C.c_void_p, C.c_void_p]),
    "tlsan_score_workspace_bytes": (C.c_int, [C.POINTER(Dims), C.POINTER(C.c_size_t)]),
    "tlsan_score_ws": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.POINTER(Batch), C.c_int32, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tlsan_workspace_bytes": (C.c_int, [C.POINTER(Dims), C.POINTER(C.c_size_t)]),
    "tlsan_train_step": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.POINTER(Batch), C.c_float, C.c_float,
                                   C.c_float, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "tlsan_train_step_pipelined": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.POINTER(Batch), C.POINTER(Next),
                                             C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_size_t, C.c_void_p,
                                             C.c_void_p]),
    "tlsan_step_grads_pipelined": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.POINTER(Batch), C.POINTER(Next),
                                             C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "tlsan_flat_count": (C.c_int, [C.POINTER(Dims), C.POINTER(C.c_int64)]),
    "tlsan_step_grads": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.POINTER(Batch), C.c_void_p, C.c_size_t,
                                   C.c_void_p, C.c_void_p]),
    "tlsan_apply_flat": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.c_void_p, C.c_float, C.c_float, C.c_float,
                                   C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "tlsan_apply_flat_opt": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.c_void_p, C.c_float, C.c_float, C.c_float,
                                       C.POINTER(Opt), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "tlsan_dp_arena_bytes": (C.c_int, [C.POINTER(Dims), C.c_int32, C.POINTER(C.c_size_t)]),
    "tlsan_dp_arena_create": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    "tlsan_dp_arena_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "tlsan_dp_arena_release": (C.c_int, [C.c_void_p, C.c_int32]),
    "tlsan_dp_exchange": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.POINTER(C.c_void_p), C.c_int32,
                                    C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_size_t,
                                    C.c_void_p, C.c_void_p]),
    "tlsan_label_rank": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "tlsan_rank_workspace_bytes": (C.c_int, [C.POINTER(Dims), C.POINTER(C.c_size_t)]),
    "tlsan_label_rank_ws": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    "tlsan_label_rank_shard": (C.c_int, [C.c_int32, C.c_int64] + [C.c_void_p] * 7 + [C.c_int32, C.c_int32, C.c_void_p,
                                                                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "tlsan_pack_batch_host": (C.c_int, [C.POINTER(Dims)] + [C.c_void_p] * 11 + [C.c_int64, C.c_int32, C.c_int32]),
    "tlsan_pack_batch_host_i32": (C.c_int, [C.POINTER(Dims)] + [C.c_void_p] * 11 + [C.c_int64, C.c_int32, C.c_int32]),
    "tlsan_stage_batch_host_i32": (C.c_int, [C.POINTER(Dims)] + [C.c_void_p] * 12 + [C.c_int64, C.c_int32, C.c_int32,
                                                                                      C.c_void_p]),
    "tlsan_stage_words": (C.c_int, [C.POINTER(Dims), C.POINTER(C.c_int64)]),
    "tlsan_stage_batch_host": (C.c_int, [C.POINTER(Dims)] + [C.c_void_p] * 12 + [C.c_int64, C.c_int32, C.c_int32,
                                                                                  C.c_void_p]),
    "tlsan_stage_packed": (C.c_int, [C.POINTER(Dims), C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "tlsan_collate": (C.c_int, [C.POINTER(Dataset), C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                C.c_int64, C.c_void_p]),
    "tlsan_ds_plan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tlsan_ds_lengths": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 7),
    "tlsan_ds_emit": (C.c_int, [C.c_void_p] * 5 + [C.c_int32] + [C.c_void_p] * 7 + [C.POINTER(Dataset), C.POINTER(Dataset),
                                                                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "tlsan_shard_pack_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "tlsan_shard_unpack_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    "tlsan_shard_pack_grads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "tlsan_shard_accum_grads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tlsan_shard_apply_grads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                          C.c_void_p]),
    "tlsan_route_bitmap_words": (C.c_int, [C.c_int64, C.c_int32, C.POINTER(C.c_int64)]),
    "tlsan_route_ids": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int32, C.c_int64,
                                  C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "tlsan_shard_apply_replicated": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                               C.c_size_t, C.c_void_p, C.c_void_p]),
    "tlsan_reduce_cate": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p]),
    "tlsan_sgd_dense": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "tlsan_sumsq": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "tlsan_launch_count": (C.c_longlong, []),
    "tlsan_profile_begin": (C.c_int, [C.c_int32]),
    "tlsan_profile_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
}
EXPORTS = tuple(_SIGS)
PHASES = ("sort", "long_fwd", "dense_fwd", "short", "dense_bwd", "bwd_long", "reduce", "apply")
# kernel that dominates each phase (names as ncu prints them, default `pf` formulation)
PHASE_KERNEL = {"long_fwd": "k_pf_long<1>", "short": "k_pf_short", "bwd_long": "k_pf_long<3>",
                "dense_fwd": "k_dense_fwd_mma", "dense_bwd": "k_dense_bwd_mma", "reduce": "k_row_reduce_bal"}


def _drain():
    """At interpreter exit: let the kernels still queued on the library's side streams (a presort announced for a
    batch nobody trained on, table norms) finish before torch tears its allocations down under them."""
    try:
        import torch
        if torch.cuda.is_available():
            for d in range(torch.cuda.device_count()):
                torch.cuda.synchronize(d)
    except Exception:
        pass


def lib():
    """Load tlsan_b200/libtlsan_b200.so (built by tlsan_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise TlsanError("CUDA library %s is missing: run `python -m tlsan_b200.build` "
                             "(there is no CPU fallback)" % LIB)
        h = C.CDLL(LIB)
        for name, (res, args) in _SIGS.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        if h.tlsan_abi_version() != 2:
            raise TlsanError("ABI version mismatch")
        _lib = h
        atexit.register(_drain)
    return _lib


def check(rc):
    if rc != 0:
        raise TlsanError("tlsan_b200 error %d: %s" % (rc, lib().tlsan_last_error().decode()))
