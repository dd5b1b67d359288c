"""The C-ABI library: builds, loads, exports every symbol include/tlsan_b200.h declares, and
rejects bad arguments before touching the GPU (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from tlsan_b200 import _lib
from tlsan_b200.build import LIB, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build()
    return _lib.lib()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "tlsan_b200.h")).read()
    declared = set(re.findall(r"\b(tlsan_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    raw = C.CDLL(LIB)
    for name in sorted(declared):
        assert hasattr(raw, name), "missing export %s" % name
    assert declared == set(_lib.EXPORTS)
    assert lib.tlsan_abi_version() == 2


def test_constants_match_header():
    hdr = open(os.path.join(ROOT, "include", "tlsan_b200.h")).read()
    d = dict(re.findall(r"#define TLSAN_(\w+) (\d+)", hdr))
    assert int(d["DENSE_COUNT"]) == _lib.DENSE_COUNT and int(d["DENSE_PAD"]) == _lib.DENSE_PAD
    for k, v in _lib.OFF.items():
        assert int(d["OFF_" + k]) == v
    assert int(d["MAX_L"]) == _lib.MAX_L
    assert C.sizeof(_lib.Dims) == 32 and C.sizeof(_lib.Params) == 56 and C.sizeof(_lib.Batch) == 88


def test_argument_validation_without_gpu(lib):
    n = C.c_size_t()
    good = _lib.Dims(B=64, L=10, S=3, NI=100, NU=50, NC=5, B_global=64, reserved=0)
    assert lib.tlsan_workspace_bytes(C.byref(good), C.byref(n)) == 0 and n.value > 64 * 15 * 256
    cnt = C.c_int64()
    assert lib.tlsan_flat_count(C.byref(good), C.byref(cnt)) == 0
    assert cnt.value == 105 * 64 + 100 + 50 * 44 + 4456
    for field, val, msg in (("B", 0, b"B must"), ("L", 97, b"L must"), ("S", 0, b"S must"), ("NI", 0, b"table")):
        bad = _lib.Dims(B=64, L=10, S=3, NI=100, NU=50, NC=5, B_global=64, reserved=0)
        setattr(bad, field, val)
        assert lib.tlsan_workspace_bytes(C.byref(bad), C.byref(n)) == -1
        assert msg in lib.tlsan_last_error()
    p, b = _lib.Params(), _lib.Batch()
    assert lib.tlsan_score(C.byref(good), C.byref(p), C.byref(b), 1, None, None, None) == -3      # NULL tables
    assert lib.tlsan_score(C.byref(good), None, C.byref(b), 1, None, None, None) == -3
    assert lib.tlsan_train_step(C.byref(good), C.byref(p), C.byref(b), 1.0, 0.0, 5.0, None, 0, None, None) == -3
    p.emb = 8; p.usert = 16; p.item_b = 16; p.dense = 16; p.icl = 16                             # misaligned emb
    assert lib.tlsan_score(C.byref(good), C.byref(p), C.byref(b), 1, None, None, None) == -2
    with pytest.raises(_lib.TlsanError):
        _lib.check(-2)


def test_stage_words_cover_the_packed_layout_and_the_ragged_tail(lib):
    """tlsan_stage_words = packed batch layout (model._pack_offsets) + session offsets + worst-case session items."""
    from tlsan_b200.model import _pack_offsets
    up4 = lambda n: (n + 3) // 4 * 4
    for B, L, S in ((1, 1, 1), (7, 10, 3), (65536, 10, 18), (333, 90, 41)):
        d = _lib.Dims(B=B, L=L, S=S, NI=100, NU=50, NC=5, B_global=B, reserved=0)
        w = C.c_int64()
        assert lib.tlsan_stage_words(C.byref(d), C.byref(w)) == 0
        _, total = _pack_offsets(B, L, S)
        assert w.value == total + up4(B) + up4(B * S)


def test_model_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tlsan_b200.model import Model
    with pytest.raises(_lib.TlsanError):
        Model({"item_count": 10, "user_count": 10, "cate_count": 3, "Ls": 10}, [0] * 10)


def test_product_package_never_imports_the_oracle():
    import ast
    for root, _, files in os.walk(os.path.join(ROOT, "tlsan_b200")):
        for f in files:
            if f.endswith(".py"):
                tree = ast.parse(open(os.path.join(root, f)).read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        names = [node.module or ""]
                    assert not any(n.split(".")[0] == "oracle" for n in names), f


def test_host_packer_matches_numpy_and_validates(lib):
    import numpy as np
    from tests.util import synth_batch
    from tlsan_b200.model import _pack_offsets, pack_batch
    rng = np.random.default_rng(2)
    for B, L, S, is_test in ((5, 3, 1, False), (1000, 10, 7, True), (30000, 10, 18, False)):
        NI, NU, NC = 500, 300, 11
        batch = synth_batch(rng, B, L, S, NI, NU, NC, is_test=is_test)
        dims = _lib.Dims(B=B, L=L, S=S, NI=NI, NU=NU, NC=NC, B_global=B, reserved=0)
        offs, total = _pack_offsets(B, L, S)
        out = np.full(total, -7, np.int32)
        pack_batch(lib, batch, dims, is_test, out)
        ref = {"u": batch[0], "i": batch[1], "c": batch[8], "sl": batch[6], "sl_new": batch[7],
               "hist_i": batch[3].ravel(), "hist_i_new": batch[4].ravel()}
        for k, v in ref.items():
            assert np.array_equal(out[offs[k]:offs[k] + len(v)], np.asarray(v, np.int32)), k
        assert np.array_equal(out.view(np.float32)[offs["hist_t"]:offs["hist_t"] + B * L], batch[5].ravel())
        sec = out[offs["second"]:offs["second"] + B]
        if is_test:
            assert np.array_equal(sec, batch[2].astype(np.int32))
        else:
            assert np.array_equal(sec.view(np.float32), batch[2].astype(np.float32))
        good = out.copy()
        for field, val in ((1, NI), (3, -1), (8, NC), (6, 0), (0, NU), (4, 2 ** 40)):
            bad = [np.array(f) for f in batch]
            bad[field].reshape(-1)[B // 2] = val
            with pytest.raises(IndexError):
                pack_batch(lib, tuple(bad), dims, is_test, out)
        # lists (what the reference batcher returns) are accepted too
        as_lists = tuple(f.tolist() if j in (0, 1, 2, 6, 7, 8) else f for j, f in enumerate(batch))
        out2 = np.full(total, -7, np.int32)
        pack_batch(lib, as_lists, dims, is_test, out2)
        assert np.array_equal(out2, good)
