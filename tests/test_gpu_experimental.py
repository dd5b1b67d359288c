"""Experimental kernels that are compiled into the library but are NOT on the default path (round-1 leftovers that
could not be run before the GPU budget ended).  Skipped unless TLSAN_TEST_EXPERIMENTAL=1:

    TLSAN_TEST_EXPERIMENTAL=1 TLSAN_BWD_LONG=diet   python -m pytest tests/test_gpu_experimental.py -q
    TLSAN_TEST_EXPERIMENTAL=1 TLSAN_SORT_IMPL=count python -m pytest tests/test_gpu_experimental.py -q

run the train-step parity checks with the register-diet long-term backward (csrc/tlsan_fused_diet.cu) / the
counting sort (csrc/tlsan_sort.cu).  The counting sort must additionally be BIT-identical to the radix sort:
compare `python tools/state_hash.py` with and without TLSAN_SORT_IMPL=count."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("TLSAN_TEST_EXPERIMENTAL") != "1", reason="experimental kernels are opt-in")]


@pytest.mark.parametrize("B,L,S,full", [(32, 10, 3, False), (257, 10, 4, True), (40, 90, 5, False), (64, 17, 2, True)])
def test_train_step_parity_with_selected_experimental_kernels(B, L, S, full):
    assert os.environ.get("TLSAN_BWD_LONG") == "diet" or os.environ.get("TLSAN_SORT_IMPL") == "count", \
        "select the kernel under test: TLSAN_BWD_LONG=diet and / or TLSAN_SORT_IMPL=count"
    from tests.test_gpu_parity import _cfg, _check_step, _params
    from tests.util import synth_batch
    rng = np.random.default_rng(B * 1000 + L)
    NU, NI, NC = 50, 301, 7
    cfg = _cfg(NU, NI, NC, L)
    params = _params(cfg, seed=B)
    icl = rng.integers(0, NC, NI).astype(np.int32)
    _check_step(params, icl, cfg, synth_batch(rng, B, L, S, NI, NU, NC, full=full), lr=0.5)
