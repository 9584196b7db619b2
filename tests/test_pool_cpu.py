"""CPU: the batch scheduler handle (wb_pool_*) — argument checks and the loud failure without a GPU.  No compute."""
import ctypes as C

import pytest


def test_pool_create_rejects_bad_slot_counts(wb):
    cfg = wb.default_cfg("toy")
    for n in (0, -1, 65):
        with pytest.raises(wb.WbError, match="n_slots"):
            wb.Pool(cfg, n)


def test_pool_without_gpu_fails_loudly(wb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(wb.WbError, match="no CPU fallback"):
        wb.Pool(wb.default_cfg("toy"), 2)


def test_pool_null_handle_calls_fail_instead_of_crashing(wb):
    L = wb.lib()
    assert L.wb_pool_slots(None) == 0
    n = C.c_int(0)
    assert L.wb_pool_wait(None, 0, C.byref(n)) == -1
    assert b"null pool" in L.wb_last_error()
    assert L.wb_pool_submit(None, None, None, 0, None, 0, 0, 0, None, 0, None, 0, None, None, None, 0) == -1
    L.wb_pool_destroy(None)                      # no-op
