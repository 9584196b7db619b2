"""CPU: the C-ABI library loads, exports every symbol include/whisper_b200.h declares, and fails
loudly (no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "whisper_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wb_[a-z0-9_]+)\s*\(", src)))


def test_header_compiles_as_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "whisper_b200.h"\nint main(void){ wb_model_cfg c; (void)c; return sizeof(wb_timing) > 0 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o", str(tmp_path / "t.o")])


def test_library_exports_every_declared_symbol(wb):
    L = wb.lib()
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_default_cfg_and_struct_layout(wb):
    cfg = wb.default_cfg("base")
    assert (cfg.n_mels, cfg.d_model, cfg.n_heads, cfg.ffn_dim, cfg.enc_layers, cfg.dec_layers, cfg.vocab) == (80, 512, 8, 2048, 6, 6, 51865)
    big = wb.default_cfg("large-v3")
    assert (big.n_mels, big.d_model, big.enc_layers, big.vocab) == (128, 1280, 32, 51866)
    with pytest.raises(wb.WbError, match="unknown model name"):
        wb.default_cfg("huge")
    assert wb.binding.model_cfg_of(cfg) == wb.weights.WHISPER_BASE


def test_no_gpu_means_loud_failure_not_cpu_fallback(wb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(wb.WbError, match="no CPU fallback"):
        wb.Whisper(wb.default_cfg("toy"))


def test_fft_selftest_matches_numpy(wb):
    # the kernel's 20x20 four-step FFT (csrc/mel_math.h) emulated on the host
    L = wb.lib()
    rng = np.random.default_rng(0)
    re_, im = rng.normal(size=400).astype(np.float32), rng.normal(size=400).astype(np.float32)
    ore, oim = np.empty(400, np.float32), np.empty(400, np.float32)
    fp = C.POINTER(C.c_float)
    L.wb_selftest_fft400.argtypes = [fp, fp, fp, fp]
    assert L.wb_selftest_fft400(re_.ctypes.data_as(fp), im.ctypes.data_as(fp), ore.ctypes.data_as(fp), oim.ctypes.data_as(fp)) == 0
    ref = np.fft.fft(re_.astype(np.float64) + 1j * im.astype(np.float64))
    assert np.abs(ore - ref.real).max() < 3e-5 and np.abs(oim - ref.imag).max() < 3e-5
