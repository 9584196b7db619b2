"""CPU: the ONNX-initializer reader (csrc/host/onnx.cpp) on REAL torch.onnx exports of HF's Whisper (tests/torch_export.py:
the exporter optimum calls, toy shape) — every tensor the library asks for must come back bit-identical to the HF
state_dict, for distinct weights and for HF's default initialisation (where the exporter de-duplicates the all-zero biases
and all-one LayerNorm weights into Identity aliases)."""
import ctypes as C

import numpy as np
import pytest

pytest.importorskip("transformers")
import torch_export  # noqa: E402


@pytest.fixture(scope="module", params=[True, False], ids=["distinct-weights", "hf-default-init"])
def exported(request, tmp_path_factory):
    d = tmp_path_factory.mktemp("torch_export")
    try:
        sd = torch_export.export(str(d), randomize=request.param)
    except (ImportError, AttributeError) as e:             # exporter internals moved in another torch version
        pytest.skip(f"torch.onnx TorchScript exporter not usable here: {e}")
    return d, sd


def test_every_tensor_is_recovered_bit_exactly(wb, exported):
    d, sd = exported
    L = wb.lib()
    L.wb_onnx_read_tensor.argtypes = [C.c_char_p, C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.c_int64]
    L.wb_last_error.restype = C.c_char_p
    cfg = wb.default_cfg("toy")
    specs = wb.weights.tensor_specs(wb.weights.WHISPER_TOY)
    assert len(specs) == 89
    for name, shape, *_ in specs:
        out = np.empty(shape, np.float32)
        rc = L.wb_onnx_read_tensor(str(d).encode(), C.byref(cfg), name.encode(), out.ctypes.data_as(C.POINTER(C.c_float)), out.size)
        assert rc == 0, (name, L.wb_last_error().decode())
        assert np.array_equal(out, sd[name].reshape(shape)), name


def test_the_export_has_the_real_exporters_quirks(exported):
    d, _ = exported
    enc = (d / "encoder_model.onnx").read_bytes()
    assert b"onnx::MatMul_" in enc and b"onnx::Add_" in enc and b"embed_positions.weight" not in enc     # folded position table
    assert b"q_proj.weight" not in enc and b"q_proj.bias" in enc                                         # anonymous Linear weights
