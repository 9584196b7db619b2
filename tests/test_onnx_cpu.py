"""CPU: the ONNX-initializer reader (csrc/host/onnx.cpp) on synthetic optimum-style exports —
protobuf wire parsing (packed/unpacked dims, raw_data/float_data, f16), name recovery of the
anonymous transposed MatMul weights by graph order, shape checks and error paths."""
import ctypes as C

import numpy as np
import pytest

from onnx_writer import export_like_optimum


def read(wb, d, cfg, name, shape):
    L = wb.lib()
    L.wb_onnx_read_tensor.argtypes = [C.c_char_p, C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.c_int64]
    out = np.empty(shape, np.float32)
    rc = L.wb_onnx_read_tensor(str(d).encode(), C.byref(cfg), name.encode(), out.ctypes.data_as(C.POINTER(C.c_float)), out.size)
    if rc != 0:
        raise RuntimeError(L.wb_last_error().decode())
    return out


@pytest.mark.parametrize("f16,prefix_enc,prefix_dec", [(False, "", "model.decoder."), (False, "model.encoder.", "decoder."), (True, "encoder.", "")])
def test_reader_recovers_every_tensor(wb, tmp_path, f16, prefix_enc, prefix_dec):
    mc = wb.weights.WHISPER_TOY
    W = wb.weights.generate(mc, 3)
    export_like_optimum(mc, W, str(tmp_path), prefix_enc, prefix_dec, f16)
    cfg = wb.default_cfg("toy")
    for name, shape, *_ in wb.weights.tensor_specs(mc):
        got = read(wb, tmp_path, cfg, name, shape)
        ref = W[name].astype(np.float16).astype(np.float32) if f16 else W[name]
        assert np.array_equal(got, ref), name


def test_reader_error_paths(wb, tmp_path):
    mc = wb.weights.WHISPER_TOY
    W = wb.weights.generate(mc, 3)
    cfg = wb.default_cfg("toy")
    with pytest.raises(RuntimeError, match="Failed to load"):
        read(wb, tmp_path, cfg, "model.encoder.conv1.bias", (128,))
    export_like_optimum(mc, W, str(tmp_path))
    wrong = wb.default_cfg("toy")
    wrong.ffn_dim = 512
    with pytest.raises(RuntimeError, match="unexpected shape|graph order"):
        read(wb, tmp_path, wrong, "model.encoder.conv1.bias", (128,))
    (tmp_path / "decoder_model.onnx").write_bytes(b"\x08\x08not a graph")
    with pytest.raises(RuntimeError, match="ONNX"):
        read(wb, tmp_path, cfg, "model.encoder.conv1.bias", (128,))


def _swap(keys, a, b):
    """every '<layer>.<a>' trades places with '<layer>.<b>' in the node order"""
    out = list(keys)
    for i, k in enumerate(keys):
        if k.endswith(a):
            j = keys.index(k[: -len(a)] + b)
            out[i], out[j] = keys[j], keys[i]
    return out


def test_matmul_weights_are_identified_by_their_bias_not_by_position(wb, tmp_path):
    """A different exporter / topological order: out_proj's MatMul is emitted after fc1's.  Matching by position would
    hand fc1's weight to out_proj (here caught by the shapes; with two [d, d] weights swapped it would load SILENTLY
    WRONG).  The reader follows MatMul -> Add -> named bias instead, so this export loads correctly."""
    mc = wb.weights.WHISPER_TOY
    W = wb.weights.generate(mc, 3)
    cfg = wb.default_cfg("toy")
    export_like_optimum(mc, W, str(tmp_path), reorder=lambda ks: _swap(ks, "self_attn.out_proj", "fc1"))
    for name, shape, *_ in wb.weights.tensor_specs(mc):
        assert np.array_equal(read(wb, tmp_path, cfg, name, shape), W[name]), name


def test_bias_less_k_proj_out_of_place_is_an_error_not_a_guess(wb, tmp_path):
    """k_proj has no bias to identify it: it must be the only unidentified MatMul between q_proj and v_proj.  An export
    that emits v_proj before q_proj leaves no such slot -> the load fails loudly."""
    mc = wb.weights.WHISPER_TOY
    W = wb.weights.generate(mc, 3)
    cfg = wb.default_cfg("toy")
    export_like_optimum(mc, W, str(tmp_path), reorder=lambda ks: _swap(ks, "self_attn.q_proj", "self_attn.v_proj"))
    with pytest.raises(RuntimeError, match="cannot place the weight of .*k_proj"):
        read(wb, tmp_path, cfg, "model.encoder.conv1.bias", (128,))
