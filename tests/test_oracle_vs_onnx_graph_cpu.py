"""CPU: oracle/whisper_ref.py against the exported ONNX GRAPH itself.  tests/torch_export.py writes a real torch.onnx export of
HF's Whisper (the exporter optimum calls for the reference's whisper-base-with-past); tests/onnx_eval.py evaluates those
graphs node by node in numpy — the arithmetic `ort::Session::run` performs at main.rs:703 (encoder) and :770 (decoder) —
and the oracle, given the same weights, must reproduce the graph's outputs: encoder hidden states and next-token logits
within 1e-4 absolute (the fp32 bar of north_star), same arg-max.  This ties the oracle to what ONNX Runtime executes, not only
to HF's eager forward (tests/golden/)."""
import numpy as np
import pytest

pytest.importorskip("transformers")
import onnx_eval  # noqa: E402
import torch_export  # noqa: E402
import whisper_ref as wr  # noqa: E402


@pytest.fixture(scope="module")
def exported(wb, tmp_path_factory):
    d = tmp_path_factory.mktemp("graph")
    try:
        sd = torch_export.export(str(d), randomize=True, seed=9)
    except (ImportError, AttributeError) as e:
        pytest.skip(f"torch.onnx TorchScript exporter not usable here: {e}")
    mc = wb.weights.WHISPER_TOY
    W = {name: sd[name].reshape(shape) for name, shape, *_ in wb.weights.tensor_specs(mc)}
    return d, wr.WhisperRef(mc, W)


def test_encoder_and_decoder_graphs(exported):
    d, oracle = exported
    rng = np.random.default_rng(4)
    for trial, ids in enumerate([[1, 5, 7, 9], [1, 1030, 2, 400], [3, 3, 3, 3]]):
        mel = rng.normal(0, 0.6, (1, 80, 3000)).astype(np.float32)
        hidden = onnx_eval.run(str(d / "encoder_model.onnx"), {"input_features": mel})["last_hidden_state"]
        want = oracle.encode(mel)
        assert hidden.shape == want.shape == (1, 1500, 128)
        assert np.abs(hidden - want).max() <= 1e-4
        logits = onnx_eval.run(str(d / "decoder_model.onnx"), {"input_ids": np.array([ids], np.int64), "encoder_hidden_states": hidden,
                                                                 "position_ids": np.arange(4, dtype=np.int64)[None]})["logits"]
        for k in range(1, 5):                                # causal mask: position k-1 of the graph sees the first k tokens only
            toks, lg = oracle.greedy(want, ids[:k], 1, 1030, [], [], return_logits=True)
            assert np.abs(logits[0, k - 1] - lg[0][0]).max() <= 1e-4
            assert int(logits[0, k - 1].argmax()) == toks[0][-1]


def test_whisper_base_shape(wb, tmp_path):
    """The same at the benchmarked architecture (d = 512, 8 heads, 6 + 6 layers, vocabulary 51865): one clip through the exported
    encoder graph, one prompt through the exported decoder graph."""
    try:
        sd = torch_export.export(str(tmp_path), randomize=True, seed=3, shape="base")
    except (ImportError, AttributeError) as e:
        pytest.skip(f"torch.onnx TorchScript exporter not usable here: {e}")
    mc = wb.weights.WHISPER_BASE
    # the product's ONNX reader on the same files (a few representative tensors; every tensor is checked at the toy shape)
    import ctypes as C
    L = wb.lib()
    L.wb_onnx_read_tensor.argtypes = [C.c_char_p, C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.c_int64]
    cfg = wb.default_cfg("base")
    shapes = {name: shape for name, shape, *_ in wb.weights.tensor_specs(mc)}
    for name in ["model.encoder.embed_positions.weight", "model.encoder.layers.3.fc1.weight", "model.decoder.embed_tokens.weight",
                 "model.decoder.layers.5.encoder_attn.k_proj.weight", "model.decoder.layers.0.self_attn.out_proj.bias"]:
        out = np.empty(shapes[name], np.float32)
        assert L.wb_onnx_read_tensor(str(tmp_path).encode(), C.byref(cfg), name.encode(), out.ctypes.data_as(C.POINTER(C.c_float)), out.size) == 0, name
        assert np.array_equal(out, sd[name].reshape(shapes[name])), name
    oracle = wr.WhisperRef(mc, {name: sd[name].reshape(shape) for name, shape, *_ in wb.weights.tensor_specs(mc)})
    mel = np.random.default_rng(8).normal(0, 0.6, (1, 80, 3000)).astype(np.float32)
    hidden = onnx_eval.run(str(tmp_path / "encoder_model.onnx"), {"input_features": mel})["last_hidden_state"]
    want = oracle.encode(mel)
    assert hidden.shape == want.shape == (1, 1500, 512) and np.abs(hidden - want).max() <= 1e-4
    ids = [50258, 50259, 50359, 50363]
    logits = onnx_eval.run(str(tmp_path / "decoder_model.onnx"), {"input_ids": np.array([ids], np.int64), "encoder_hidden_states": hidden,
                                                                    "position_ids": np.arange(4, dtype=np.int64)[None]})["logits"]
    toks, lg = oracle.greedy(want, ids, 1, 50257, [], [], return_logits=True)
    assert np.abs(logits[0, -1] - lg[0][0]).max() <= 1e-4 and int(logits[0, -1].argmax()) == toks[0][-1]
