"""GPU: the toy architecture (d=128, 2 heads, ffn 256, vocab 1031) end to end in both builds —
small odd shapes exercise the generic (non-whisper-base) kernel paths: SIMT skinny GEMMs with a
ragged vocabulary, K=128/256 mma variants, batch sizes that are not multiples of anything."""
import numpy as np
import pytest

import mel_oracle as mo
import whisper_ref as wr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle(wb):
    cfg = wb.weights.WHISPER_TOY
    return wr.WhisperRef(cfg, wb.weights.generate(cfg, 0))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("B", [1, 3])
def test_toy_end_to_end(wb, oracle, precision, B):
    prec = wb.WB_PREC_FP32 if precision == "fp32" else wb.WB_PREC_BF16
    m = wb.Whisper(wb.default_cfg("toy", precision=prec, max_batch=3, max_chunks=6))
    x = [wb.synth.clip(0, 0, 31.7), wb.synth.clip(1, 0, 0.013), wb.synth.clip(2, 0, 4.0)][:B]
    mels, n_chunks = m.log_mel(x)
    ref_chunks = np.concatenate([mo.chunk_mels(mo.log_mel(c), len(c)) for c in x])
    assert n_chunks == len(ref_chunks)
    prompt, eot, sup, bsup = [1, 2, 3, 4], 1030, [5, 9], [6]
    toks, fidx = m.transcribe_batch(x, prompt, 5, eot, sup, bsup)
    assert len(toks) == n_chunks
    ref = []
    for c0 in range(0, n_chunks, 3):
        ref += wr.transcribe_tokens(oracle, ref_chunks[c0:c0 + 3], prompt, 5, eot, sup, bsup)
    if precision == "fp32":
        assert toks == ref
    else:
        enc = m.encode(ref_chunks[:min(3, n_chunks)])
        r = oracle.encode(ref_chunks[:min(3, n_chunks)])
        assert np.linalg.norm(enc - r) / np.linalg.norm(r) <= 2e-2
        assert all(len(t) == len(q) for t, q in zip(toks, ref))
    m.close()


def test_weights_from_onnx_export_and_blob(wb, tmp_path):
    """north_star: 'Weights come from the whisper-base-with-past ONNX initializers when present'."""
    from onnx_writer import export_like_optimum
    mc = wb.weights.WHISPER_TOY
    W = wb.weights.generate(mc, 3)
    ref = wr.WhisperRef(mc, W)
    mel = np.random.default_rng(1).normal(0, 0.5, (1, 80, 3000)).astype(np.float32)
    want = ref.encode(mel)
    onnx_dir = tmp_path / "onnx"
    onnx_dir.mkdir()
    export_like_optimum(mc, W, str(onnx_dir))
    blob = tmp_path / "w.wb200"
    wb.weights.save_blob(str(blob), mc, W)
    for src in (str(onnx_dir), str(blob)):
        m = wb.Whisper(wb.default_cfg("toy", max_batch=1, max_chunks=1), weights_path=src)
        assert np.array_equal(m.tensor("model.decoder.layers.1.encoder_attn.k_proj.weight", (128, 128)),
                              W["model.decoder.layers.1.encoder_attn.k_proj.weight"])
        assert np.abs(m.encode(mel) - want).max() <= 1e-4
        toks = m.greedy_decode(1, [1, 2, 3, 4], 4, 1030)
        assert toks == ref.greedy(want, [1, 2, 3, 4], 4, 1030)
        m.close()


def test_weights_from_a_real_torch_onnx_export(wb, tmp_path):
    """The same path on a REAL torch.onnx export of HF's WhisperForConditionalGeneration (tests/torch_export.py: the exporter
    optimum calls; anonymous MatMul weights, folded encoder position table): the context created from the export directory
    encodes and greedy-decodes like the oracle run on the HF state_dict."""
    pytest.importorskip("transformers")
    import torch_export
    try:
        sd = torch_export.export(str(tmp_path / "onnx"), randomize=True, seed=5)
    except (ImportError, AttributeError) as e:
        pytest.skip(f"torch.onnx TorchScript exporter not usable here: {e}")
    mc = wb.weights.WHISPER_TOY
    W = {name: sd[name].reshape(shape) for name, shape, *_ in wb.weights.tensor_specs(mc)}
    ref = wr.WhisperRef(mc, W)
    mel = np.random.default_rng(2).normal(0, 0.5, (2, 80, 3000)).astype(np.float32)
    want = ref.encode(mel)
    m = wb.Whisper(wb.default_cfg("toy", max_batch=2, max_chunks=2), weights_path=str(tmp_path / "onnx"))
    assert np.array_equal(m.tensor("model.encoder.embed_positions.weight", (1500, 128)), W["model.encoder.embed_positions.weight"])
    assert np.abs(m.encode(mel) - want).max() <= 1e-4
    assert m.greedy_decode(2, [1, 2, 3, 4], 6, 1030) == ref.greedy(want, [1, 2, 3, 4], 6, 1030)
    # ... and, without the oracle in between, like the exported GRAPHS evaluated node by node (tests/onnx_eval.py: what
    # ort::Session::run computes at main.rs:703 / :770): encoder hidden states and the first new token's logits within 1e-4
    import onnx_eval
    hidden = onnx_eval.run(str(tmp_path / "onnx" / "encoder_model.onnx"), {"input_features": mel[:1]})["last_hidden_state"]
    got = m.encode(mel)
    assert np.abs(got[:1] - hidden).max() <= 1e-4
    ids = [1, 5, 7, 9]
    graph_logits = onnx_eval.run(str(tmp_path / "onnx" / "decoder_model.onnx"), {"input_ids": np.array([ids], np.int64), "encoder_hidden_states": hidden,
                                                                                  "position_ids": np.arange(4, dtype=np.int64)[None]})["logits"]
    toks, lg = m.greedy_decode(2, ids, 1, 1030, want_logits=True)
    assert np.abs(np.asarray(lg)[0, 0] - graph_logits[0, -1]).max() <= 1e-4 and toks[0][-1] == int(graph_logits[0, -1].argmax())
    m.close()
