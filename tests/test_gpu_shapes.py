"""GPU: decode/encode over a grid of batch sizes, prompt lengths and token budgets on the toy
architecture in both builds — every combination must match the oracle's control flow (fp32: identical
ids) and terminate.  Guards the warp-uniformity and graph-cache paths that only show up off the
B=32 / 128-token benchmark shape."""
import numpy as np
import pytest

import whisper_ref as wr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle(wb):
    cfg = wb.weights.WHISPER_TOY
    return wr.WhisperRef(cfg, wb.weights.generate(cfg, 0))


@pytest.fixture(scope="module")
def mel():
    return np.random.default_rng(11).normal(0.0, 0.6, (7, 80, 3000)).astype(np.float32)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_decode_shape_grid(wb, oracle, mel, precision):
    prec = wb.WB_PREC_FP32 if precision == "fp32" else wb.WB_PREC_BF16
    m = wb.Whisper(wb.default_cfg("toy", precision=prec, max_batch=7, max_chunks=7))
    ref_enc = oracle.encode(mel)
    for B in (1, 2, 5, 7):
        enc = m.encode(mel[:B])
        if precision == "fp32":
            assert np.abs(enc - ref_enc[:B]).max() <= 1e-4
        for prompt in ([1, 2, 3], [1, 2, 3, 4], [9]):
            for max_new in (1, 2, 17, 33):
                got = m.greedy_decode(B, prompt, max_new, 1030, [5], [6, 7])
                assert [len(s) for s in got] == [len(prompt) + max_new] * B
                if precision == "fp32":
                    assert got == oracle.greedy(ref_enc[:B], prompt, max_new, 1030, [5], [6, 7])
    # the longest decode the text context allows (448 positions)
    long = m.greedy_decode(2, [1, 2, 3, 4], 444, 1030)
    assert [len(s) for s in long] == [448, 448]
    with pytest.raises(wb.WbError, match="n_text_ctx"):
        m.greedy_decode(2, [1, 2, 3, 4], 445, 1030)
    with pytest.raises(wb.WbError, match="exceeds max_batch"):
        m.encode(np.zeros((8, 80, 3000), np.float32))
    m.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cross_attention_cta_shapes_agree(wb, precision):
    """8 heads x 40 sequences = 320 (b,h) pairs overflow one wave of 8-warp CTAs on 148 SMs, so the
    cross-attention runs its 4-warp shape; the first 20 sequences decoded alone take the 8-warp shape.
    Same keys, different merge order: fp32 logits agree to rounding.  (bf16: batches above 32 also leave the
    tensor-core decode GEMMs for the SIMT ones, which keep fp32 activations, so only the bf16 tolerance holds.)"""
    prec = wb.WB_PREC_FP32 if precision == "fp32" else wb.WB_PREC_BF16
    B = 40
    m = wb.Whisper(wb.default_cfg("base", precision=prec, max_batch=B, max_chunks=B))
    mel = np.random.default_rng(5).normal(0.0, 0.6, (B, 80, 3000)).astype(np.float32)
    m.encode(mel, want_hidden=False)
    prompt = [50258, 50259, 50359, 50363]
    forced = np.random.default_rng(6).integers(0, 50000, (B, 3))
    _, big = m.greedy_decode(B, prompt, 3, 50257, forced=forced, want_logits=True)
    _, small = m.greedy_decode(20, prompt, 3, 50257, forced=forced[:20], want_logits=True)
    assert np.isfinite(big).all()
    assert np.abs(big[:20] - small).max() <= (1e-4 if precision == "fp32" else 5e-2)
    m.close()
