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
