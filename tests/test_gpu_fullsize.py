"""GPU: BASELINE.json's full sizes, checked through size-independent properties (the CPU oracle takes
minutes at these sizes, so it only anchors a few rows):

* configs[3] (batch 32, 128 new tokens, both builds): a clip's tokens do not depend on its position in
  the batch, duplicated clips decode identically, a batch of 32 equals its two halves of 16, and rows
  0..1 equal the oracle's greedy ids on the fp32 build;
* configs[2] (encoder, batch 64, bf16): one batch of 64 equals two batches of 32 element for element;
* configs[1] (log-mel of many 30 s clips in one launch): every copy of a clip gives the same 80x3000
  block wherever it sits, and that block is the C oracle's.
"""
import numpy as np
import pytest

import mel_oracle as mo
import whisper_ref as wr

pytestmark = pytest.mark.gpu
EOT = 50257
PROMPT = [50258, 50259, 50359, 50363]


def _decode(m, pcm, n_new):
    B = pcm.shape[0]
    m.upload_pcm(pcm)
    m.run_log_mel()
    m.encode(None, 0, B, want_hidden=False)
    return m.greedy_decode(B, PROMPT, n_new, EOT)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_configs3_batch32_128_tokens_properties(wb, precision):
    prec = wb.WB_PREC_FP32 if precision == "fp32" else wb.WB_PREC_BF16
    m = wb.Whisper(wb.default_cfg("base", precision=prec, max_batch=32, max_chunks=32))
    uniq = wb.synth.batch(12, seed=21)
    idx = np.array([i % 12 for i in range(32)])            # 12 distinct clips, each 2-3 times
    pcm = uniq[idx]
    a = _decode(m, pcm, 128)
    assert all(len(s) == 4 + 128 for s in a)
    for i in range(32):                                    # duplicates decode identically
        assert a[i] == a[idx[i]]
    perm = np.random.default_rng(0).permutation(32)        # position in the batch does not matter
    b = _decode(m, pcm[perm], 128)
    for j in range(32):
        assert b[j] == a[perm[j]]
    lo, hi = _decode(m, pcm[:16], 128), _decode(m, pcm[16:], 128)   # 32 = 16 + 16
    assert lo + hi == a
    if precision == "fp32":                                # anchor: the oracle's ids for two of the rows
        cfg = wb.weights.WHISPER_BASE
        ref = wr.WhisperRef(cfg, wb.weights.generate(cfg, 0))
        mel = np.stack([mo.log_mel(c) for c in uniq[:2]])
        assert ref.greedy(ref.encode(mel), PROMPT, 24, EOT) == [s[:4 + 24] for s in a[:2]]
    m.close()


def test_configs2_encoder_batch64_equals_two_batches_of_32(wb):
    m = wb.Whisper(wb.default_cfg("base", precision=wb.WB_PREC_BF16, max_batch=64, max_chunks=64))
    mel = np.random.default_rng(4).normal(0.0, 0.6, (64, 80, 3000)).astype(np.float32)
    full = m.encode(mel)
    assert np.isfinite(full).all()
    for _ in range(3):                                     # run-to-run identical: this caught a race on the single-buffered P tile
        assert np.array_equal(m.encode(mel), full)
    assert np.array_equal(m.encode(mel[:32]), full[:32])
    assert np.array_equal(m.encode(mel[32:]), full[32:])
    m.close()


def test_configs1_many_clips_one_launch(wb):
    n = 192
    m = wb.Whisper(wb.default_cfg("toy", max_batch=4, max_chunks=n))
    uniq = wb.synth.batch(6, seed=31)
    idx = (np.arange(n) * 7) % 6
    m.upload_pcm(uniq[idx])
    m.run_log_mel()
    first = {}
    for c in range(n):
        blk = m.chunk_mel(c, 1)[0]
        k = int(idx[c])
        if k not in first:
            first[k] = blk
            assert np.abs(blk - mo.log_mel(uniq[k])).max() <= 1e-4
        else:
            assert np.array_equal(blk, first[k])
    m.close()


def test_configs4_large_v3_full_depth_two_steps_vs_oracle(wb):
    """BASELINE.json configs[4] at its real depth (32 + 32 layers, d = 1280, 128 mel bins, vocab 51866), not the 2 + 2
    layers of tests/test_gpu_wide.py: one clip from PCM through the 128-bin log-mel, the full encoder and two greedy
    steps (main.rs:753-829), against the numpy oracle of the same seeded weights (about two minutes of host time:
    1.5 G parameters to generate, 2.3 TFLOP of fp32 encoder on the CPU).  fp32 build: encoder within 1e-3 absolute
    (activations reach ~1e1 after 32 residual layers), identical token ids, logits within 1e-3.  bf16 build: encoder
    within 2e-2 relative (north_star), teacher-forced logits within 1e-1 (32 layers of bf16 rounding, logit std ~1)."""
    mc = wb.weights.WHISPER_LARGE_V3
    oracle = wr.WhisperRef(mc, wb.weights.generate(mc, 0))
    x = wb.synth.batch(1, seed=6)
    mel = np.stack([mo.log_mel(c, n_mels=128) for c in x])
    ref_enc = oracle.encode(mel)
    prompt = [50258, 50259, 50360, 50364]
    ref_t, ref_l = oracle.greedy(ref_enc, prompt, 2, EOT, [], [], return_logits=True)
    ref_l = np.stack(ref_l, 1)
    forced = np.array([s[len(prompt):] for s in ref_t])

    m = wb.Whisper(wb.default_cfg("large-v3", precision=wb.WB_PREC_FP32, max_batch=1, max_chunks=1))
    mels, _ = m.log_mel(list(x))
    assert np.abs(mels[0] - mel[0]).max() <= 1e-4
    enc = m.encode(None, 0, 1)
    assert np.abs(enc - ref_enc).max() <= 1e-3 * max(1.0, float(np.abs(ref_enc).max()))
    toks, lg = m.greedy_decode(1, prompt, 2, EOT, want_logits=True)
    assert toks == ref_t
    assert np.abs(lg - ref_l).max() <= 1e-3 * max(1.0, float(np.abs(ref_l).max()))
    m.close()

    m = wb.Whisper(wb.default_cfg("large-v3", precision=wb.WB_PREC_BF16, max_batch=1, max_chunks=1))
    m.log_mel(list(x), want_mel=False)
    enc = m.encode(None, 0, 1)
    assert np.linalg.norm(enc - ref_enc) / np.linalg.norm(ref_enc) <= 2e-2
    toks, lg = m.greedy_decode(1, prompt, 2, EOT, forced=forced, want_logits=True)
    assert np.abs(lg - ref_l).max() <= 1e-1 * max(1.0, float(ref_l.std()))
    m.close()
