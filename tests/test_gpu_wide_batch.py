"""GPU: decode chains wider than 32 sequences (bf16 build).  One chain carries up to 64 sequences on the tensor
cores: the per-layer GEMMs take 64 sequences per pass (or walk groups of 32 with the same register-resident weights
where 64 staged rows do not fit shared memory), the tcgen05 vocabulary projection runs with N = 64.  Checked

* against the fp32 oracle (main.rs:753-829 greedy ids teacher-forced; logits within 5e-2, margin-gated arg-max),
* against its own groups of 32 decoded separately (teacher-forced logits within 2e-2; the free-running fused arg-max
  path, the one the bench runs, picks the arg-max of those logits), for a full (64) and a ragged (40, 33) batch,
* on the toy widths (K = 128 / 256 instantiations).
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


@pytest.fixture(scope="module")
def wide(wb):
    m = wb.Whisper(wb.default_cfg("base", precision=wb.WB_PREC_BF16, max_batch=64, max_chunks=64))
    yield m
    m.close()


@pytest.mark.parametrize("B", [64, 40, 33])
def test_wide_batch_agrees_with_its_groups_of_32(wb, wide, B):
    """Free-running fused path (the one the bench runs) at B > 32, then the same tokens teacher-forced through the
    wide chain and through its two groups decoded on their own.  The GEMMs compute every (row, sequence) element in the
    same order whatever the batch, but 8 heads x B > 37 sequences switch the cross-attention to its 4-warp CTA shape
    (other merge order of the key partials), so logits agree to f32 rounding of bf16-fed sums, not bit for bit."""
    uniq = wb.synth.batch(9, seed=31)
    idx = np.arange(B) % 9
    pcm = uniq[idx]
    n_new = 24
    a = _decode(wide, pcm, n_new)
    assert all(len(s) == 4 + n_new for s in a)
    for i in range(B):                                     # duplicates decode identically wherever they sit
        assert a[i] == a[idx[i]]
    forced = np.array([s[4:] for s in a])
    _, big = wide.greedy_decode(B, PROMPT, n_new, EOT, forced=forced, want_logits=True)
    # the fused tcgen05 arg-max (no logits) picked what the logits path shows, wherever the top-1 margin is above noise
    top2 = np.sort(np.partition(big, -2, axis=-1)[..., -2:], -1)
    clear = (top2[..., 1] - top2[..., 0]) > 2e-2
    assert clear.mean() > 0.2
    assert np.all(big.argmax(-1)[clear] == forced[clear])
    _decode(wide, pcm[:32], 1)
    _, lo = wide.greedy_decode(32, PROMPT, n_new, EOT, forced=forced[:32], want_logits=True)
    _decode(wide, pcm[32:], 1)
    _, hi = wide.greedy_decode(B - 32, PROMPT, n_new, EOT, forced=forced[32:], want_logits=True)
    assert np.abs(np.concatenate([lo, hi]) - big).max() <= 2e-2          # bf16 activation roundings flip on 1-ulp differences


def test_wide_batch_teacher_forced_vs_oracle(wb, wide, golden_dir):
    g = np.load(f"{golden_dir}/hf_whisper_base_seed0.npz")
    cfg = wb.weights.WHISPER_BASE
    oracle = wr.WhisperRef(cfg, wb.weights.generate(cfg, 0))
    uniq = wb.synth.batch(4, seed=5)
    idx = np.arange(64) % 4
    steps = 24
    mel = np.stack([mo.log_mel(c) for c in uniq])
    ref_t, ref_l = oracle.greedy(oracle.encode(mel), g["prompt"], steps, EOT, g["suppress"], g["begin_suppress"], return_logits=True)
    ref_l = np.stack(ref_l, 1)
    forced = np.array([s[len(g["prompt"]):] for s in ref_t])
    wide.encode(mel[idx])
    toks, lg = wide.greedy_decode(64, g["prompt"], steps, EOT, g["suppress"], g["begin_suppress"], forced=forced[idx], want_logits=True)
    got = np.array([s[len(g["prompt"]):] for s in toks])
    sup = np.isin(np.arange(ref_l.shape[-1]), g["suppress"])
    top2 = np.sort(np.where(sup, -np.inf, ref_l), -1)[..., -2:]
    clear = (top2[..., 1] - top2[..., 0]) > 0.1
    for r in range(64):
        k = idx[r]
        assert np.abs(lg[r] - ref_l[k]).max() <= 5e-2, (r, float(np.abs(lg[r] - ref_l[k]).max()))
        assert np.all(got[r][clear[k]] == forced[k][clear[k]])
        assert np.array_equal(lg[r], lg[k])
    # the free-running fused path (tcgen05 vocabulary projection, N = 64) picks the oracle's tokens where the margin is clear
    free = wide.greedy_decode(64, g["prompt"], steps, EOT, g["suppress"], g["begin_suppress"])
    first = np.array([s[len(g["prompt"])] for s in free])
    for r in range(64):
        if clear[idx[r]][0]:
            assert first[r] == forced[idx[r]][0]


def test_wide_batch_toy_widths(wb):
    m = wb.Whisper(wb.default_cfg("toy", precision=wb.WB_PREC_BF16, max_batch=48, max_chunks=48))
    mel = np.random.default_rng(12).normal(0.0, 0.6, (48, 80, 3000)).astype(np.float32)
    m.encode(mel)
    a = m.greedy_decode(48, [1, 2, 3, 4], 20, 1030, [5], [6, 7])
    m.encode(mel[:32])
    lo = m.greedy_decode(32, [1, 2, 3, 4], 20, 1030, [5], [6, 7])
    m.encode(mel[32:])
    hi = m.greedy_decode(16, [1, 2, 3, 4], 20, 1030, [5], [6, 7])
    assert lo + hi == a
    m.close()


def test_batch_above_64_falls_back_to_the_logits_path(wb):
    """Above 64 sequences the tcgen05 vocabulary kernel does not apply: the GEMMs walk two groups of 64 (fc2: three of
    32) and the arg-max runs over written logits.  Same tokens as the first 32 sequences decoded on their own wherever
    the top-1 margin is clear."""
    B = 72
    m = wb.Whisper(wb.default_cfg("base", precision=wb.WB_PREC_BF16, max_batch=B, max_chunks=B))
    mel = np.random.default_rng(21).normal(0.0, 0.6, (B, 80, 3000)).astype(np.float32)
    m.encode(mel, want_hidden=False)
    n_new = 6
    a = m.greedy_decode(B, PROMPT, n_new, EOT)
    forced = np.array([s[4:] for s in a])
    _, big = m.greedy_decode(B, PROMPT, n_new, EOT, forced=forced, want_logits=True)
    top2 = np.sort(np.partition(big, -2, axis=-1)[..., -2:], -1)
    clear = (top2[..., 1] - top2[..., 0]) > 2e-2
    assert np.all(big.argmax(-1)[clear] == forced[clear])
    m.encode(mel[:32], want_hidden=False)
    _, small = m.greedy_decode(32, PROMPT, n_new, EOT, forced=forced[:32], want_logits=True)
    assert np.abs(small - big[:32]).max() <= 2e-2
    m.close()
