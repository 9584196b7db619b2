"""GPU: the optional decode kernels (switched on per context through the environment at wb_create):
  WB_DEC_CLUSTER=1  all decoder layers of a step in one cluster-chained launch (dec_layers_kernel, dec_cluster.cu)
  WB_DEC_VOCAB=1    final LayerNorm + vocabulary projection + masked arg-max + token bookkeeping in one launch
  WB_XATTN_TC=1     stand-alone cross-attention on the TMA ring + tensor cores
  WB_SKINNY_TC=1    per-layer decode GEMMs on tcgen05 (skinny_tc_kernel, vocab_tc.cu)
  WB_DEC_LEAN=2     register-capped mma.sync decode GEMMs
Process-wide switches (read once per process) are covered by tests/test_gpu_switches.py in subprocesses.
Each is held to the same bar as the default bf16 path: teacher-forced logits within 5e-2 of the fp32 oracle
(/root/reference/src/main.rs:753-829 restated in oracle/whisper_ref.py), arg-max identical wherever the oracle's margin is
clear of bf16 noise, results independent of the batch position, eos / suppress bookkeeping identical to the default path."""
import os

import numpy as np
import pytest

import mel_oracle as mo
import whisper_ref as wr

pytestmark = pytest.mark.gpu
EOT = 50257
VARIANTS = [{"WB_DEC_CLUSTER": "1", "WB_DEC_VOCAB": "1"}, {"WB_DEC_VOCAB": "1"}, {"WB_XATTN_TC": "1"}, {"WB_DEC_CLUSTER": "1"},
            {"WB_SKINNY_TC": "1"},          # vocab_tc.cu: the per-layer decode GEMMs on tcgen05 (measured slower, kept as a record)
            {"WB_DEC_LEAN": "2"}]           # register-capped decode GEMM kernels
SWITCHES = ("WB_DEC_CLUSTER", "WB_DEC_VOCAB", "WB_XATTN_TC", "WB_SKINNY_TC", "WB_DEC_LEAN")


def make(wb, env, batch):
    old = {k: os.environ.get(k) for k in SWITCHES}
    for k in old:
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        return wb.Whisper(wb.default_cfg("base", precision=wb.WB_PREC_BF16, max_batch=batch, max_chunks=batch))
    finally:
        for k, v in old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


@pytest.fixture(scope="module")
def oracle_run(wb, golden_dir):
    g = np.load(f"{golden_dir}/hf_whisper_base_seed0.npz")
    cfg = wb.weights.WHISPER_BASE
    ref = wr.WhisperRef(cfg, wb.weights.generate(cfg, 0))
    uniq = wb.synth.batch(3, seed=11)
    mel = np.stack([mo.log_mel(c) for c in uniq])
    steps = 20
    ref_t, ref_l = ref.greedy(ref.encode(mel), g["prompt"], steps, EOT, g["suppress"], g["begin_suppress"], return_logits=True)
    return g, mel, np.array([s[len(g["prompt"]):] for s in ref_t]), np.stack(ref_l, 1)


@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: "+".join(sorted(e)))
def test_optional_kernels_match_oracle_and_are_batch_position_independent(wb, oracle_run, env):
    g, mel, forced, ref_l = oracle_run
    idx = np.array([0, 1, 2, 2, 1, 0, 0, 2, 1, 1, 0])           # 11 sequences: clusters of unequal size, every clip in several positions
    m = make(wb, env, len(idx))
    m.encode(mel[idx])
    toks, lg = m.greedy_decode(len(idx), g["prompt"], forced.shape[1], EOT, g["suppress"], g["begin_suppress"], forced=forced[idx], want_logits=True)
    got = np.array([s[len(g["prompt"]):] for s in toks])
    sup = np.isin(np.arange(ref_l.shape[-1]), g["suppress"])
    top2 = np.sort(np.where(sup, -np.inf, ref_l), -1)[..., -2:]
    clear = (top2[..., 1] - top2[..., 0]) > 0.1
    first = {}
    for r, k in enumerate(idx):
        assert np.abs(lg[r] - ref_l[k]).max() <= 5e-2, (r, float(np.abs(lg[r] - ref_l[k]).max()))
        assert np.all(got[r][clear[k]] == forced[k][clear[k]])
        if k in first:
            assert np.array_equal(lg[r], lg[first[k]])                  # same clip, other batch position: same bits
        first.setdefault(k, r)
    m.close()


@pytest.mark.parametrize("env", VARIANTS[:2], ids=lambda e: "+".join(sorted(e)))
def test_optional_kernels_free_running_decode_and_eot_bookkeeping(wb, env):
    """Free-running graph + PDL path (no logits), including a forced early end: the token every sequence emits first is
    declared the end-of-text id, so each sequence must stop after exactly one generated token (main.rs:781-783)."""
    pcm = wb.synth.batch(5, seed=3)
    prompt = [50258, 50259, 50359, 50363]
    base = make(wb, {}, 5)
    opt = make(wb, env, 5)
    outs = []
    for m in (base, opt):
        m.upload_pcm(pcm); m.run_log_mel(); m.encode(None, 0, 5, want_hidden=False)
        outs.append(m.greedy_decode(5, prompt, 40, EOT))
    a, b = np.array(outs[0]), np.array(outs[1])
    assert a.shape == b.shape == (5, 44)
    assert (a == b).mean() > 0.8                                       # different summation order: near-ties may flip and a row then diverges
    first_tok = int(b[0][4])
    short = opt.greedy_decode(5, prompt, 40, first_tok)
    assert short[0] == prompt + [first_tok]
    for row, full in zip(short, b.tolist()):
        n = len(row)
        assert row == full[:n] and (row[-1] == first_tok or n == 44)
    base.close(); opt.close()


def test_tcgen05_vocabulary_projection_argmax(wb, oracle_run):
    """vocab_tc.cu (default whenever a decode does not ask for logits): final LayerNorm + vocabulary projection on
    tcgen05.mma (swap-AB) + masked arg-max partials.  Teacher-forced on the oracle's ids WITHOUT logits, so this kernel
    (not the mma.sync one that also writes logits) produces the reported arg-max of every step: identical to the oracle
    wherever its top-1 margin is clear of bf16 noise (begin-suppress at step 0, suppress everywhere, vocabulary tail of
    the last 128-row tile excluded), identical between batch positions, and equal to the mma.sync kernel's on the same
    margin-gated positions.  A batch of 32 (all TMEM columns) and a ragged one (7)."""
    g, mel, forced, ref_l = oracle_run
    sup = np.isin(np.arange(ref_l.shape[-1]), g["suppress"])
    masked = np.where(sup, -np.inf, ref_l)
    masked[:, 0, g["begin_suppress"]] = -np.inf
    top2 = np.sort(masked, -1)[..., -2:]
    clear = (top2[..., 1] - top2[..., 0]) > 0.1
    want = masked.argmax(-1)
    assert clear.mean() > 0.3
    for n in (32, 7):
        idx = np.arange(n) % 3
        res = {}
        for flag in ("1", "0"):
            old = os.environ.get("WB_VOCAB_TC")
            os.environ["WB_VOCAB_TC"] = flag
            try:
                m = make(wb, {}, n)
            finally:
                os.environ.pop("WB_VOCAB_TC", None)
                if old is not None:
                    os.environ["WB_VOCAB_TC"] = old
            m.encode(mel[idx])
            toks = m.greedy_decode(n, g["prompt"], forced.shape[1], EOT, g["suppress"], g["begin_suppress"], forced=forced[idx])
            res[flag] = np.array([s[len(g["prompt"]):] for s in toks])
            m.close()
        for r, k in enumerate(idx):
            assert np.all(res["1"][r][clear[k]] == want[k][clear[k]]), r
            assert np.all(res["0"][r][clear[k]] == want[k][clear[k]]), r
            assert np.array_equal(res["1"][r], res["1"][k])                 # same clip, other batch position: same ids
