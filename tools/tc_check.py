"""Bring-up check of the tcgen05 GEMM and the bf16 build (prints measured errors)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import wb200
m = wb200.Whisper(wb200.default_cfg("base", precision=wb200.WB_PREC_BF16, max_batch=4, max_chunks=8))
for (M, N, K, lda, batch, f32) in [(128, 128, 64, None, 1, False), (256, 128, 128, None, 1, True), (300, 256, 512, None, 1, False),
                                   (3000, 512, 240, 80, 2, False), (1500, 512, 1536, 1024, 2, True), (6000, 1536, 512, None, 1, False),
                                   (6000, 2048, 512, None, 1, False), (6000, 512, 2048, None, 1, True)]:
    d, a = m.selftest_gemm(M, N, K, lda, batch, f32)
    print(f"gemm M={M} N={N} K={K} lda={lda} batch={batch} f32={f32}: max|diff|={d:.3e} max|val|={a:.3f} rel={d/max(a,1e-9):.2e}", flush=True)
import mel_oracle as mo, whisper_ref as wr
x = wb200.synth.batch(2, seed=0)
mel = np.stack([mo.log_mel(c) for c in x])
t = time.time(); enc = m.encode(mel); print("encode bf16", time.time() - t, m.timing()["encoder_ms"])
cfg = wb200.weights.WHISPER_BASE
o = wr.WhisperRef(cfg, wb200.weights.generate(cfg, 0))
ref = o.encode(mel)
print("enc bf16 vs oracle: max abs", np.abs(enc - ref).max(), "ref max", np.abs(ref).max(), "rel fro", np.linalg.norm(enc - ref) / np.linalg.norm(ref))
g = np.load("tests/golden/hf_whisper_base_seed0.npz")
steps = 24
rt, rl = o.greedy(ref, g["prompt"], steps, 50257, g["suppress"], g["begin_suppress"], return_logits=True)
forced = np.array([s[4:] for s in rt])
toks, lg = m.greedy_decode(2, g["prompt"], steps, 50257, g["suppress"], g["begin_suppress"], forced=forced, want_logits=True)
rl = np.stack(rl, 1)
print("teacher-forced logits bf16 vs oracle: max abs", np.abs(lg - rl).max(), "logit std", rl.std())
agree = sum(int(a == b) for s, r in zip(toks, rt) for a, b in zip(s[4:], r[4:]))
print("argmax agreement", agree, "of", 2 * steps)
