"""One-off measurements of the other BASELINE.json configs (not the bench contract; see bench.py):
  configs[1] log-mel only, 1024 x 30 s clips          -> achieved HBM GB/s
  configs[2] encoder forward bf16, batch 64            -> achieved TFLOP/s
  configs[4] whisper-large-v3 shapes, batch 16, random-init (decode shortened to 16 tokens)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wb200

which = sys.argv[1] if len(sys.argv) > 1 else "mel,enc"
out = {}
if "mel" in which:
    n = 1024
    m = wb200.Whisper(wb200.default_cfg("toy", max_batch=4, max_chunks=n))
    m.upload_pcm(wb200.synth.fast_batch(n, seed=1))
    best = 1e9
    for _ in range(5):
        m.run_log_mel(); best = min(best, m.timing()["mel_ms"])
    k_ms, k_bytes = m.bench_kernel("logmel", 1, 10)
    out["configs[1] log-mel 1024 x 30 s"] = {"K1a+K1b_ms": best, "K1a_ms": k_ms, "K1a_GBps_algorithmic": k_bytes / k_ms / 1e6,
                                            "audio_s_per_s": n * 30 / (best * 1e-3), "frac_of_hbm_6553.6": k_bytes / k_ms / 1e6 / 6553.6}
    m.close()
if "enc" in which:
    B = 64
    m = wb200.Whisper(wb200.default_cfg("base", precision=wb200.WB_PREC_BF16, max_batch=B, max_chunks=B))
    m.upload_pcm(wb200.synth.fast_batch(B, seed=1)); m.run_log_mel()
    best = 1e9
    for _ in range(5):
        m.encode(None, 0, B, want_hidden=False); best = min(best, m.timing()["encoder_ms"])
    out["configs[2] encoder bf16 B=64"] = {"encoder_ms": best, "TFLOPs": 87.368e9 * B / (best * 1e-3) / 1e12,
                                          "frac_of_bf16_sustained_1359.7": 87.368e9 * B / (best * 1e-3) / 1e12 / 1359.7,
                                          "audio_s_per_s": B * 30 / (best * 1e-3)}
    m.close()
if "large" in which:
    B = 16
    t0 = time.time()
    m = wb200.Whisper(wb200.default_cfg("large-v3", precision=wb200.WB_PREC_BF16, max_batch=B, max_chunks=B))
    t_create = time.time() - t0
    mel = np.random.default_rng(0).normal(0, 0.5, (B, 128, 3000)).astype(np.float32)
    m.encode(mel, want_hidden=False); m.encode(mel, want_hidden=False)
    enc_ms = m.timing()["encoder_ms"]; ckv_ms = m.timing()["cross_kv_ms"]
    toks = m.greedy_decode(B, [50258, 50259, 50360, 50364], 16, 50257)
    toks = m.greedy_decode(B, [50258, 50259, 50360, 50364], 16, 50257)
    dec_ms = m.timing()["decode_ms"]
    out["configs[4] large-v3 shapes B=16"] = {"create_s": t_create, "encoder_ms": enc_ms, "encoder_TFLOPs": 2273.8e9 * B / (enc_ms * 1e-3) / 1e12,
                                             "cross_kv_ms": ckv_ms, "decode_ms_16_tokens": dec_ms, "ms_per_step": dec_ms / 19}
    m.close()
print(json.dumps(out, indent=1))
