import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wb200
prec = wb200.WB_PREC_BF16 if sys.argv[1] == "bf16" else wb200.WB_PREC_FP32
B = 32
m = wb200.Whisper(wb200.default_cfg("base", precision=prec, max_batch=B, max_chunks=B))
m.upload_pcm(wb200.synth.fast_batch(B, seed=1)); m.run_log_mel(); m.encode(None, 0, B, want_hidden=False)
m.greedy_decode(B, [50258, 50259, 50359, 50363], 4, 50257)
for k in ("cross_attn", "vocab_proj"):
    ms, by = m.bench_kernel(k, B, 30)
    print(sys.argv[1], k, f"{ms*1000:.1f} us  {by/ms/1e6:.0f} GB/s")
