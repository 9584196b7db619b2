import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wb200
for B in (32, 37, 74):
    m = wb200.Whisper(wb200.default_cfg("base", precision=wb200.WB_PREC_BF16, max_batch=B, max_chunks=B))
    m.upload_pcm(wb200.synth.fast_batch(B, seed=1)); m.run_log_mel(); m.encode(None, 0, B, want_hidden=False)
    m.greedy_decode(min(B, 32), [50258, 50259, 50359, 50363], 4, 50257)
    ms, by = m.bench_kernel("cross_attn", B, 30)
    print(f"B={B} CTAs={8*B}: cross_attn {ms*1000:.1f} us  {by/ms/1e6:.0f} GB/s")
    m.close()
