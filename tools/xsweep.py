"""Time the cross-attention kernels over Tk (WB_BENCH_TK) to separate fixed cost from the streaming rate."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "one":
    sys.path.insert(0, ROOT)
    import wb200
    prec = wb200.WB_PREC_BF16 if sys.argv[2] == "bf16" else wb200.WB_PREC_FP32
    B = 32
    m = wb200.Whisper(wb200.default_cfg("base", precision=prec, max_batch=B, max_chunks=B))
    m.upload_pcm(wb200.synth.fast_batch(B, seed=1)); m.run_log_mel(); m.encode(None, 0, B, want_hidden=False)
    m.greedy_decode(B, [50258, 50259, 50359, 50363], 4, 50257)
    for k in sys.argv[3].split(","):
        ms, by = m.bench_kernel(k, B, 300)
        print(sys.argv[2], k, "Tk", os.environ.get("WB_BENCH_TK", "1500"), "pdl", "WB_BENCH_PDL" in os.environ, f"{ms*1000:.2f} us", flush=True)
    sys.exit(0)
for prec, ks in (("bf16", "cross_attn"), ("fp32", "cross_attn")):
    for pdl in (0, 1):
        for tk in (256, 512, 768, 1024, 1280, 1500):
            env = dict(os.environ, WB_BENCH_TK=str(tk))
            if pdl: env["WB_BENCH_PDL"] = "1"
            subprocess.run([sys.executable, __file__, "one", prec, ks], env=env, timeout=120)
