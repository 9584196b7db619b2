"""One small pass of every stage (for the ncu launch list): mel, encoder on B clips, a few decode steps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wb200  # noqa: E402

prec = wb200.WB_PREC_BF16 if len(sys.argv) > 1 and sys.argv[1] == "bf16" else wb200.WB_PREC_FP32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
new = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m = wb200.Whisper(wb200.default_cfg("base", precision=prec, max_batch=B, max_chunks=B))
x = wb200.synth.fast_batch(B, seed=1)
m.upload_pcm(x)
m.run_log_mel()
m.encode(None, 0, B, want_hidden=False)
toks = m.greedy_decode(B, [50258, 50259, 50359, 50363], new, 50257)
print("ok", toks[0])
