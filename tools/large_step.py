"""configs[4] widths (whisper-large-v3, batch 16, bf16): encode once, decode a few tokens. Meant to run under
ncu with a kernel-name filter to list the decode kernels of one step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wb200
B = 16
m = wb200.Whisper(wb200.default_cfg("large-v3", precision=wb200.WB_PREC_BF16, max_batch=B, max_chunks=B))
mel = np.random.default_rng(0).normal(0, 0.5, (B, 128, 3000)).astype(np.float32)
m.encode(mel, want_hidden=False)
m.greedy_decode(B, [50258, 50259, 50360, 50364], int(sys.argv[1]) if len(sys.argv) > 1 else 2, 50257)
print("decode_ms", m.timing()["decode_ms"])
