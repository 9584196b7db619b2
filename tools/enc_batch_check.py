"""Encoder batch invariance (bf16): encode 64 clips at once vs the two halves; report where they differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wb200
m = wb200.Whisper(wb200.default_cfg("base", precision=wb200.WB_PREC_BF16, max_batch=64, max_chunks=64))
mel = np.random.default_rng(4).normal(0.0, 0.6, (64, 80, 3000)).astype(np.float32)
for rep in range(3):
    full = m.encode(mel)
    again = m.encode(mel)
    lo = m.encode(mel[:32])
    hi = m.encode(mel[32:])
    for name, a, b in (("full vs full", full, again), ("lo", lo, full[:32]), ("hi", hi, full[32:])):
        d = np.abs(a - b)
        bad = np.argwhere(d > 0)
        print(rep, name, "max", float(d.max()), "n_diff", len(bad), "first", bad[:3].tolist(), "seqs", sorted(set(bad[:, 0].tolist()))[:8],
              "rows", sorted(set(bad[:, 1].tolist()))[:6], flush=True)
