import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import wb200
m = wb200.Whisper(wb200.default_cfg("base", precision=wb200.WB_PREC_BF16, max_batch=4, max_chunks=8))
for B in (1, 3):
    d, a = m.selftest_attn(B)
    print(f"attn B={B}: max|diff|={d:.4e} max|val|={a:.4f}", flush=True)
import mel_oracle as mo, whisper_ref as wr
x = wb200.synth.batch(2, seed=0)
mel = np.stack([mo.log_mel(c) for c in x])
enc = m.encode(mel); print("encoder_ms (B=2)", m.timing()["encoder_ms"])
cfg = wb200.weights.WHISPER_BASE
o = wr.WhisperRef(cfg, wb200.weights.generate(cfg, 0))
ref = o.encode(mel)
print("enc bf16 vs oracle: max abs", np.abs(enc - ref).max(), "rel fro", np.linalg.norm(enc - ref) / np.linalg.norm(ref))
