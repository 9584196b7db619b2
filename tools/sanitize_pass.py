"""Smallest end-to-end pass for compute-sanitizer: toy + base shapes, both builds, ragged mel input."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wb200
for name, prec in (("toy", wb200.WB_PREC_FP32), ("toy", wb200.WB_PREC_BF16), ("base", wb200.WB_PREC_BF16)):
    m = wb200.Whisper(wb200.default_cfg(name, precision=prec, max_batch=3, max_chunks=6))
    x = [wb200.synth.clip(0, 0, 31.7), wb200.synth.clip(1, 0, 0.013), wb200.synth.clip(2, 0, 4.0)]
    mels, n = m.log_mel(x)
    vocab = m.cfg.vocab
    prompt = [1, 2, 3, 4] if vocab < 50000 else [50258, 50259, 50359, 50363]
    toks, fidx = m.transcribe_batch(x, prompt, 3, vocab - 1, [5], [6])
    print(name, prec, n, [len(t) for t in toks])
    m.close()
