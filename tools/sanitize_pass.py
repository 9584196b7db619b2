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
# decode chains wider than 32 sequences (NT = 8 GEMM passes, groups of 32 for fc2, tcgen05 vocabulary projection with N = 64)
m = wb200.Whisper(wb200.default_cfg("base", precision=wb200.WB_PREC_BF16, max_batch=40, max_chunks=40))
mel = np.random.default_rng(3).normal(0.0, 0.6, (40, 80, 3000)).astype(np.float32)
m.encode(mel, want_hidden=False)
toks = m.greedy_decode(40, [50258, 50259, 50359, 50363], 3, 50257)
print("wide", [len(t) for t in toks][:4])
m.close()
# the pool: two slots, three batches from one thread
pool = wb200.Pool(wb200.default_cfg("toy", max_batch=2, max_chunks=2), 2)
tk = [pool.submit([wb200.synth.clip(i, 0, 2.0), wb200.synth.clip(i + 9, 0, 1.0)], [1, 2, 3, 4], 3, 1030) for i in range(3)]
print("pool", [len(pool.wait(t)[0]) for t in tk])
pool.close()
