"""Scratch timing of each stage through the C ABI (not the bench contract; see bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wb200

prec = wb200.WB_PREC_BF16 if len(sys.argv) > 1 and sys.argv[1] == "bf16" else wb200.WB_PREC_FP32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 128
m = wb200.Whisper(wb200.default_cfg("base", precision=prec, max_batch=B, max_chunks=max(B, 256)))
x = wb200.synth.fast_batch(256, seed=1)
n = m.upload_pcm(x)
for _ in range(3):
    m.run_log_mel()
    print("mel 256 clips ms", m.timing()["mel_ms"])
prompt = [50258, 50259, 50359, 50363]
for it in range(3):
    t = time.time()
    m.encode(None, 0, B, want_hidden=False)
    tm = m.timing()
    toks = m.greedy_decode(B, prompt, steps, 50257)
    tm2 = m.timing()
    print(f"B={B} enc {tm['encoder_ms']:.2f} ms ckv {tm['cross_kv_ms']:.2f} ms dec {tm2['decode_ms']:.2f} ms "
          f"({tm2['decode_launches']} launches, {tm2['decode_steps']} steps) wall {time.time()-t:.3f}s "
          f"-> {B*30/(time.time()-t):.0f}x RT", flush=True)
print("tokens[0][:12]", toks[0][:12])
