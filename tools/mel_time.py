import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wb200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = wb200.Whisper(wb200.default_cfg("toy", max_batch=4, max_chunks=n))
x = wb200.synth.fast_batch(n, seed=1)
m.upload_pcm(x)
for _ in range(3):
    m.run_log_mel(); t = m.timing()["mel_ms"]
ms, by = m.bench_kernel("logmel", 1, 10)
print(f"{n} clips: K1a+K1b {t:.3f} ms ; K1a alone {ms:.3f} ms = {by/ms/1e6:.0f} GB/s algorithmic ({by/1e6:.0f} MB)")
