"""Throughput of S independent contexts (one host thread each) decoding batches of 32 concurrently on one GPU."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wb200
S = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
B = 32
prompt = [50258, 50259, 50359, 50363]
ms = [wb200.Whisper(wb200.default_cfg("base", precision=wb200.WB_PREC_BF16, max_batch=B, max_chunks=B)) for _ in range(S)]
x = wb200.synth.fast_batch(B, seed=1)
for m in ms:
    m.upload_pcm(x)
    for _ in range(2):
        m.transcribe_resident(B, prompt, 128, 50257)
def work(m):
    for _ in range(steps):
        m.transcribe_resident(B, prompt, 128, 50257)
t0 = time.perf_counter()
th = [threading.Thread(target=work, args=(m,)) for m in ms]
[t.start() for t in th]; [t.join() for t in th]
dt = time.perf_counter() - t0
print(f"S={S}: {S*steps} batches of {B} in {dt*1000:.1f} ms -> {S*steps*B*30/dt:.0f} audio-s/s, {dt/steps*1000:.1f} ms per batch latency")
