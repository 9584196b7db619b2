"""Write N synthetic 16 kHz mono s16 WAV clips (seeded tone/noise mixes, wb200.synth.clip) into a directory:
the `--audio-dir` input of whisper_b200_cli when no real audio is at hand (BASELINE.json: "synthetic 30 s clips").
    python tools/make_audio_dir.py <dir> [n_clips=64] [seconds=30]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wb200

out = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
sec = float(sys.argv[3]) if len(sys.argv) > 3 else 30.0
os.makedirs(out, exist_ok=True)
for i in range(n):
    wb200.synth.write_wav(os.path.join(out, f"clip_{i:05d}.wav"), wb200.synth.clip(i, 8, sec), fmt="s16")
print(f"{n} clips of {sec} s in {out}")
