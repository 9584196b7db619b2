"""GPU box: how fast the drop-in CLI ingests MP3.  Writes N copies of one generated 30 s 44.1 kHz joint-stereo Layer III
stream (tests/mp3_writer.py: no encoder exists offline), runs whisper_b200_cli in throughput mode over them with 1 and 8
decode threads per loader (--intra-op) and prints files, wall seconds and audio-s/s of each run."""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mp3_writer as mw  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
t0 = time.time()
frames, sr = mw.make_stream(seed=1, version=0, sr_idx=0, br_idx=9, channels=2, n_frames=1150, ms=True)
blob = b"".join(frames)
print("generated %.1f s of %d Hz stereo MP3 (%d bytes) in %.1f s" % (1150 * 1152 / sr, sr, len(blob), time.time() - t0), flush=True)
with tempfile.TemporaryDirectory() as d:
    audio, onnx = os.path.join(d, "audio"), os.path.join(d, "onnx")
    os.makedirs(audio); os.makedirs(onnx)
    for i in range(n):
        with open(os.path.join(audio, "c%04d.mp3" % i), "wb") as f:
            f.write(blob)
    exe = os.path.join(ROOT, "whisper-rust-ort_b200", "whisper_b200_cli")
    for threads in (1, 8):
        out = os.path.join(d, "out%d" % threads)
        cmd = [exe, "--audio-dir", audio, "--onnx-dir", onnx, "--max-new-tokens", "128", "--warmup", "1", "--precision", "bf16",
               "--file-batch", "32", "--in-flight", "2", "--intra-op", str(threads),
               "--out-csv", os.path.join(out, "p.csv"), "--out-json", os.path.join(out, "p.json"), "--out-summary-json", os.path.join(out, "s.json")]
        t0 = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, WB_CLI_TRACE="1"))
        wall = time.time() - t0
        assert r.returncode == 0, r.stderr[-2000:]
        rows = json.load(open(os.path.join(out, "p.json")))
        audio_s = sum(x["duration_s"] for x in rows)
        s = json.load(open(os.path.join(out, "s.json")))
        print(json.dumps({"decode_threads": threads, "files": len(rows), "audio_s": round(audio_s, 1), "wall_s": round(wall, 2),
                          "audio_s_per_s_whole_process": round(audio_s / wall, 1), "load_s_mean": s["breakdown_s"]["load_s"]["mean"]}), flush=True)
        trace = [l for l in r.stderr.splitlines() if "groups" in l or "steady" in l]
        for l in trace[-3:]:
            print("   ", l)
