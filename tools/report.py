"""Turns saved bench.py JSON lines into the table BASELINE.json's north_star asks for: throughput (audio-s/s) and p95
end-to-end latency per clip at each GPU count, as absolute numbers and as a fraction of the relevant roofline, next to
the CPU baseline measured in the same run.
    python tools/report.py profiles/r1_bench_bf16_v7.json profiles/r1_bench_bf16_v6_n8.json > profiles/r1_report.md"""
import json
import sys

# per 960 audio-seconds (one batch of 32 x 30 s, 128 new tokens) on one B200 at the MEASURED peaks:
# decode 91.34 GB over HBM (6553.6 GB/s) = 13.94 ms, encoder 32 x 87.368 GFLOP at 1359.7 TFLOP/s = 2.06 ms (+0.22 ms cross-K/V)
ROOFLINE_MS_PER_BATCH = 13.94 + 2.28
ROOFLINE_AUDIO_S_PER_S = 960.0 / (ROOFLINE_MS_PER_BATCH * 1e-3)

lines = []
for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    d["_file"] = path.split("/")[-1]
    lines.append(d)
lines.sort(key=lambda d: d["n_gpus"])

print("# whisper-base, 30 s clips, 128 new tokens, batch 32 per GPU, %s batches in flight per GPU (bf16 build)\n" %
      lines[0]["config"].get("batches_in_flight_per_gpu", "?"))
print("| GPUs | audio-s/s (PCM resident) | audio-s/s end to end (host PCM in, tokens out) | p95 latency per clip | per GPU | fraction of the "
      "per-GPU roofline (%.1f k audio-s/s) | source |" % (ROOFLINE_AUDIO_S_PER_S / 1e3))
print("|---|---|---|---|---|---|---|")
for d in lines:
    n = d["n_gpus"]
    print("| %d | %.0f | %.0f | %.3f s | %.0f | %.1f %% | `%s` |" % (
        n, d["value"], d["e2e"]["value"], d["e2e"].get("p95_latency_s_per_clip", float("nan")), d["value"] / n,
        100.0 * d["value"] / n / ROOFLINE_AUDIO_S_PER_S, d["_file"]))
one = lines[0]
if len(lines) > 1:
    print("\nWeak-scaling efficiency vs the %d-GPU line: %s." % (
        one["n_gpus"], ", ".join("%d GPUs %.2f" % (d["n_gpus"], d["value"] / d["n_gpus"] / (one["value"] / one["n_gpus"])) for d in lines[1:])))
r = one["roofline"]
print("\nDominant kernel (%s): %.0f %s of %.0f = %.0f %% (%s bound), %.1f us per launch, %.1f MB algorithmic, ncu DRAM traffic %s MB." % (
    r["kernel"], r["achieved"], r["unit"], r["peak"], 100 * r["frac"], r["bound"], r["ms_per_launch"] * 1e3, r["bytes_per_launch"] / 1e6,
    ("%.1f" % (r["traffic"] / 1e6)) if r.get("traffic") else "n/a"))
sr = one.get("stage_rooflines_single_batch")
if sr:
    print("Stages of one batch alone on the GPU: log-mel %.0f GB/s = %.1f %% of HBM; encoder %.0f TFLOP/s = %.1f %% of the sustained bf16 "
          "peak; decode %.0f GB/s = %.1f %% of HBM (%.3f ms per step)." % (
              sr["log_mel"]["achieved"], 100 * sr["log_mel"]["frac"], sr["encoder"]["achieved"], 100 * sr["encoder"]["frac"],
              sr["decode"]["achieved"], 100 * sr["decode"]["frac"], sr["decode"]["ms_per_step"]))
sb = one.get("single_batch_in_flight")
if sb:
    print("One batch in flight: %.0f audio-s/s, %.1f ms per batch (mel %.2f / encoder %.2f / cross-K/V %.2f / decode %.2f ms)." % (
        sb["value"], sb["latency_s_per_clip"] * 1e3, sb["stage_ms"]["mel_ms"], sb["stage_ms"]["encoder_ms"], sb["stage_ms"]["cross_kv_ms"],
        sb["stage_ms"]["decode_ms"]))
cb = one.get("cpu_baseline")
if cb:
    print("CPU baseline in the same run (%s, %d host cores; %s): %.1f audio-s/s -> the GPU path is %.0fx it end to end. "
          "Reference's published number (EPYC 9654, 4 cores, Rust + ORT, other hardware): 20.3x real time." % (
              cb["kind"], cb["cores"], cb["sample"], cb["value"], one["e2e"]["value"] / cb["value"]))
