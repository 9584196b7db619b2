#!/usr/bin/env bash
# GPU counterpart of the reference's core-count sweep (run_container_benchmarks.sh: `for cores in ${CORES_LIST}`):
# runs the drop-in CLI once per entry of GPUS_LIST and leaves, per entry, the same three result files the
# Rust SUT writes, under <OUT_ROOT>/gpu_<N>g/without_hf_pipeline_rust/ -- the directory layout
# compare_container_benchmarks.py reads (--results-dir <OUT_ROOT>/gpu_<N>g --log-dir <OUT_ROOT>/gpu_<N>g/logs), plus
# logs/without_hf_pipeline_rust.time.txt with the two `/usr/bin/time -v` lines that script parses (elapsed wall clock,
# maximum resident set size; written by GNU time when it is installed, else from getrusage).  One table row per entry.
#
#   GPUS_LIST="1 2 4 8" AUDIO_DIR=audio ONNX_DIR=whisper-base-with-past scripts/run_gpu_benchmarks.sh
#
# Without AUDIO_DIR content, N_SYNTH synthetic 30 s clips are generated first (tools/make_audio_dir.py).
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
CLI="${CLI:-${ROOT}/whisper-rust-ort_b200/whisper_b200_cli}"
GPUS_LIST="${GPUS_LIST:-1 2 4 8}"
AUDIO_DIR="${AUDIO_DIR:-/tmp/wb200_audio_synth}"          # not under gpurun_out/: that directory is copied back and is capped at 64 MiB
ONNX_DIR="${ONNX_DIR:-${ROOT}/gpurun_out/onnx_empty}"
OUT_ROOT="${OUT_ROOT:-${ROOT}/gpurun_out/results/benchmarks}"
N_SYNTH="${N_SYNTH:-256}"
IN_FLIGHT="${IN_FLIGHT:-4}"
FILE_BATCH="${FILE_BATCH:-32}"
MAX_NEW_TOKENS="${MAX_NEW_TOKENS:-128}"
PRECISION="${PRECISION:-bf16}"

mkdir -p "${ONNX_DIR}" "${OUT_ROOT}"
if ! ls "${AUDIO_DIR}"/*.wav >/dev/null 2>&1; then
  python "${ROOT}/tools/make_audio_dir.py" "${AUDIO_DIR}" "${N_SYNTH}" 30
fi

printf "%-5s %-7s %-10s %-9s %-12s %-10s\n" gpus files audio_s wall_s audio_s/s p95_e2e_s
for g in ${GPUS_LIST}; do
  out="${OUT_ROOT}/gpu_${g}g/without_hf_pipeline_rust"
  logs="${OUT_ROOT}/gpu_${g}g/logs"
  mkdir -p "${out}" "${logs}"
  t0=$(date +%s.%N)
  python "${ROOT}/tools/time_v.py" "${logs}/without_hf_pipeline_rust.time.txt" \
  "${CLI}" --audio-dir "${AUDIO_DIR}" --onnx-dir "${ONNX_DIR}" --language en --task transcribe \
    --max-new-tokens "${MAX_NEW_TOKENS}" --warmup 1 --write-txt --precision "${PRECISION}" \
    --gpus "${g}" --in-flight "${IN_FLIGHT}" --file-batch "${FILE_BATCH}" \
    --out-csv "${out}/inference_per_file.csv" --out-json "${out}/inference_per_file.json" \
    --out-summary-json "${out}/inference_summary.json" > "${out}/stdout.log" 2> "${out}/stderr.log"
  t1=$(date +%s.%N)
  python - "${out}" "${g}" "${t0}" "${t1}" <<'PY'
import json, sys
out, g, t0, t1 = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4])
rows = json.load(open(f"{out}/inference_per_file.json"))
summ = json.load(open(f"{out}/inference_summary.json"))
audio = sum(r["duration_s"] for r in rows)
print(f"{g:<5} {len(rows):<7} {audio:<10.1f} {t1 - t0:<9.2f} {audio / (t1 - t0):<12.1f} {summ['latency_end_to_end_s']['p95']:<10.4f}")
PY
done
