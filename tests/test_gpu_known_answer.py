"""GPU: the known-answer run of BASELINE.json configs[0] — `audio/audio.wav` through the drop-in CLI on
the real whisper-base-with-past ONNX initializers — against the one transcript the reference's Rust
binary committed (results.old/benchmarks/container_4c4g/epyc-9654/without_hf_pipeline_rust/
audio.transcript.txt: 301.574 s of audio, --max-new-tokens 128, language en, task transcribe, see the
inference_summary.json beside it).  The transcript itself is the reference's file and is not copied into this
repo: the test holds its SHA-256, length and word count (and reads the file for a readable diff where the
reference checkout exists).

Neither the audio file nor the ONNX export exists offline, so the test is skipped until both are
pointed at:  WB_REAL_AUDIO_DIR (a directory holding audio.wav)  and  WB_REAL_ONNX_DIR (encoder_model.onnx,
decoder_model.onnx / decoder_with_past_model.onnx, tokenizer.json, generation_config.json as
/root/reference/scripts/export_onnx_whisper.py:20-28 leaves them), or dropped into assets/audio and
assets/whisper-base-with-past at the repo root.  fp32 build: greedy tokens, hence the text, must be
identical (north_star); the bf16 build is reported, not asserted."""
import hashlib
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "whisper-rust-ort_b200", "whisper_b200_cli")
AUDIO = os.environ.get("WB_REAL_AUDIO_DIR", os.path.join(ROOT, "assets", "audio"))
ONNX = os.environ.get("WB_REAL_ONNX_DIR", os.path.join(ROOT, "assets", "whisper-base-with-past"))
HAVE = os.path.isfile(os.path.join(AUDIO, "audio.wav")) and os.path.isfile(os.path.join(ONNX, "encoder_model.onnx")) \
    and os.path.isfile(os.path.join(ONNX, "tokenizer.json"))


def _run(tmp_path, precision):
    out = tmp_path / precision
    cmd = [EXE, "--audio-dir", AUDIO, "--onnx-dir", ONNX, "--language", "en", "--task", "transcribe", "--max-new-tokens", "128",
           "--warmup", "1", "--write-txt", "--precision", precision,
           "--out-csv", str(out / "per_file.csv"), "--out-json", str(out / "per_file.json"), "--out-summary-json", str(out / "summary.json")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    rows = {row["file"]: row for row in json.loads((out / "per_file.json").read_text())}
    return rows["audio.wav"], json.loads((out / "summary.json").read_text())


REF_TRANSCRIPT = "/root/reference/results.old/benchmarks/container_4c4g/epyc-9654/without_hf_pipeline_rust/audio.transcript.txt"
WANT_SHA256 = "62920c33d23be0f8e02db4bd3a45e0ea2607250a2f14bb7072eddd0cfe8aec8c"      # of the stripped UTF-8 text
WANT_CHARS, WANT_WORDS = 4637, 748


@pytest.mark.skipif(not HAVE, reason="real assets absent (set WB_REAL_AUDIO_DIR / WB_REAL_ONNX_DIR)")
def test_reference_transcript_of_audio_wav(tmp_path):
    row, summary = _run(tmp_path, "fp32")
    assert summary["notes"]["token_decode"] == "Tokenizer decode (skip_special_tokens=true)"
    assert row["duration_s"] == 301.574                                    # inference_per_file.csv of the reference run
    got = row["text"].strip()
    if os.path.exists(REF_TRANSCRIPT):
        assert got == open(REF_TRANSCRIPT, encoding="utf-8").read().strip()
    assert (len(got), len(got.split())) == (WANT_CHARS, WANT_WORDS) and got.startswith("Meet Emma, a graphic designer")
    assert hashlib.sha256(got.encode()).hexdigest() == WANT_SHA256
    row16, _ = _run(tmp_path, "bf16")
    same = hashlib.sha256(row16["text"].strip().encode()).hexdigest() == WANT_SHA256
    print("bf16 transcript identical to the reference's:", same)
