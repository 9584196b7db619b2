"""GPU: the drop-in CLI end to end on synthetic WAV files — flags, the three output files, their
schema (key order, rounding) and the transcripts, against the reference's committed outputs and
the library called directly."""
import csv
import json
import os
import subprocess

import numpy as np
import pytest

import host_ref as hr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "whisper-rust-ort_b200", "whisper_b200_cli")

SUMMARY_KEYS = ["breakdown_s", "config_used", "language", "latency_end_to_end_s", "max_new_tokens", "model_id", "n_files",
                "notes", "onnx_dir", "rtf_end_to_end", "task", "timestamps", "tokenizer_json"]
STAT_KEYS = ["max", "mean", "median", "min", "p90", "p95"]


def test_cli_end_to_end(wb, tmp_path):
    audio, onnx, out = tmp_path / "audio", tmp_path / "onnx", tmp_path / "out"
    audio.mkdir(); onnx.mkdir()
    clips = {"b_long.wav": wb.synth.clip(1, 3, 41.0), "a_short.WAV": wb.synth.clip(2, 3, 7.3), "c.txt": None}
    for name, x in clips.items():
        if x is None:
            (audio / name).write_text("not audio")
        else:
            wb.synth.write_wav(str(audio / name), x, fmt="f32")
    (onnx / "generation_config.json").write_text(json.dumps({"suppress_tokens": [1, 2, 7], "begin_suppress_tokens": [220, 50257]}))
    cmd = [EXE, "--audio-dir", str(audio), "--onnx-dir", str(onnx), "--max-new-tokens", "6", "--warmup", "1", "--write-txt",
           "--precision", "fp32", "--batch", "4", "--intra-op", "3", "--chunk-parallelism", "4",
           "--out-csv", str(out / "per_file.csv"), "--out-json", str(out / "per_file.json"), "--out-summary-json", str(out / "summary.json")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "DONE" and lines[1] == "Config used:" and any(l.startswith("End-to-end p95(s): ") for l in lines)

    summary = json.loads((out / "summary.json").read_text())
    assert list(summary.keys()) == SUMMARY_KEYS                               # alphabetical, like serde_json
    assert list(summary["breakdown_s"].keys()) == ["decode_s", "load_s", "model_only_s", "preprocess_s"]
    assert list(summary["latency_end_to_end_s"].keys()) == STAT_KEYS
    assert summary["n_files"] == 2 and summary["max_new_tokens"] == 6 and summary["config_used"]["intra_op"] == 3
    assert summary["tokenizer_json"] == "" and summary["notes"]["token_decode"].startswith("Prints token IDs")
    ref_p = "/root/reference/results.old/benchmarks/container_4c4g/epyc-9654/without_hf_pipeline_rust/inference_summary.json"
    if os.path.exists(ref_p):
        ref = json.loads(open(ref_p).read())
        assert list(ref.keys()) == SUMMARY_KEYS and list(ref["config_used"].keys()) == list(summary["config_used"].keys())

    rows = json.loads((out / "per_file.json").read_text())
    assert [r_["file"] for r_ in rows] == ["a_short.WAV", "b_long.wav"]             # sorted, extension case-insensitive
    assert list(rows[0].keys()) == ["file", "duration_s", "end_to_end_s", "rtf", "text"]
    assert rows[0]["duration_s"] == 7.3 and rows[1]["duration_s"] == 41.0
    with open(out / "per_file.csv", newline="") as f:
        table = list(csv.reader(f))
    assert table[0] == ["file", "duration_s", "end_to_end_s", "rtf", "text"]
    assert table[1][1] == "7.300" and table[2][0] == "b_long.wav" and table[2][4] == rows[1]["text"]

    # transcripts: no tokenizer -> "[TOKENS:...]" per chunk, stitched with a space (main.rs:644-647, 659-684)
    m = wb.Whisper(wb.default_cfg("base", max_batch=4, max_chunks=8))
    prompt = [50258, 50259, 50359, 50363]
    for row, name in zip(rows, ["a_short.WAV", "b_long.wav"]):
        toks, fidx = m.transcribe_batch([clips[name]], prompt, 6, 50257, [1, 2, 7], [220, 50257])
        texts = [hr.decode_tokens_fallback(t[4:]) for t in toks]
        assert row["text"] == hr.stitch_texts(texts)
        stem = name.rsplit(".", 1)[0]
        assert (out / f"{stem}.transcript.txt").read_text() == row["text"].strip() + "\n"
    assert len(m.transcribe_batch([clips["b_long.wav"]], prompt, 6, 50257)[0]) == 2    # 41 s -> 2 chunks
    m.close()


def test_cli_unsupported_container_aborts_like_anyhow(wb, tmp_path):
    audio, onnx = tmp_path / "audio", tmp_path / "onnx"
    audio.mkdir(); onnx.mkdir()
    (audio / "x.mp3").write_bytes(b"ID3" + b"\0" * 100)
    r = subprocess.run([EXE, "--audio-dir", str(audio), "--onnx-dir", str(onnx), "--arch", "toy", "--out-csv", str(tmp_path / "o.csv"),
                        "--out-json", str(tmp_path / "o.json"), "--out-summary-json", str(tmp_path / "s.json")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "unsupported audio container" in r.stderr


def test_cli_file_batch_gives_the_same_transcripts(wb, tmp_path):
    """--file-batch packs the chunks of several files into shared GPU batches (the log-mel clamp stays
    per file): transcripts must not depend on the grouping."""
    audio, onnx = tmp_path / "audio", tmp_path / "onnx"
    audio.mkdir(); onnx.mkdir()
    for i, sec in enumerate([3.0, 33.0, 9.5, 30.0, 61.0]):
        wb.synth.write_wav(str(audio / f"f{i}.wav"), wb.synth.clip(i, 8, sec) * (0.02 if i == 1 else 1.0), fmt="s16")
    texts = {}
    for fb in (1, 2, 5):
        out = tmp_path / f"out{fb}"
        r = subprocess.run([EXE, "--audio-dir", str(audio), "--onnx-dir", str(onnx), "--arch", "base", "--precision", "fp32",
                            "--max-new-tokens", "5", "--batch", "4", "--file-batch", str(fb),
                            "--out-csv", str(out / "a.csv"), "--out-json", str(out / "a.json"), "--out-summary-json", str(out / "s.json")],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        rows = json.loads((out / "a.json").read_text())
        texts[fb] = [(x["file"], x["text"]) for x in rows]
        assert json.loads((out / "s.json").read_text())["n_files"] == 5
    assert texts[1] == texts[2] == texts[5]
    assert all(t.startswith("[TOKENS:") for _, t in texts[1])


def test_cli_scheduler_gpus_and_in_flight_give_the_same_rows(wb, tmp_path):
    """The batch scheduler (--gpus N: one worker process per GPU, rows gathered by the parent;
    --in-flight S: S contexts per GPU) only changes WHERE a file group runs: file order, durations and
    transcripts are those of the serial loop.  On a 1-GPU box the second worker shares device 0."""
    audio, onnx = tmp_path / "audio", tmp_path / "onnx"
    audio.mkdir(); onnx.mkdir()
    secs = [3.0, 33.0, 9.5, 30.0, 12.0, 5.0, 41.0]
    for i, sec in enumerate(secs):
        wb.synth.write_wav(str(audio / f"f{i}.wav"), wb.synth.clip(i, 8, sec), fmt="s16")
    rows = {}
    for name, extra in (("serial", []), ("inflight", ["--in-flight", "3"]), ("gpus", ["--gpus", "2", "--file-batch", "2"]),
                        ("both", ["--gpus", "2", "--in-flight", "2", "--warmup", "1"])):
        out = tmp_path / name
        r = subprocess.run([EXE, "--audio-dir", str(audio), "--onnx-dir", str(onnx), "--arch", "base", "--precision", "fp32",
                            "--max-new-tokens", "5", "--batch", "4", "--write-txt", *extra,
                            "--out-csv", str(out / "a.csv"), "--out-json", str(out / "a.json"), "--out-summary-json", str(out / "s.json")],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        assert r.stdout.startswith("DONE\n")
        got = json.loads((out / "a.json").read_text())
        rows[name] = [(x["file"], x["duration_s"], x["text"]) for x in got]
        summ = json.loads((out / "s.json").read_text())
        assert summ["n_files"] == len(secs) and summ["latency_end_to_end_s"]["min"] > 0
        assert sorted(p.name for p in out.iterdir() if p.name.endswith(".transcript.txt")) == [f"f{i}.transcript.txt" for i in range(len(secs))]
        assert not [p for p in out.iterdir() if p.name.endswith(".rows")]          # worker hand-off files are removed
    assert rows["serial"] == rows["inflight"] == rows["gpus"] == rows["both"]
    assert [f for f, _, _ in rows["serial"]] == [f"f{i}.wav" for i in range(len(secs))]


def test_cli_scheduler_worker_failure_fails_the_run(wb, tmp_path):
    audio, onnx = tmp_path / "audio", tmp_path / "onnx"
    audio.mkdir(); onnx.mkdir()
    for i in range(3):
        wb.synth.write_wav(str(audio / f"f{i}.wav"), wb.synth.clip(i, 8, 4.0), fmt="s16")
    (audio / "f1.mp3").write_bytes(b"ID3 not really audio")
    out = tmp_path / "out"
    r = subprocess.run([EXE, "--audio-dir", str(audio), "--onnx-dir", str(onnx), "--arch", "base", "--gpus", "2",
                        "--out-csv", str(out / "a.csv"), "--out-json", str(out / "a.json"), "--out-summary-json", str(out / "s.json")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 1 and "Error:" in r.stderr
    assert not (out / "a.csv").exists()


def test_cli_reads_mp3_next_to_wav(wb, tmp_path):
    """The reference's extension filter takes .mp3 too (main.rs:1116) and symphonia decodes it (main.rs:266-275): an MPEG
    Layer III file goes through the same pipeline as a WAV file; its row is what the library gives for the PCM the loader
    (csrc/host/mp3.cpp, held to libavcodec in tests/test_mp3_cpu.py) produced."""
    import ctypes as C
    import mp3_writer as mw
    audio, onnx, out = tmp_path / "audio", tmp_path / "onnx", tmp_path / "out"
    audio.mkdir(); onnx.mkdir()
    frames, sr = mw.make_stream(seed=11, version=1, sr_idx=2, br_idx=8, channels=2, n_frames=60, ms=True)      # 16 kHz, 2.16 s
    (audio / "a.mp3").write_bytes(b"".join(frames))
    wb.synth.write_wav(str(audio / "b.wav"), wb.synth.clip(4, 1, 3.0), fmt="s16")
    r = subprocess.run([EXE, "--audio-dir", str(audio), "--onnx-dir", str(onnx), "--max-new-tokens", "5", "--precision", "fp32",
                        "--out-csv", str(out / "p.csv"), "--out-json", str(out / "p.json"), "--out-summary-json", str(out / "s.json")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rows = json.loads((out / "p.json").read_text())
    assert [x["file"] for x in rows] == ["a.mp3", "b.wav"]
    L = wb.lib()
    f32p = C.POINTER(C.c_float)
    L.wb_host_load_audio_16k_mono.argtypes = [C.c_char_p, C.POINTER(f32p), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.wb_host_free.argtypes = [C.c_void_p]
    buf, n, dur = f32p(), C.c_int64(), C.c_double()
    assert L.wb_host_load_audio_16k_mono(str(audio / "a.mp3").encode(), C.byref(buf), C.byref(n), C.byref(dur)) == 0
    pcm = np.ctypeslib.as_array(buf, (n.value,)).copy()
    L.wb_host_free(buf)
    assert n.value == 60 * 576 and rows[0]["duration_s"] == round(dur.value, 3)
    m = wb.Whisper(wb.default_cfg("base", max_batch=4, max_chunks=8))
    toks, _ = m.transcribe_batch([pcm], [50258, 50259, 50359, 50363], 5, 50257)
    assert rows[0]["text"] == hr.stitch_texts([hr.decode_tokens_fallback(t[4:]) for t in toks])
    m.close()
