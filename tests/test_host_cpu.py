"""CPU: the C++ host half of the path (csrc/host/) through the C ABI against oracle/host_ref.py,
which restates the reference's Rust line by line."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import host_ref as hr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_OUT = "/root/reference/results.old/benchmarks/container_4c4g/epyc-9654/without_hf_pipeline_rust"


@pytest.fixture(scope="module")
def L(wb):
    L = wb.lib()
    f32p, i64p, f64p = C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_double)
    L.wb_host_load_audio_16k_mono.argtypes = [C.c_char_p, C.POINTER(f32p), i64p, f64p]
    L.wb_host_resample_linear.argtypes = [f32p, C.c_int64, C.c_uint32, C.c_uint32, f32p, C.c_int64]
    L.wb_host_resample_linear.restype = C.c_int64
    L.wb_host_free.argtypes = [C.c_void_p]
    L.wb_host_chunk_starts.argtypes = [C.c_int64, C.c_int64, C.c_int64, i64p, C.c_int]
    L.wb_host_word_overlap.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    L.wb_host_stitch_texts.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_int64]
    L.wb_host_stitch_texts.restype = C.c_int64
    L.wb_host_percentile.argtypes = [f64p, C.c_int, C.c_double]
    L.wb_host_percentile.restype = C.c_double
    L.wb_host_stat_block.argtypes = [f64p, C.c_int, f64p]
    L.wb_tokenizer_load.argtypes = [C.POINTER(C.c_void_p), C.c_char_p]
    L.wb_tokenizer_free.argtypes = [C.c_void_p]
    L.wb_tokenizer_token_to_id.argtypes = [C.c_void_p, C.c_char_p]
    L.wb_tokenizer_token_to_id.restype = C.c_int64
    L.wb_host_special_tokens.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, i64p]
    L.wb_host_decode_tokens.argtypes = [C.c_void_p, i64p, C.c_int, C.c_char_p, C.c_int64]
    L.wb_host_decode_tokens.restype = C.c_int64
    L.wb_cli_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    return L


def load_audio(L, path):
    p, n, dur = C.POINTER(C.c_float)(), C.c_int64(0), C.c_double(0)
    rc = L.wb_host_load_audio_16k_mono(str(path).encode(), C.byref(p), C.byref(n), C.byref(dur))
    if rc != 0:
        raise RuntimeError(L.wb_last_error().decode())
    out = np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.float32)
    L.wb_host_free(p)
    return out, dur.value


def stitch(L, chunks):
    arr = (C.c_char_p * len(chunks))(*[c.encode() for c in chunks])
    n = L.wb_host_stitch_texts(arr, len(chunks), None, 0)
    buf = C.create_string_buffer(n + 1)
    L.wb_host_stitch_texts(arr, len(chunks), buf, n + 1)
    return buf.value.decode()


# ---------------- audio ingest (main.rs:207-316) ----------------
@pytest.mark.parametrize("fmt,channels", [("s16", 1), ("s16", 2), ("u8", 1), ("f32", 1), ("f32", 3)])
def test_wav_formats_and_downmix(wb, L, tmp_path, fmt, channels):
    x = wb.synth.clip(0, seed=4, seconds=0.5)
    p = tmp_path / f"a_{fmt}_{channels}.wav"
    wb.synth.write_wav(str(p), x, fmt=fmt, channels=channels)
    got, dur = load_audio(L, p)
    raw = {"s16": np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16),
           "u8": np.clip(np.round(x * 128.0 + 128.0), 0, 255).astype(np.uint8), "f32": x}[fmt]
    ref = hr.decode_wav_samples(np.repeat(raw[:, None], channels, 1).reshape(-1), fmt, channels)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    assert dur == len(x) / 16000.0


@pytest.mark.parametrize("sr", [8000, 22050, 44100, 48000])
def test_linear_resample_to_16k(wb, L, tmp_path, sr):
    n = int(sr * 0.37)
    x = np.sin(np.arange(n) * 0.01).astype(np.float32) * 0.5
    p = tmp_path / "r.wav"
    wb.synth.write_wav(str(p), x, sr=sr, fmt="f32")
    got, dur = load_audio(L, p)
    ref = hr.resample_linear(x, sr, 16000)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    assert dur == len(ref) / 16000.0


def test_unsupported_audio_is_an_error(wb, L, tmp_path):
    # FLAC: symphonia decodes it to S32, which the reference's match rejects (main.rs:303) -- same message here
    (tmp_path / "x.flac").write_bytes(b"fLaC" + b"\0" * 64)
    with pytest.raises(RuntimeError, match="Unsupported decoded sample format"):
        load_audio(L, tmp_path / "x.flac")
    (tmp_path / "x.mp3").write_bytes(b"ID3" + b"\0" * 64)
    with pytest.raises(RuntimeError, match="unsupported audio container"):
        load_audio(L, tmp_path / "x.mp3")
    with pytest.raises(RuntimeError, match="Failed to open audio"):
        load_audio(L, tmp_path / "missing.wav")
    # 24-bit PCM decodes to S24 in symphonia, which the reference rejects (main.rs:303)
    import struct
    data = b"\0" * 300
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 16000, 48000, 3, 24) + b"data" + struct.pack("<I", len(data))
    (tmp_path / "s24.wav").write_bytes(hdr + data)
    with pytest.raises(RuntimeError, match="Unsupported decoded sample format"):
        load_audio(L, tmp_path / "s24.wav")


# ---------------- text + stats ----------------
CASES = [
    ["Hello world this is", "this is a test", "a test of stitching."],
    ["  One two  ", "", "   ", "THREE four", "three FOUR five"],
    ["no overlap here", "completely different"],
    ["same same", "same same"],
    ["Ünïcode Wörds here", "wörds HERE again and more"],
    ["a b c d e f g h i j k l m n o p q r", "b c d e f g h i j k l m n o p q r s"],
]


@pytest.mark.parametrize("chunks", CASES)
def test_stitch_texts_and_word_overlap(L, chunks):
    assert stitch(L, chunks) == hr.stitch_texts(chunks)
    for a, b in zip(chunks, chunks[1:]):
        assert L.wb_host_word_overlap(a.encode(), b.encode(), 16) == hr.word_overlap(a, b, 16)


def test_reference_transcript_is_a_fixed_point_of_stitching(L):
    # the only committed transcript of the Rust path: stitching it as a single chunk must not change it
    p = os.path.join(REF_OUT, "audio.transcript.txt")
    if not os.path.exists(p):
        pytest.skip("reference not mounted (GPU box)")
    text = open(p, encoding="utf-8").read()
    assert stitch(L, [text]) == text.strip() == hr.stitch_texts([text])


@pytest.mark.parametrize("xs", [[3.0], [1.0, 2.0], [5.0, 1.0, 4.0, 2.0, 3.0], list(np.random.default_rng(0).random(37))])
def test_percentile_and_stat_block(L, xs):
    a = (C.c_double * len(xs))(*xs)
    for p in (0.0, 50.0, 90.0, 95.0, 100.0):
        assert L.wb_host_percentile(a, len(xs), p) == hr.percentile(xs, p)
    out = (C.c_double * 6)()
    assert L.wb_host_stat_block(a, len(xs), out) == 0
    ref = hr.stat_block(xs)
    assert list(out) == [ref["min"], ref["median"], ref["p90"], ref["p95"], ref["max"], ref["mean"]]


def test_chunk_starts(L):
    for n in (1, 479999, 480000, 480001, 880000, 880001, 301574 * 16):
        buf = (C.c_int64 * 64)()
        k = L.wb_host_chunk_starts(n, 0, 0, buf, 64)
        assert list(buf[:k]) == hr.chunk_starts(n)
    assert L.wb_host_chunk_starts(int(301.574 * 16000), 0, 0, None, 0) == 12


# ---------------- tokenizer ----------------
@pytest.fixture(scope="module")
def tok_json(tmp_path_factory):
    b2u = hr.bytes_to_unicode()
    enc = lambda s: "".join(b2u[b] for b in s.encode("utf-8"))
    vocab = {enc(w): i for i, w in enumerate([" Hello", " world", "!", " caf", "é", " 日本", "\n", " a", "b"])}
    added = [{"id": 50257, "content": "<|endoftext|>", "special": True}, {"id": 50258, "content": "<|startoftranscript|>", "special": True},
             {"id": 50259, "content": "<|en|>", "special": True}, {"id": 50276, "content": "<|hi|>", "special": True},
             {"id": 50358, "content": "<|translate|>", "special": True}, {"id": 50359, "content": "<|transcribe|>", "special": True},
             {"id": 50363, "content": "<|notimestamps|>", "special": True}, {"id": 50364, "content": "<|0.00|>", "special": False}]
    p = tmp_path_factory.mktemp("tok") / "tokenizer.json"
    p.write_text(json.dumps({"version": "1.0", "added_tokens": added, "decoder": {"type": "ByteLevel"},
                             "model": {"type": "BPE", "vocab": vocab, "merges": []}}, ensure_ascii=True))
    return p


def decode(L, tok, ids):
    a = (C.c_int64 * len(ids))(*ids)
    n = L.wb_host_decode_tokens(tok, a, len(ids), None, 0)
    buf = C.create_string_buffer(n + 1)
    L.wb_host_decode_tokens(tok, a, len(ids), buf, n + 1)
    return buf.raw[:n].decode()              # the text may hold U+0000 (byte-level 'Ā'): length-delimited, not NUL-delimited


def test_tokenizer_decode_and_special_tokens(L, tok_json):
    t = C.c_void_p()
    assert L.wb_tokenizer_load(C.byref(t), str(tok_json).encode()) == 0, L.wb_last_error()
    assert L.wb_tokenizer_token_to_id(t, b"<|en|>") == 50259 and L.wb_tokenizer_token_to_id(t, b"<|xx|>") == -1
    assert decode(L, t, [50258, 50259, 0, 1, 2, 50257]) == " Hello world!"            # specials skipped
    assert decode(L, t, [3, 4, 5, 6, 50364, 99999, -5]) == " café 日本\n<|0.00|>"      # non-special added token kept, bad ids dropped
    out = (C.c_int64 * 5)()
    assert L.wb_host_special_tokens(t, b"en", b"transcribe", out) == 0 and list(out) == [50258, 50257, 50259, 50359, 50363]
    assert L.wb_host_special_tokens(t, b"zz", b"transcribe", out) != 0 and b"Tokenizer missing token: <|zz|>" in L.wb_last_error()
    L.wb_tokenizer_free(t)
    assert L.wb_tokenizer_load(C.byref(t), b"/nonexistent/tokenizer.json") != 0


def test_fallbacks_without_tokenizer(L):
    out = (C.c_int64 * 5)()
    for lang, task in (("en", "transcribe"), ("hi", "translate"), ("fr", "summarise")):
        assert L.wb_host_special_tokens(None, lang.encode(), task.encode(), out) == 0
        assert tuple(out) == hr.special_tokens(lang, task)
    ids = list(range(1000, 1300))
    assert decode(L, None, ids) == hr.decode_tokens_fallback(ids)
    assert decode(L, None, []) == "[TOKENS:]"


# ---------------- CLI surface (no GPU here: only argument handling and early errors) ----------------
def run_cli(*argv):
    exe = os.path.join(ROOT, "whisper-rust-ort_b200", "whisper_b200_cli")
    return subprocess.run([exe, *argv], capture_output=True, text=True)


def test_cli_flag_surface_matches_reference():
    r = run_cli("--help")
    assert r.returncode == 0
    for flag in ("--audio-dir", "--model-id", "--onnx-dir", "--language", "--task", "--max-new-tokens", "--warmup",
                 "--limit-files", "--discovery-best-json", "--out-csv", "--out-json", "--out-summary-json", "--intra-op",
                 "--inter-op", "--write-txt", "--tokenizer-json", "--timestamps", "--chunk-parallelism", "--chunk-length-s", "--overlap-s"):
        assert flag in r.stdout, flag
    assert run_cli("--no-such-flag").returncode == 2                       # clap usage error
    assert run_cli("--max-new-tokens", "abc").returncode == 2


def test_cli_early_errors_mirror_anyhow(tmp_path):
    out = ["--out-csv", str(tmp_path / "o/a.csv"), "--out-json", str(tmp_path / "o/a.json"), "--out-summary-json", str(tmp_path / "o/s.json")]
    r = run_cli("--onnx-dir", str(tmp_path / "nope"), *out)
    assert r.returncode == 1 and "onnx_dir does not exist or is not a directory" in r.stderr
    assert (tmp_path / "o").is_dir()                                       # parents are created first (main.rs:1069-1071)
    r = run_cli("--tokenizer-json", str(tmp_path / "tok.json"), "--onnx-dir", str(tmp_path), *out)
    assert r.returncode == 1 and "tokenizer_json not found" in r.stderr


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def test_cli_scheduler_flags_and_file_listing_errors(wb, tmp_path):
    r = run_cli("--help")
    for flag in ("--file-batch", "--gpus", "--in-flight", "--device", "--precision", "--batch", "--weights", "--arch", "--seed"):
        assert flag in r.stdout, flag
    assert run_cli("--gpus", "x").returncode == 2 and run_cli("--in-flight").returncode == 2
    out = ["--out-csv", str(tmp_path / "o/a.csv"), "--out-json", str(tmp_path / "o/a.json"), "--out-summary-json", str(tmp_path / "o/s.json")]
    onnx, audio = tmp_path / "onnx", tmp_path / "audio"
    onnx.mkdir(); audio.mkdir()
    r = run_cli("--onnx-dir", str(onnx), "--audio-dir", str(tmp_path / "missing"), *out)
    assert r.returncode == 1 and "cannot read audio dir" in r.stderr
    (audio / "notes.txt").write_text("not audio")
    r = run_cli("--onnx-dir", str(onnx), "--audio-dir", str(audio), *out)
    assert r.returncode == 1 and "No audio files found" in r.stderr             # main.rs:1126-1128


@pytest.mark.skipif(not _no_gpu(), reason="checks the loud failure of a box WITHOUT a GPU")
def test_cli_without_a_gpu_fails_loudly_in_both_scheduler_modes(wb, tmp_path):
    """No CPU fallback: the serial path dies in wb_create, the --gpus path in its worker processes (the parent then
    fails the run and removes the workers' hand-off files)."""
    onnx, audio = tmp_path / "onnx", tmp_path / "audio"
    onnx.mkdir(); audio.mkdir()
    for i in range(3):
        wb.synth.write_wav(str(audio / f"f{i}.wav"), wb.synth.clip(i, 8, 1.0), fmt="s16")
    for extra, msg in (([], "no CPU fallback"), (["--gpus", "2"], "a GPU worker process failed")):
        out_dir = tmp_path / ("o" + str(len(extra)))
        r = run_cli("--onnx-dir", str(onnx), "--audio-dir", str(audio), "--arch", "toy", *extra,
                    "--out-csv", str(out_dir / "a.csv"), "--out-json", str(out_dir / "a.json"), "--out-summary-json", str(out_dir / "s.json"))
        assert r.returncode == 1 and msg in r.stderr, r.stderr
        assert not (out_dir / "a.csv").exists()
        assert not [p for p in out_dir.iterdir() if p.name.endswith(".rows")]


def test_known_answer_digest_is_the_reference_transcripts():
    # tests/test_gpu_known_answer.py compares the drop-in's transcript of audio/audio.wav with the Rust binary's by digest
    # (the reference's file is not copied into this repo): the digest it holds must be that file's
    import hashlib
    import test_gpu_known_answer as ka
    if not os.path.exists(ka.REF_TRANSCRIPT):
        pytest.skip("reference checkout not present")
    t = open(ka.REF_TRANSCRIPT, encoding="utf-8").read().strip()
    assert (hashlib.sha256(t.encode()).hexdigest(), len(t), len(t.split())) == (ka.WANT_SHA256, ka.WANT_CHARS, ka.WANT_WORDS)


def test_architecture_from_hf_config_json(wb, tmp_path):
    # the export directory's config.json sizes the model when --arch is not given (the reference takes shapes from the graphs)
    L = wb.lib()
    L.wb_cfg_from_hf_config.argtypes = [C.c_char_p, C.c_void_p]
    transformers = pytest.importorskip("transformers")
    for name, kw in [("toy", dict(vocab_size=1031, num_mel_bins=80, d_model=128, encoder_layers=2, decoder_layers=2, encoder_attention_heads=2,
                                  decoder_attention_heads=2, encoder_ffn_dim=256, decoder_ffn_dim=256, max_source_positions=1500,
                                  max_target_positions=448)),
                     ("large-v3", dict(vocab_size=51866, num_mel_bins=128, d_model=1280, encoder_layers=32, decoder_layers=32,
                                       encoder_attention_heads=20, decoder_attention_heads=20, encoder_ffn_dim=5120, decoder_ffn_dim=5120))]:
        d = tmp_path / name
        transformers.WhisperConfig(**kw).save_pretrained(str(d))            # writes config.json the way optimum leaves it
        cfg = wb.default_cfg("base", precision=wb.WB_PREC_BF16, max_batch=7)
        assert L.wb_cfg_from_hf_config(str(d / "config.json").encode(), C.byref(cfg)) == 0, L.wb_last_error()
        assert wb.binding.model_cfg_of(cfg) == wb.binding.model_cfg_of(wb.default_cfg(name))
        assert (cfg.precision, cfg.max_batch) == (wb.WB_PREC_BF16, 7)        # everything else is left alone
    bad = tmp_path / "bad.json"
    bad.write_text(json.dumps({"d_model": 512, "num_mel_bins": 80}))
    cfg = wb.default_cfg("base")
    assert L.wb_cfg_from_hf_config(str(bad).encode(), C.byref(cfg)) != 0 and b"encoder_attention_heads" in L.wb_last_error()
    assert wb.binding.model_cfg_of(cfg) == wb.weights.WHISPER_BASE            # untouched on failure
    ref_cfg = "/root/reference/config.json"
    if os.path.exists(ref_cfg) and "d_model" in open(ref_cfg).read():
        assert L.wb_cfg_from_hf_config(ref_cfg.encode(), C.byref(cfg)) == 0
