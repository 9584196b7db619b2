"""CPU: randomised (hypothesis) agreement of the C++ host functions with oracle/host_ref.py — the hand-picked
cases of test_host_cpu.py widened to arbitrary word lists, sample counts and rate pairs, plus the structural
properties the reference relies on (chunk windows cover the file, stitching never duplicates an overlap)."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import host_ref as hr
from test_host_cpu import L, load_audio, stitch            # noqa: F401  (fixture + helpers)

WORDS = ["the", "The", "cat", "sat", "on", "mat", "über", "naïve", "日本", "a", "A", "and", "then", "went", "home", "x1", "...",
         "ΣΟΦΟΣ", "σοφος", "İZ", "i̇z", "a\x1fb", "a\x1cB", "no\u200bbreak", "Ǆ", "ǆ"]
SEPS = [" ", "  ", "\t", "\n", "\u00a0", "\u2003", "\u3000", "\x0b", "\x85", "\u2028"]            # all White_Space
words = st.lists(st.sampled_from(WORDS), min_size=0, max_size=24)


def _join(ws, rng_spaces):
    return "".join(w + SEPS[s % len(SEPS)] for w, s in zip(ws, rng_spaces)) if ws else ""


@settings(max_examples=150, deadline=None)
@given(st.lists(words, min_size=0, max_size=6), st.integers(0, 2**31 - 1))
def test_stitching_matches_oracle_on_random_word_lists(L, chunks_w, seed):
    rng = np.random.default_rng(seed)
    chunks = []
    for i, ws in enumerate(chunks_w):
        if i and chunks_w[i - 1] and rng.random() < 0.6:          # make a real overlap with the previous chunk
            k = int(rng.integers(1, min(len(chunks_w[i - 1]), 18) + 1))
            ws = chunks_w[i - 1][-k:] + ws
        chunks.append(_join(ws, rng.integers(0, 1000, len(ws))))
    assert stitch(L, chunks) == hr.stitch_texts(chunks)
    for a, b in zip(chunks, chunks[1:]):
        for mw in (1, 4, 16):
            assert L.wb_host_word_overlap(a.encode(), b.encode(), mw) == hr.word_overlap(a, b, mw)


@settings(max_examples=150, deadline=None)
@given(st.lists(st.floats(min_value=-1e6, max_value=1e6, allow_nan=False, width=64), min_size=1, max_size=60),
       st.floats(min_value=0.0, max_value=100.0))
def test_percentile_matches_oracle_and_is_monotone(L, xs, p):
    a = (C.c_double * len(xs))(*xs)
    v = L.wb_host_percentile(a, len(xs), p)
    assert v == hr.percentile(xs, p)
    assert min(xs) <= v <= max(xs)
    assert L.wb_host_percentile(a, len(xs), 0.0) <= v <= L.wb_host_percentile(a, len(xs), 100.0)
    out = (C.c_double * 6)()
    assert L.wb_host_stat_block(a, len(xs), out) == 0
    ref = hr.stat_block(xs)
    assert list(out) == [ref["min"], ref["median"], ref["p90"], ref["p95"], ref["max"], ref["mean"]]


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 5_000_000), st.integers(1, 600_000), st.integers(0, 600_000))
def test_chunk_windows_match_oracle_and_cover_the_file(L, n, chunk_len, overlap):
    step = max(chunk_len - overlap, 1) if chunk_len > overlap else 1
    k = L.wb_host_chunk_starts(n, chunk_len, step, None, 0)
    if k > 4096:
        return
    buf = (C.c_int64 * max(k, 1))()
    assert L.wb_host_chunk_starts(n, chunk_len, step, buf, k) == k
    starts = list(buf[:k])
    assert starts == hr.chunk_starts(n, chunk_len, step)
    assert starts[0] == 0 and all(b - a == step for a, b in zip(starts, starts[1:]))
    assert starts[-1] < n <= starts[-1] + max(chunk_len, step)          # the last window reaches the end, none starts past it


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 4000), st.sampled_from([8000, 11025, 16000, 22050, 32000, 44100, 48000]), st.integers(0, 2**31 - 1))
def test_resample_matches_oracle(L, n, sr, seed):
    x = np.random.default_rng(seed).uniform(-1, 1, n).astype(np.float32)
    f32p = C.POINTER(C.c_float)
    m = L.wb_host_resample_linear(x.ctypes.data_as(f32p), n, sr, 16000, None, 0)
    ref = hr.resample_linear(x, sr, 16000)
    assert m == len(ref)
    out = np.zeros(max(m, 1), np.float32)
    L.wb_host_resample_linear(x.ctypes.data_as(f32p), n, sr, 16000, out.ctypes.data_as(f32p), m)
    assert np.array_equal(out[:m], ref)


# ---------------- f64 formatting of the output files (serde_json / ryu) ----------------
def fmt(L, v):
    L.wb_host_format_f64.argtypes = [C.c_double, C.c_char_p, C.c_int64]
    L.wb_host_format_f64.restype = C.c_int64
    buf = C.create_string_buffer(64)
    n = L.wb_host_format_f64(v, buf, 64)
    assert n == len(buf.value)
    return buf.value.decode()


# every float the reference's Rust binary wrote into its committed result files
# (/root/reference/results.old/**/without_hf_pipeline_rust*/*.json), as printed by serde_json
RUST_PRINTED = ['7.2111', '9.1312', '14.8844', '17.0939', '301.574', '0.023911', '0.030278', '0.049356', '0.056682', '0.2028574',
                '0.42281892', '0.68677424', '8.65831095', '0.000482736', '0.000507354', '0.000559743', '0.000967648', '0.048492999',
                '0.053896242', '0.271739968', '0.412314905', '0.647183284', '6.741848167', '9.131170268', '14.031795815',
                '16.130229141', '17.093930471', '14.884440201999999', '7.2110552420000005', '0.04935592882663761',
                '0.05668246868839603', '0.023911435328369368', '0.030278423893346063']
# ryu's documented layout switches
RYU_LAYOUT = [(1.0, "1.0"), (0.0, "0.0"), (-0.0, "-0.0"), (1e16, "1e16"), (1e15, "1000000000000000.0"),
              (123456789012345680.0, "1.2345678901234568e17"), (0.1 + 0.2, "0.30000000000000004"), (1e-5, "0.00001"), (1e-6, "1e-6"),
              (4.82e-6, "4.82e-6"), (0.0000482, "0.0000482"), (5e-324, "5e-324"), (1.7976931348623157e308, "1.7976931348623157e308"),
              (-2.5, "-2.5"), (100.0, "100.0"), (1.5e-7, "1.5e-7"),
              (2.0 ** -24, "5.960464477539063e-8"),       # exact decimal tie: the correctly ROUNDED 16 digits do not read back
              (9007199254740993.0, "9007199254740992.0"), (0.3, "0.3"), (2.0 ** 70, "1.1805916207174113e21")]


def test_f64_formatting_reproduces_what_the_rust_binary_printed(L):
    for t in RUST_PRINTED:
        assert fmt(L, float(t)) == t
    for v, t in RYU_LAYOUT:
        assert fmt(L, v) == t, (v, t)
    assert fmt(L, float("nan")) == "null" and fmt(L, float("inf")) == "null"
    import glob, os, re
    files = glob.glob("/root/reference/results.old/**/without_hf_pipeline_rust*/*.json", recursive=True)
    for f in files:                                   # live re-check when the reference is mounted (this container)
        for m in re.finditer(r'(?<![\\w"])-?\\d+\\.\\d+(?:e-?\\d+)?', open(f).read()):
            assert fmt(L, float(m.group(0))) == m.group(0)


@settings(max_examples=400, deadline=None)
@given(st.floats(allow_nan=False, allow_infinity=False, width=64))
def test_f64_formatting_round_trips_with_shortest_digits(L, v):
    s = fmt(L, v)
    assert float(s) == v and (v != 0 or s in ("0.0", "-0.0"))
    digits = lambda t: t.lstrip("-").split("e")[0].replace(".", "").strip("0")
    assert digits(s) == digits(repr(v))               # same shortest digit string as Python's repr (Gay / ryu agree)
    assert ("e" in s) or ("." in s)


# ---------------- byte-level BPE decode: bytes -> String::from_utf8_lossy ----------------
@pytest.fixture(scope="module")
def byte_tok(L, tmp_path_factory):
    """A tokenizer.json whose vocabulary is the 256 single-byte tokens (ids 0..255) plus a few specials."""
    import json
    b2u = hr.bytes_to_unicode()
    vocab = {b2u[b]: b for b in range(256)}
    added = [{"id": 300, "content": "<|endoftext|>", "special": True}, {"id": 301, "content": "<|0.00|>", "special": False},
             {"id": 302, "content": "café au lait", "special": False}]
    p = tmp_path_factory.mktemp("bytetok") / "tokenizer.json"
    p.write_text(json.dumps({"model": {"type": "BPE", "vocab": vocab}, "added_tokens": added, "decoder": {"type": "ByteLevel"}}), encoding="utf-8")
    L.wb_tokenizer_load.argtypes = [C.POINTER(C.c_void_p), C.c_char_p]
    t = C.c_void_p()
    assert L.wb_tokenizer_load(C.byref(t), str(p).encode()) == 0, L.wb_last_error()
    yield t
    L.wb_tokenizer_free(t)


def _decode(L, t, ids):
    arr = (C.c_int64 * max(len(ids), 1))(*ids)
    n = L.wb_host_decode_tokens(t, arr, len(ids), None, 0)
    buf = C.create_string_buffer(n + 1)
    L.wb_host_decode_tokens(t, arr, len(ids), buf, n + 1)
    return buf.raw[:n]


@settings(max_examples=400, deadline=None)
@given(st.binary(min_size=0, max_size=40))
def test_decode_of_arbitrary_bytes_is_from_utf8_lossy(L, byte_tok, data):
    got = _decode(L, byte_tok, list(data))
    assert got == data.decode("utf-8", errors="replace").encode("utf-8")       # same maximal-subpart policy as Rust


def test_decode_mixes_specials_added_tokens_and_partial_sequences(L, byte_tok):
    e_acute = list("é".encode())                                               # 0xC3 0xA9 as two single-byte tokens
    assert _decode(L, byte_tok, [0x41, 300, *e_acute, 301]).decode() == "Aé<|0.00|>"       # special skipped, added kept
    assert _decode(L, byte_tok, [e_acute[0], 301]).decode() == "�<|0.00|>"            # dangling lead byte
    # an added token with a character outside the byte-level alphabet (the space) is taken as raw text, like the crate
    assert _decode(L, byte_tok, [302]).decode() == "café au lait"


# ---------------- str::to_lowercase (full Unicode mapping + Final_Sigma) ----------------
def lower(L, s):
    L.wb_host_to_lowercase.argtypes = [C.c_char_p, C.c_char_p, C.c_int64]
    L.wb_host_to_lowercase.restype = C.c_int64
    b = s.encode("utf-8")
    n = L.wb_host_to_lowercase(b, None, 0)
    buf = C.create_string_buffer(n + 1)
    L.wb_host_to_lowercase(b, buf, n + 1)
    return buf.raw[:n].decode("utf-8")


def test_lowercase_of_every_scalar_value_matches_the_unicode_database(L):
    """CPython's str.lower() implements the algorithm Rust's str::to_lowercase does (same tables, same Final_Sigma
    rule), so it is the oracle here; the header is generated from it by tools/gen_unicode_tables.py."""
    chunk = []
    for c in range(1, 0x110000):
        if 0xD800 <= c <= 0xDFFF:
            continue
        chunk.append(chr(c))
        if len(chunk) == 4096:
            s = " ".join(chunk)                        # spaces: no Final_Sigma context between neighbours
            assert lower(L, s) == " ".join(ch.lower() for ch in chunk)
            chunk = []
    s = " ".join(chunk)
    assert lower(L, s) == " ".join(ch.lower() for ch in chunk)


SIGMA_WORDS = ["ΟΔΥΣΣΕΥΣ", "ΣΟΦΟΣ", "Σ", "ΑΣ", "ΑΣ.", "ΑΣ'Α", "ΑΣ́", "ΆΣ", "ΣΑΣ", "aΣ", "Σa", "1Σ", "ΑΣ1", "ΑΣ·Β", "İSTANBUL", "ǅUNGLA",
               "ẞ", "ΆΈΉ", "ԱԲԳ", "ᲐᲑᲒ", "Ｈｅｌｌｏ", "ǄǇ", "Ⅷ", "Ⓐ", "𐐀𐐁", "Straße", "ÀÉÎÕÜ"]


@settings(max_examples=300, deadline=None)
@given(st.lists(st.one_of(st.sampled_from(SIGMA_WORDS), st.text(alphabet=st.characters(blacklist_categories=("Cs",)), min_size=0, max_size=8)),
                min_size=0, max_size=6))
def test_lowercase_of_random_text_matches_python(L, parts):
    s = "".join(parts)
    if "\x00" in s:
        return                                          # C strings end at NUL
    assert lower(L, s) == s.lower()


def test_word_overlap_is_case_insensitive_in_every_script(L):
    a, b = "είπε ο ΟΔΥΣΣΕΥΣ ΣΤΗΝ Ιθάκη", "οδυσσευς στην ιθάκη και μετά"
    assert L.wb_host_word_overlap(a.encode(), b.encode(), 16) == hr.word_overlap(a, b, 16) == 3
    a, b = "geldik İSTANBUL ŞEHRİ", "i̇stanbul şehri̇ çok güzel"
    assert L.wb_host_word_overlap(a.encode(), b.encode(), 16) == hr.word_overlap(a, b, 16) == 2


# ---------------- string output: serde_json literals and csv-crate fields ----------------
def _call_str(L, fn, s):
    f = getattr(L, fn)
    f.argtypes = [C.c_char_p, C.c_char_p, C.c_int64]
    f.restype = C.c_int64
    b = s.encode("utf-8")
    n = f(b, None, 0)
    buf = C.create_string_buffer(n + 1)
    f(b, buf, n + 1)
    return buf.raw[:n].decode("utf-8")


TEXT = st.text(alphabet=st.one_of(st.sampled_from(list(' ,";\'\\/\t\n\r\x08\x0c\x01\x1f\x7fé日𐐀')), st.characters(blacklist_categories=("Cs",), min_codepoint=1)),
               min_size=0, max_size=30)


@settings(max_examples=300, deadline=None)
@given(TEXT)
def test_json_string_matches_serde_json_escaping(L, s):
    import json
    got = _call_str(L, "wb_host_json_string", s)
    assert got == json.dumps(s, ensure_ascii=False)          # \\" \\\\ \\b \\f \\n \\r \\t, \\u00xx below 0x20, everything else raw
    assert json.loads(got) == s


@settings(max_examples=300, deadline=None)
@given(TEXT)
def test_csv_field_matches_minimal_quoting(L, s):
    import csv, io
    buf = io.StringIO()
    csv.writer(buf, quoting=csv.QUOTE_MINIMAL, lineterminator="\n").writerow(["x", s, "y"])
    want = buf.getvalue()[2:-3]                               # strip 'x,' and ',y\\n'
    assert _call_str(L, "wb_host_csv_field", s) == want
    assert next(csv.reader(io.StringIO("x," + want + ",y\n")))[1] == s.replace("\r\n", "\r\n")


# ---------------- RIFF/WAVE container variations ----------------
GUID_TAIL = bytes.fromhex("000000001000800000aa00389b71")


def _riff(chunks, riff_len=None, tail=b""):
    body = b"WAVE"
    for cid, payload, declared in chunks:
        body += cid + np.uint32(len(payload) if declared is None else declared).tobytes() + payload
        if len(payload) & 1:
            body += b"\0"                                    # RIFF chunks are word aligned
    body += tail
    return b"RIFF" + np.uint32(len(body) if riff_len is None else riff_len).tobytes() + body


@settings(max_examples=200, deadline=None)
@given(st.sampled_from(["u8", "s16", "f32", "alaw", "mulaw"]), st.integers(1, 3), st.integers(0, 2600), st.sampled_from([16, 18, 40]),
       st.lists(st.tuples(st.sampled_from([b"LIST", b"fact", b"bext", b"junk"]), st.binary(min_size=0, max_size=33)), max_size=3),
       st.sampled_from(["exact", "streamed", "too_long", "short_riff", "mask0"]), st.binary(max_size=40), st.integers(0, 2**31 - 1))
def test_wav_container_variations(L, tmp_path_factory, fmt, channels, frames, fmt_len, extra, size_kind, tail, seed):
    """The RIFF/WAVE reader against the restatement of symphonia-format-riff 0.5.5 (oracle/host_ref.py::read_wav_symphonia):
    fmt chunks of 16/18/40 bytes, foreign chunks with odd sizes before the data, odd-length data + pad byte, more than two
    channels (extensible only), A-law/mu-law, and the sizes streaming writers leave behind: `streamed` = RIFF and data
    length both 0xFFFFFFFF (accepted; whole 1152-frame packets only, the partial packet at EOF is dropped), `too_long` =
    data length beyond its parent (the reference fails), `short_riff` = RIFF length that hides the data chunk."""
    import struct
    rng = np.random.default_rng(seed)
    if fmt == "u8":
        raw = rng.integers(0, 256, frames * channels).astype(np.uint8); tag, bits = 1, 8
    elif fmt == "s16":
        raw = rng.integers(-32768, 32768, frames * channels).astype("<i2"); tag, bits = 1, 16
    elif fmt == "f32":
        raw = rng.uniform(-1, 1, frames * channels).astype("<f4"); tag, bits = 3, 32
    else:
        raw = rng.integers(0, 256, frames * channels).astype(np.uint8); tag, bits = (6 if fmt == "alaw" else 7), 8
    block = channels * bits // 8
    mask = 0 if size_kind == "mask0" else (1 << channels) - 1
    f = struct.pack("<HHIIHH", 0xFFFE if fmt_len == 40 else tag, channels, 16000, 16000 * block, block, bits)
    if fmt_len == 18:
        f += struct.pack("<H", 0)
    elif fmt_len == 40:
        f += struct.pack("<HHI", 22, bits, mask) + struct.pack("<H", tag) + GUID_TAIL
    data = raw.tobytes()
    chunks = [(b"fmt ", f, None)] + [(cid, payload, None) for cid, payload in extra]
    if size_kind == "streamed":
        blob = _riff(chunks + [(b"data", data, 0xFFFFFFFF)], riff_len=0xFFFFFFFF, tail=tail)
    elif size_kind == "too_long":
        blob = _riff(chunks + [(b"data", data, len(data) + 1000)])
    elif size_kind == "short_riff":
        blob = _riff(chunks + [(b"data", data, None)], riff_len=4 + 8 + len(f) - 2)
    else:
        blob = _riff(chunks + [(b"data", data, None)], tail=tail)
    p = tmp_path_factory.mktemp("wav") / "x.wav"
    p.write_bytes(blob)
    try:
        want, sr = hr.read_wav_symphonia(blob)
    except hr.WavError as e:
        with pytest.raises(RuntimeError) as ei:
            load_audio(L, p)
        assert str(e).split(" for ")[0][:40] in str(ei.value)
        return
    assert sr == 16000
    got, dur = load_audio(L, p)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert dur == len(want) / 16000.0
    # what the contract says, spelled out independently of the restatement
    if size_kind in ("exact",):
        assert len(want) == frames
    if size_kind == "streamed":
        assert len(want) == (len(data) + (len(data) & 1) + len(tail)) // block // 1152 * 1152


def test_wav_contract_examples(L, tmp_path):
    """Known answers of the settled contract: an odd-length u8 data chunk keeps its pad byte out of
    the samples; an over-declared data chunk inside a correctly sized RIFF is the reference's decode error; plain PCM
    with 3 channels is rejected, the same samples as WAVE_FORMAT_EXTENSIBLE are accepted; G.711 goes through."""
    import struct, audioop
    def fmt_chunk(tag, ch, bits, ext=None):
        block = ch * bits // 8
        body = struct.pack("<HHIIHH", tag if ext is None else 0xFFFE, ch, 16000, 16000 * block, block, bits)
        if ext is not None:
            body += struct.pack("<HHI", 22, bits, ext) + struct.pack("<H", tag) + GUID_TAIL
        return (b"fmt ", body, None)
    def load(blob):
        p = tmp_path / "t.wav"
        p.write_bytes(blob)
        return load_audio(L, p)[0]
    one = np.array([200], np.uint8).tobytes()
    assert load(_riff([fmt_chunk(1, 1, 8), (b"data", one, None)])).tolist() == [(200 - 128) / 128]
    with pytest.raises(RuntimeError, match="chunk length exceeds parent"):
        load(_riff([fmt_chunk(1, 1, 8), (b"data", one, 1001)]))
    with pytest.raises(RuntimeError, match="missing data chunk"):
        load(_riff([fmt_chunk(1, 1, 8)]))
    s3 = np.arange(6, dtype="<i2").tobytes()
    with pytest.raises(RuntimeError, match="not stereo or mono"):
        load(_riff([fmt_chunk(1, 3, 16), (b"data", s3, None)]))
    assert load(_riff([fmt_chunk(1, 3, 16, ext=0b111), (b"data", s3, None)])).tolist() == [np.float32(1 / 32768), np.float32(4 / 32768)]
    with pytest.raises(RuntimeError, match="channel mask mismatch"):
        load(_riff([fmt_chunk(1, 3, 16, ext=0b011), (b"data", s3, None)]))
    # streamed sizes: 2500 frames present -> two whole packets survive
    x = np.arange(2500, dtype="<i2")
    got = load(_riff([fmt_chunk(1, 1, 16), (b"data", x.tobytes(), 0xFFFFFFFF)], riff_len=0xFFFFFFFF))
    assert np.array_equal(got, x[:2304].astype(np.float32) / np.float32(32768))
    # G.711 against Python's audioop (an independent implementation of the ITU tables)
    codes = np.arange(256, dtype=np.uint8).tobytes()
    for tag, conv in ((6, audioop.alaw2lin), (7, audioop.ulaw2lin)):
        body = struct.pack("<HHIIHHH", tag, 1, 16000, 16000, 1, 8, 0)
        got = load(_riff([(b"fmt ", body, None), (b"data", codes, None)]))
        want = np.frombuffer(conv(codes, 2), "<i2").astype(np.float32) / np.float32(32768)
        assert np.array_equal(got, want)


# ---------------- tokenizer.json parsing: JSON string escapes, surrogate pairs, odd layouts ----------------
@settings(max_examples=60, deadline=None)
@given(st.lists(st.text(alphabet=st.characters(blacklist_categories=("Cs",), min_codepoint=1), min_size=1, max_size=8), min_size=1, max_size=12,
                unique=True), st.booleans(), st.sampled_from([None, 0, 2]))
def test_tokenizer_json_keys_survive_any_json_spelling(L, tmp_path_factory, keys, ascii_only, indent):
    """The same vocabulary written with \\uXXXX escapes (incl. surrogate pairs for astral characters) or raw UTF-8, compact
    or indented: every key must be found under its real spelling (wbjson::parse vs Python's json)."""
    import json
    L.wb_tokenizer_load.argtypes = [C.POINTER(C.c_void_p), C.c_char_p]
    L.wb_tokenizer_token_to_id.argtypes = [C.c_void_p, C.c_char_p]
    L.wb_tokenizer_token_to_id.restype = C.c_int64
    vocab = {k: i for i, k in enumerate(keys)}
    doc = {"version": "1.0", "truncation": None, "padding": None, "unknown": [1, 2.5e3, -0.0, True, {"a": []}],
           "added_tokens": [{"id": len(keys), "content": "<|endoftext|>", "special": True, "lstrip": False}],
           "decoder": {"type": "ByteLevel", "add_prefix_space": True}, "model": {"type": "BPE", "dropout": None, "vocab": vocab, "merges": ["a b"]}}
    p = tmp_path_factory.mktemp("tokjson") / "tokenizer.json"
    p.write_text(json.dumps(doc, ensure_ascii=ascii_only, indent=indent), encoding="utf-8")
    t = C.c_void_p()
    assert L.wb_tokenizer_load(C.byref(t), str(p).encode()) == 0, L.wb_last_error()
    try:
        for k, i in vocab.items():
            assert L.wb_tokenizer_token_to_id(t, k.encode("utf-8")) == i, k
        assert L.wb_tokenizer_token_to_id(t, b"<|endoftext|>") == len(keys)
    finally:
        L.wb_tokenizer_free(t)
