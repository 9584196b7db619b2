"""oracle/host_ref.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python restatement of the host-side (non-tensor) functions on the reference's hot path, each
following /root/reference/src/main.rs line by line; used by tests/ to check the C++ host in
whisper-rust-ort_b200/csrc/host/.  Pinned by the reference's own committed outputs where they
exist (results.old/.../inference_summary.json, inference_per_file.csv: schema, rounding, key order).

Audio decode: the RIFF/WAVE reader is restated here (read_wav_symphonia).  MPEG Layer III has NO restatement in
oracle/: its decoder lives in a dependency (symphonia-bundle-mp3 0.5.x) absent from /root/reference, and MP3
decoding is only defined up to floating-point rounding, so a second implementation by the same hand would pin
nothing.  The checker for csrc/host/mp3.cpp is an independent conforming decoder instead (tests/libav_ref.py:
libavcodec's mp3float), fed by tests/mp3_writer.py; see DESIGN.md section 3.
"""
from __future__ import annotations

import math

import numpy as np


def resample_linear(x: np.ndarray, sr_in: int, sr_out: int) -> np.ndarray:
    """main.rs:207-226."""
    x = np.asarray(x, np.float32)
    if sr_in == sr_out:
        return x.copy()
    ratio = sr_out / sr_in
    n_out = int(math.floor(len(x) * ratio + 0.5))          # f64::round (half away from zero, x >= 0)
    i = np.arange(n_out, dtype=np.float64)
    t = i / ratio
    i0 = np.floor(t).astype(np.int64)
    i1 = i0 + 1
    a = t - i0
    xp = np.concatenate([x, np.zeros(2, np.float32)])
    s0 = np.where((i0 >= 0) & (i0 < len(x)), xp[np.clip(i0, 0, len(x))], np.float32(0))
    s1 = np.where((i1 >= 0) & (i1 < len(x)), xp[np.clip(i1, 0, len(x))], np.float32(0))
    return ((1.0 - a).astype(np.float32) * s0 + a.astype(np.float32) * s1).astype(np.float32)


def decode_wav_samples(raw: np.ndarray, fmt: str, channels: int) -> np.ndarray:
    """Sample conversion + channel-mean downmix of load_audio_16k_mono (main.rs:266-301)."""
    if fmt == "u8":
        v = (raw.astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
    elif fmt == "s16":
        v = raw.astype(np.float32) / np.float32(32768.0)
    else:
        v = raw.astype(np.float32)
    v = v.reshape(-1, channels)
    acc = np.zeros(v.shape[0], np.float32)
    for c in range(channels):
        acc = acc + v[:, c]
    return acc / np.float32(channels)


_ALAW = None


def _g711_tables():
    """ITU-T G.711 expansion tables (what symphonia-codec-pcm applies to PCM_ALAW / PCM_MULAW packets -> S16)."""
    global _ALAW
    if _ALAW is None:
        al, mu = np.zeros(256, np.int16), np.zeros(256, np.int16)
        for v in range(256):
            a = v ^ 0x55
            t, seg = (a & 0x0F) << 4, (a & 0x70) >> 4
            t = t + 8 if seg == 0 else t + 0x108 if seg == 1 else (t + 0x108) << (seg - 1)
            al[v] = t if a & 0x80 else -t
            u = (~v) & 0xFF
            t = (((u & 0x0F) << 3) + 0x84) << ((u & 0x70) >> 4)
            mu[v] = (0x84 - t) if u & 0x80 else (t - 0x84)
        _ALAW = (al, mu)
    return _ALAW


class WavError(ValueError):
    pass


def read_wav_symphonia(blob: bytes):
    """What load_audio_16k_mono (main.rs:228-316) gets out of symphonia 0.5.5's RIFF/WAVE reader + PCM decoder, before
    resampling: -> (mono f32 samples, sample_rate).  symphonia-format-riff is a Cargo.lock dependency that is not under
    /root/reference; this restates its published algorithm (ChunksReader::next, WavReader::try_new, next_packet):
    the RIFF length bounds the chunk walk, a chunk longer than the rest of its parent is an error unless both lengths
    are 0xFFFFFFFF, chunks are word aligned, the walk stops at the first data chunk, packets are <= 1152 blocks cut
    from the DECLARED data length and a packet running past EOF is dropped whole (IoError -> `break`, main.rs:258-262)."""
    import struct
    if len(blob) < 12:
        raise WavError("short read")
    if blob[:4] == b"fLaC":
        raise WavError("Unsupported decoded sample format")          # FLAC decodes to S32 -> main.rs:303
    if blob[:4] != b"RIFF":
        raise WavError("unsupported audio container")
    if blob[8:12] != b"WAVE":
        raise WavError("wav: riff form is not wave")
    riff_len = struct.unpack_from("<I", blob, 4)[0]
    consumed, pos, fmt = 0, 12, None
    while True:
        if consumed & 1:
            if pos >= len(blob):
                raise WavError("end of stream")
            pos += 1
            consumed += 1
        if consumed + 8 > riff_len:
            raise WavError("wav: missing data chunk")
        if pos + 8 > len(blob):
            raise WavError("end of stream")
        cid, ln = blob[pos:pos + 4], struct.unpack_from("<I", blob, pos + 4)[0]
        pos += 8
        consumed += 8
        if riff_len - consumed < ln and not (riff_len == ln == 0xFFFFFFFF):
            raise WavError("riff: chunk length exceeds parent (list) chunk length")
        consumed = min(consumed + ln, 0xFFFFFFFF)
        if cid == b"fmt ":
            if ln < 16:
                raise WavError("wav: malformed fmt chunk")
            if pos + ln > len(blob):
                raise WavError("end of stream")
            tag, ch, sr, _, align, bits = struct.unpack_from("<HHIIHH", blob, pos)
            def lr(what):
                if ch not in (1, 2):
                    raise WavError(f"wav: channel layout is not stereo or mono for {what}")
            if tag == 1:
                if ln not in (16, 18, 40):
                    raise WavError("wav: malformed fmt_pcm chunk")
                if bits not in (8, 16, 24, 32):
                    raise WavError("wav: bits per sample for fmt_pcm must be 8, 16, 24 or 32 bits")
                lr("fmt_pcm")
                codec = {8: "u8", 16: "s16"}.get(bits, "other")
            elif tag == 3:
                if ln not in (16, 18):
                    raise WavError("wav: malformed fmt_ieee chunk")
                if ln == 18 and struct.unpack_from("<H", blob, pos + 16)[0] != 0:
                    raise WavError("wav: extension length not 0 for fmt_ieee chunk")
                if bits not in (32, 64):
                    raise WavError("wav: bits per sample for fmt_ieee must be 32 or 64 bits")
                lr("fmt_ieee")
                codec = "f32" if bits == 32 else "other"
            elif tag == 0xFFFE:
                if ln != 40:
                    raise WavError("wav: malformed fmt_ext chunk")
                ext, valid, mask = struct.unpack_from("<HHI", blob, pos + 16)
                if ext != 22:
                    raise WavError("wav: extension length not 22 for fmt_ext chunk")
                if bits & 7:
                    raise WavError("wav: bits per sample for fmt_ext must be a multiple of 8")
                if valid > bits:
                    raise WavError("wav: bits per sample exceeds coded bits per sample for fmt_ext")
                if bin(mask).count("1") != ch:
                    raise WavError("wav: channel mask mismatch for fmt_ext")
                if mask >> 26:
                    raise WavError("wav: too many channel masks")
                if blob[pos + 26:pos + 40] != bytes.fromhex("000000001000800000aa00389b71"):
                    raise WavError("wav: unsupported fmt_ext sub-type")
                sub = struct.unpack_from("<H", blob, pos + 24)[0]
                if sub == 1 and bits in (8, 16, 24, 32):
                    codec = {8: "u8", 16: "s16"}.get(bits, "other")
                elif sub == 3 and bits in (32, 64):
                    codec = "f32" if bits == 32 else "other"
                else:
                    raise WavError("wav: unsupported fmt_ext sub-type")
            elif tag in (6, 7):
                if ln != 18:
                    raise WavError("wav: malformed fmt_alaw/fmt_mulaw chunk")
                if bits != 8:
                    raise WavError("wav: bits per sample for fmt_alaw/fmt_mulaw must be 8 bits")
                lr("fmt_alaw" if tag == 6 else "fmt_mulaw")
                codec = "alaw" if tag == 6 else "mulaw"
            elif tag in (2, 0x11):
                codec = "other"                                      # ADPCM decodes to S32 -> main.rs:303
            else:
                raise WavError("wav: unsupported wave format")
            fmt = (codec, ch, sr, align, bits)
        elif cid == b"data":
            break
        pos += ln
        if pos > len(blob):
            raise WavError("end of stream")
    if fmt is None:
        raise WavError("No default track")
    codec, ch, sr, align, bits = fmt
    if sr == 0:
        raise WavError("Unknown sample rate")
    if ch == 0:
        raise WavError("Unknown channels")
    if codec == "other":
        raise WavError("Unsupported decoded sample format")
    if align == 0:
        raise WavError("riff: block size is 0")
    if align != ch * bits // 8:
        raise WavError("wav: block_align does not match")
    avail, blocks, frames = len(blob) - pos, ln // align, 0
    while frames < blocks:
        n = min(1152, blocks - frames)
        if (frames + n) * align > avail:
            break
        frames += n
    raw = blob[pos:pos + frames * align]
    if codec in ("alaw", "mulaw"):
        tab = _g711_tables()[0 if codec == "alaw" else 1]
        return decode_wav_samples(tab[np.frombuffer(raw, np.uint8)], "s16", ch), sr
    dt = {"u8": np.uint8, "s16": "<i2", "f32": "<f4"}[codec]
    return decode_wav_samples(np.frombuffer(raw, dtype=dt), codec, ch), sr


# Rust's char::is_whitespace = the Unicode White_Space property.  Python's str.split()/strip() also treat
# U+001C..U+001F as separators, which Rust does not, so the Rust set is spelled out.
RUST_WHITESPACE = set(range(0x09, 0x0E)) | {0x20, 0x85, 0xA0, 0x1680} | set(range(0x2000, 0x200B)) | {0x2028, 0x2029, 0x202F, 0x205F, 0x3000}


def split_whitespace(s: str) -> list[str]:
    """str::split_whitespace."""
    out, cur = [], ""
    for ch in s:
        if ord(ch) in RUST_WHITESPACE:
            if cur:
                out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur:
        out.append(cur)
    return out


def trim(s: str) -> str:
    """str::trim."""
    b, e = 0, len(s)
    while b < e and ord(s[b]) in RUST_WHITESPACE:
        b += 1
    while e > b and ord(s[e - 1]) in RUST_WHITESPACE:
        e -= 1
    return s[b:e]


def word_overlap(a: str, b: str, max_words: int) -> int:
    """main.rs:686-696."""
    aw = [w.lower() for w in split_whitespace(a)]
    bw = [w.lower() for w in split_whitespace(b)]
    mx = min(max_words, len(aw), len(bw))
    for k in range(mx, 0, -1):
        if aw[len(aw) - k:] == bw[:k]:
            return k
    return 0


def stitch_texts(chunks: list[str]) -> str:
    """main.rs:659-684."""
    out = ""
    for chunk in chunks:
        t = trim(chunk)
        if not t:
            continue
        if not out:
            out = t
            continue
        ov = word_overlap(out, t, 16)
        if ov > 0:
            rem = " ".join(split_whitespace(t)[ov:])
            if rem:
                out += " " + rem
        else:
            out += " " + t
    return out


def percentile(xs, p: float) -> float:
    """main.rs:1021-1031."""
    if len(xs) == 0:
        return float("nan")
    v = sorted(xs)
    k = (len(v) - 1.0) * (p / 100.0)
    f, c = int(math.floor(k)), int(math.ceil(k))
    if f == c:
        return v[f]
    return v[f] + (v[c] - v[f]) * (k - f)


def stat_block(xs) -> dict:
    """main.rs:1033-1048 (median = upper median)."""
    v = sorted(xs)
    total = 0.0
    for x in v:                 # Rust's iter().sum::<f64>() is a plain left-to-right fold over the SORTED values;
        total += x              # Python >= 3.12 sum() is Neumaier-compensated and can differ in the last bit
    return {"min": v[0], "median": v[len(v) // 2], "p90": percentile(xs, 90.0), "p95": percentile(xs, 95.0),
            "max": v[-1], "mean": total / len(v)}


def special_tokens(language: str, task: str, token_to_id=None):
    """main.rs:528-569 -> (sot, eot, lang, task, no_timestamps)."""
    if token_to_id is not None:
        def get(t):
            if t not in token_to_id:
                raise KeyError(f"Tokenizer missing token: {t}")
            return token_to_id[t]
        return (get("<|startoftranscript|>"), get("<|endoftext|>"), get(f"<|{language}|>"), get(f"<|{task}|>"),
                get("<|notimestamps|>"))
    lang = {"en": 50259, "hi": 50276}.get(language, 50259)
    task_tok = {"transcribe": 50359, "translate": 50358}.get(task, 50359)
    return 50258, 50257, lang, task_tok, 50363


def decode_tokens_fallback(tokens) -> str:
    """main.rs:644-647."""
    return "[TOKENS:" + " ".join(str(int(t)) for t in list(tokens)[:200]) + "]"


def bytes_to_unicode() -> dict[int, str]:
    """GPT-2 byte-level alphabet (what tokenizers' ByteLevel decoder inverts)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return {b: chr(c) for b, c in zip(bs, cs)}


def chunk_starts(n_samples: int, chunk_len: int = 480000, step: int = 400000) -> list[int]:
    """main.rs:875-882."""
    out, pos = [], 0
    while pos < n_samples:
        end = min(pos + chunk_len, n_samples)
        out.append(pos)
        if end == n_samples:
            break
        pos += step
    return out
