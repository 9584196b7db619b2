"""oracle/host_ref.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python restatement of the host-side (non-tensor) functions on the reference's hot path, each
following /root/reference/src/main.rs line by line; used by tests/ to check the C++ host in
whisper-rust-ort_b200/csrc/host/.  Pinned by the reference's own committed outputs where they
exist (results.old/.../inference_summary.json, inference_per_file.csv: schema, rounding, key order).
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
