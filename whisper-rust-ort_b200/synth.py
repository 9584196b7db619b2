"""Deterministic synthetic 16 kHz clips: the reference's only input,
`audio/audio.wav`, is a missing large blob, so parity tests and the bench run on these."""
from __future__ import annotations

import struct

import numpy as np

SR = 16000


def clip(i: int, seed: int = 0, seconds: float = 30.0) -> np.ndarray:
    """Clip `i` of stream `seed`: f32 mono in [-1, 1].  Three families by i % 3:
    (A) gaussian sigma 0.1, (B) three log-uniform 80-4000 Hz sines amp 0.2 + noise, (C) A with a
    slow 2 Hz envelope."""
    rng = np.random.default_rng(seed * 1_000_003 + i)
    n = int(round(seconds * SR))
    t = np.arange(n, dtype=np.float64) / SR
    kind = i % 3
    if kind == 0:
        x = rng.normal(0.0, 0.1, n)
    elif kind == 1:
        f = np.exp(rng.uniform(np.log(80.0), np.log(4000.0), 3))
        ph = rng.uniform(0, 2 * np.pi, 3)
        x = sum(0.2 * np.sin(2 * np.pi * fk * t + pk) for fk, pk in zip(f, ph))
        x = x + rng.normal(0.0, 0.01, n)
    else:
        x = rng.normal(0.0, 0.1, n) * (0.5 + 0.5 * np.sin(2 * np.pi * 2.0 * t))
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def batch(n_clips: int, seed: int = 0, seconds: float = 30.0, start: int = 0) -> np.ndarray:
    return np.stack([clip(start + i, seed, seconds) for i in range(n_clips)])


def fast_batch(n_clips: int, seed: int = 0, n_unique: int = 8, seconds: float = 30.0) -> np.ndarray:
    """Bench helper: `n_unique` real clips tiled with a per-clip gain so rows differ; avoids
    minutes of host RNG time for 1024-clip workloads."""
    base = batch(min(n_unique, n_clips), seed, seconds)
    reps = (n_clips + base.shape[0] - 1) // base.shape[0]
    x = np.tile(base, (reps, 1))[:n_clips].copy()
    gains = (0.5 + 0.5 * ((np.arange(n_clips) * 37 % 101) / 100.0)).astype(np.float32)
    x *= gains[:, None]
    return x


def write_wav(path: str, pcm: np.ndarray, sr: int = SR, fmt: str = "s16", channels: int = 1) -> None:
    """Minimal RIFF/WAVE writer for tests of the host WAV reader (`load_audio_16k_mono`,
    /root/reference/src/main.rs:228-316 accepts u8/s16/f32 decoded buffers)."""
    x = np.asarray(pcm, dtype=np.float32)
    if channels > 1:
        x = np.repeat(x[:, None], channels, axis=1).reshape(-1)
    if fmt == "s16":
        data = np.clip(np.round(x * 32768.0), -32768, 32767).astype("<i2").tobytes()
        tag, bits = 1, 16
    elif fmt == "u8":
        data = np.clip(np.round(x * 128.0 + 128.0), 0, 255).astype(np.uint8).tobytes()
        tag, bits = 1, 8
    elif fmt == "f32":
        data = x.astype("<f4").tobytes()
        tag, bits = 3, 32
    else:
        raise ValueError(fmt)
    block = channels * bits // 8
    if channels <= 2:
        fmt_body = struct.pack("<HHIIHH", tag, channels, sr, sr * block, block, bits)
    else:
        # symphonia's fmt_pcm / fmt_ieee chunks are mono or stereo only: more channels need WAVE_FORMAT_EXTENSIBLE
        # with one channel-mask bit per channel
        fmt_body = (struct.pack("<HHIIHH", 0xFFFE, channels, sr, sr * block, block, bits) + struct.pack("<HHI", 22, bits, (1 << channels) - 1)
                    + struct.pack("<H", tag) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71")
    pad = b"\0" * (len(data) & 1)
    hdr = b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt_body) + 8 + len(data) + len(pad)) + b"WAVE"
    hdr += b"fmt " + struct.pack("<I", len(fmt_body)) + fmt_body
    hdr += b"data" + struct.pack("<I", len(data))
    data += pad
    with open(path, "wb") as f:
        f.write(hdr + data)
