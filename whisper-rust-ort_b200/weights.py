"""Whisper weight inventory + the seeded random-init generator (host side).

The reference gets its weights from three ONNX files (`/root/reference/src/main.rs:1099-1108`);
none exist offline, so the operative weight source is "random-init weights of the named
architecture" (BASELINE.json north_star).  To make the CUDA library, the numpy oracle and the
tests agree bit-for-bit without shipping a 290 MB file, weights are *defined* by a counter-based
integer hash (splitmix64 finaliser -> sum of four 16-bit lanes, an Irwin-Hall(4) bell curve):
every operation is exact integer arithmetic followed by one f32 multiply and one f32 add, so the
C++ generator in `csrc/weights.cpp` and this numpy one produce identical bits.

Tensor names follow the Hugging Face `WhisperForConditionalGeneration.state_dict()` keys (the
module the reference's ONNX graphs were exported from, `scripts/export_onnx_whisper.py:20-28`),
so the same dict loads into HF for oracle validation.
"""
from __future__ import annotations

import json
import math
import struct
from dataclasses import dataclass, asdict

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
TENSOR_MUL = np.uint64(0xD1B54A32D192ED03)
M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)
# std-dev of the sum of four independent uniform 16-bit lanes
IH4_SD = math.sqrt(4.0 * (65536.0 ** 2 - 1.0) / 12.0)

KIND_RANDOM = 0      # value = u * (std / IH4_SD) + offset
KIND_SINUSOID = 1    # encoder positional table (HF `sinusoids`)


@dataclass(frozen=True)
class ModelCfg:
    """Mirror of `wb_model_cfg` in include/whisper_b200.h (field order matters: ctypes)."""
    n_mels: int = 80
    d_model: int = 512
    n_heads: int = 8
    ffn_dim: int = 2048
    enc_layers: int = 6
    dec_layers: int = 6
    vocab: int = 51865
    n_audio_ctx: int = 1500
    n_text_ctx: int = 448

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_heads


WHISPER_BASE = ModelCfg()
WHISPER_LARGE_V3 = ModelCfg(n_mels=128, d_model=1280, n_heads=20, ffn_dim=5120, enc_layers=32,
                            dec_layers=32, vocab=51866)
# a toy config for fast CPU tests of the oracle / host logic (same code paths, head_dim 64)
WHISPER_TOY = ModelCfg(n_mels=80, d_model=128, n_heads=2, ffn_dim=256, enc_layers=2, dec_layers=2,
                       vocab=1031, n_audio_ctx=1500, n_text_ctx=448)


def tensor_specs(cfg: ModelCfg):
    """Canonical, ordered list of (name, shape, kind, std, offset). The index in this list is
    the tensor id fed to the hash, so the order is part of the weight definition."""
    d, f = cfg.d_model, cfg.ffn_dim
    W, B_, LNW, LNB = 0.02, 0.02, 0.05, 0.02   # std-devs; LN weight is 1 + N(0, 0.05)
    specs = []

    def lin(prefix, out, inp, bias=True):
        specs.append((prefix + ".weight", (out, inp), KIND_RANDOM, W, 0.0))
        if bias:
            specs.append((prefix + ".bias", (out,), KIND_RANDOM, B_, 0.0))

    def ln(prefix):
        specs.append((prefix + ".weight", (d,), KIND_RANDOM, LNW, 1.0))
        specs.append((prefix + ".bias", (d,), KIND_RANDOM, LNB, 0.0))

    def attn(prefix):
        lin(prefix + ".k_proj", d, d, bias=False)     # modeling_whisper.py: k_proj has no bias
        lin(prefix + ".v_proj", d, d)
        lin(prefix + ".q_proj", d, d)
        lin(prefix + ".out_proj", d, d)

    e = "model.encoder"
    specs.append((e + ".conv1.weight", (d, cfg.n_mels, 3), KIND_RANDOM, W, 0.0))
    specs.append((e + ".conv1.bias", (d,), KIND_RANDOM, B_, 0.0))
    specs.append((e + ".conv2.weight", (d, d, 3), KIND_RANDOM, W, 0.0))
    specs.append((e + ".conv2.bias", (d,), KIND_RANDOM, B_, 0.0))
    specs.append((e + ".embed_positions.weight", (cfg.n_audio_ctx, d), KIND_SINUSOID, 0.0, 0.0))
    for i in range(cfg.enc_layers):
        p = f"{e}.layers.{i}"
        attn(p + ".self_attn")
        ln(p + ".self_attn_layer_norm")
        lin(p + ".fc1", f, d)
        lin(p + ".fc2", d, f)
        ln(p + ".final_layer_norm")
    ln(e + ".layer_norm")

    dd = "model.decoder"
    specs.append((dd + ".embed_tokens.weight", (cfg.vocab, d), KIND_RANDOM, W, 0.0))
    specs.append((dd + ".embed_positions.weight", (cfg.n_text_ctx, d), KIND_RANDOM, W, 0.0))
    for i in range(cfg.dec_layers):
        p = f"{dd}.layers.{i}"
        attn(p + ".self_attn")
        ln(p + ".self_attn_layer_norm")
        attn(p + ".encoder_attn")
        ln(p + ".encoder_attn_layer_norm")
        lin(p + ".fc1", f, d)
        lin(p + ".fc2", d, f)
        ln(p + ".final_layer_norm")
    ln(dd + ".layer_norm")
    return specs


def hash_u(seed: int, tensor_id: int, n: int) -> np.ndarray:
    """Centered Irwin-Hall(4) integer in [-131070, 131070] for elements 0..n-1, as int32."""
    with np.errstate(over="ignore"):
        key = np.uint64(seed) * GOLDEN + np.uint64(tensor_id + 1) * TENSOR_MUL
        z = key + (np.arange(n, dtype=np.uint64) + np.uint64(1)) * GOLDEN
        z = (z ^ (z >> np.uint64(30))) * M1
        z = (z ^ (z >> np.uint64(27))) * M2
        z = z ^ (z >> np.uint64(31))
    m = np.uint64(0xFFFF)
    s = (z & m) + ((z >> np.uint64(16)) & m) + ((z >> np.uint64(32)) & m) + ((z >> np.uint64(48)) & m)
    return s.astype(np.int64).astype(np.int32) - np.int32(131070)


def sinusoid_table(length: int, channels: int) -> np.ndarray:
    """HF `sinusoids` (modeling_whisper.py) evaluated in f64 and rounded once to f32."""
    half = channels // 2
    inc = math.log(10000.0) / (half - 1)
    inv = np.exp(-inc * np.arange(half, dtype=np.float64))
    t = np.arange(length, dtype=np.float64)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)


def generate(cfg: ModelCfg, seed: int = 0) -> dict[str, np.ndarray]:
    """name -> f32 array, identical (bitwise, except <=1 ulp on the sinusoid table) to what
    `wb_create(..., weights_path=NULL)` builds on the device."""
    out = {}
    for tid, (name, shape, kind, std, off) in enumerate(tensor_specs(cfg)):
        n = int(np.prod(shape))
        if kind == KIND_SINUSOID:
            out[name] = sinusoid_table(shape[0], shape[1])
            continue
        u = hash_u(seed, tid, n).astype(np.float32)
        scale = np.float32(std / IH4_SD)
        v = u * scale
        if off != 0.0:
            v = v + np.float32(off)
        out[name] = v.reshape(shape)
    return out


BLOB_MAGIC = b"WB200W01"


def save_blob(path: str, cfg: ModelCfg, tensors: dict[str, np.ndarray]) -> None:
    """`.wb200` weight blob: magic, u64 json_len, JSON index, 64-byte aligned f32 LE payload.
    The format `wb_create(weights_path=...)` mmaps; also the target of a future
    ONNX-initializer converter."""
    index, off = [], 0
    for name, shape, *_ in tensor_specs(cfg):
        a = tensors[name]
        assert tuple(a.shape) == tuple(shape), (name, a.shape, shape)
        nbytes = a.size * 4
        index.append({"name": name, "shape": list(shape), "offset": off, "nbytes": nbytes})
        off += (nbytes + 63) // 64 * 64
    meta = json.dumps({"cfg": asdict(cfg), "tensors": index}).encode()
    head = BLOB_MAGIC + struct.pack("<Q", len(meta)) + meta
    pad = (-len(head)) % 64
    with open(path, "wb") as f:
        f.write(head + b"\0" * pad)
        for ent in index:
            a = np.ascontiguousarray(tensors[ent["name"]], dtype="<f4")
            f.write(a.tobytes())
            f.write(b"\0" * ((-a.nbytes) % 64))
