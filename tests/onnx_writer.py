"""Synthesises ONNX files with the structure of an optimum Whisper export, without the `onnx`
package (hand-rolled protobuf wire format): named initializers for everything an op consumes
directly, every nn.Linear weight as an anonymous TRANSPOSED `onnx::MatMul_<n>` initializer, and
MatMul/Add nodes in the order the Hugging Face forward executes them.  Test input for
csrc/host/onnx.cpp (no real .onnx file exists offline)."""
from __future__ import annotations

import struct

import numpy as np


def _varint(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field(num: int, wt: int, payload: bytes) -> bytes:
    return _varint((num << 3) | wt) + payload


def _ld(num: int, payload: bytes) -> bytes:
    return _field(num, 2, _varint(len(payload)) + payload)


def tensor(name: str, a: np.ndarray, packed_dims=True, use_float_data=False, dtype="f32") -> bytes:
    a = np.ascontiguousarray(a)
    body = b""
    if packed_dims:
        body += _ld(1, b"".join(_varint(int(d)) for d in a.shape))
    else:
        body += b"".join(_field(1, 0, _varint(int(d))) for d in a.shape)
    if dtype == "f32":
        body += _field(2, 0, _varint(1))
        data = a.astype("<f4").tobytes()
        body += _ld(4, data) if use_float_data else b""
    elif dtype == "f16":
        body += _field(2, 0, _varint(10))
        data = a.astype("<f2").tobytes()
    else:
        raise ValueError(dtype)
    body += _ld(8, name.encode())
    if not use_float_data:
        body += _ld(9, data)
    return body


def node(op: str, inputs, outputs, name="") -> bytes:
    body = b"".join(_ld(1, i.encode()) for i in inputs) + b"".join(_ld(2, o.encode()) for o in outputs)
    if name:
        body += _ld(3, name.encode())
    return body + _ld(4, op.encode())


def model(nodes, initializers) -> bytes:
    graph = b"".join(_ld(1, n) for n in nodes) + _ld(2, b"main_graph") + b"".join(_ld(5, t) for t in initializers)
    return _field(1, 0, _varint(8)) + _ld(2, b"pytorch") + _ld(7, graph) + _ld(8, _field(2, 0, _varint(17)))   # ir_version, producer, graph, opset


def export_like_optimum(cfg, W: dict, out_dir: str, prefix_enc="", prefix_dec="model.decoder.", f16=False, reorder=None):
    """W: HF state_dict-style name -> f32 array (weights.generate).  Writes encoder_model.onnx and
    decoder_model.onnx into out_dir.  reorder(list of linear keys) -> the order their MatMul nodes are emitted in
    (default: the order the Hugging Face forward executes them)."""
    dt = "f16" if f16 else "f32"
    counter = [100]

    def build(hf_prefix, name_prefix, named, linears, tied=None):
        inits, nodes = [], []
        for i, key in enumerate(named):
            inits.append(tensor(name_prefix + key, W[hf_prefix + key], packed_dims=(i % 2 == 0), use_float_data=(i % 5 == 0 and not f16), dtype=dt))
        cur = "x0"
        nodes.append(node("Identity", ["input"], [cur]))
        for key in (reorder(list(linears)) if reorder else linears):
            counter[0] += 7
            an = f"onnx::MatMul_{counter[0]}"
            inits.append(tensor(an, W[hf_prefix + key + ".weight"].T, dtype=dt))         # [in, out]
            nxt = f"t{counter[0]}"
            nodes.append(node("MatMul", [cur, an], [nxt], name="/" + key.replace(".", "/") + "/MatMul"))
            if hf_prefix + key + ".bias" in W:
                nodes.append(node("Add", [name_prefix + key + ".bias", nxt], [nxt + "b"]))
                nxt += "b"
            nodes.append(node("Softmax", [nxt], [nxt + "s"]))                             # unrelated op in between
            cur = nxt + "s"
        if tied is not None:
            counter[0] += 7
            an = f"onnx::MatMul_{counter[0]}"
            inits.append(tensor(an, W[tied].T, dtype=dt))
            nodes.append(node("MatMul", [cur, an], ["logits"]))
        return model(nodes, inits)

    e_named = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "embed_positions.weight"]
    e_lin = []
    for i in range(cfg.enc_layers):
        L = f"layers.{i}."
        e_named += [L + n + s for n in ("self_attn_layer_norm", "final_layer_norm") for s in (".weight", ".bias")]
        e_named += [L + n + ".bias" for n in ("self_attn.q_proj", "self_attn.v_proj", "self_attn.out_proj", "fc1", "fc2")]
        e_lin += [L + n for n in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.out_proj", "fc1", "fc2")]
    e_named += ["layer_norm.weight", "layer_norm.bias"]
    open(f"{out_dir}/encoder_model.onnx", "wb").write(build("model.encoder.", prefix_enc, e_named, e_lin))

    d_named = ["embed_tokens.weight", "embed_positions.weight"]
    d_lin = []
    for i in range(cfg.dec_layers):
        L = f"layers.{i}."
        d_named += [L + n + s for n in ("self_attn_layer_norm", "encoder_attn_layer_norm", "final_layer_norm") for s in (".weight", ".bias")]
        d_named += [L + a + "." + n + ".bias" for a in ("self_attn", "encoder_attn") for n in ("q_proj", "v_proj", "out_proj")]
        d_named += [L + "fc1.bias", L + "fc2.bias"]
        d_lin += [L + a + "." + n for a in ("self_attn", "encoder_attn") for n in ("q_proj", "k_proj", "v_proj", "out_proj")]
        d_lin += [L + "fc1", L + "fc2"]
    d_named += ["layer_norm.weight", "layer_norm.bias"]
    open(f"{out_dir}/decoder_model.onnx", "wb").write(build("model.decoder.", prefix_dec, d_named, d_lin, tied="model.decoder.embed_tokens.weight"))
