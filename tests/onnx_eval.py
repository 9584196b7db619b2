"""Test infrastructure: evaluates an ONNX graph node by node in numpy (f32) — the arithmetic ONNX Runtime's CPU provider
performs on the reference's export, for the operator set a torch.onnx export of Whisper uses (Conv, MatMul, Add/Sub/Mul/Div,
Pow, Sqrt, Erf, ReduceMean, Softmax, Transpose, Reshape, Gather, Trilu, Constant, Identity).  No `onnx` / `onnxruntime`
package exists offline, so the protobuf wire format is read directly.  Used to hold oracle/whisper_ref.py to the exported
GRAPH (what `ort::Session::run` executes at main.rs:703 / :770 / :806), not only to HF's eager forward."""
import math
import struct

import numpy as np


def _varint(b, i):
    v = s = 0
    while True:
        c = b[i]
        i += 1
        v |= (c & 0x7F) << s
        s += 7
        if c < 0x80:
            return v, i


def _fields(b):
    i, out = 0, []
    while i < len(b):
        t, i = _varint(b, i)
        f, w = t >> 3, t & 7
        if w == 0:
            v, i = _varint(b, i)
        elif w == 2:
            n, i = _varint(b, i)
            v = b[i:i + n]
            i += n
        elif w == 5:
            v = b[i:i + 4]
            i += 4
        elif w == 1:
            v = b[i:i + 8]
            i += 8
        else:
            raise ValueError("wire type %d" % w)
        out.append((f, w, v))
    return out


def _signed(v):
    return v - (1 << 64) if v >= 1 << 63 else v


def _ints(fs, field):
    out = []
    for f, w, v in fs:
        if f != field:
            continue
        if w == 0:
            out.append(_signed(v))
        else:                                            # packed
            i = 0
            while i < len(v):
                x, i = _varint(v, i)
                out.append(_signed(x))
    return out


def _tensor(b):
    fs = _fields(b)
    dims = _ints(fs, 1)
    dtype = [v for f, w, v in fs if f == 2][0]
    name = b"".join(v for f, w, v in fs if f == 8).decode()
    raw = b"".join(v for f, w, v in fs if f == 9)
    np_t = {1: np.float32, 7: np.int64, 6: np.int32, 9: np.bool_, 11: np.float64, 10: np.float16}[dtype]
    if raw:
        a = np.frombuffer(raw, np_t)
    elif dtype == 1:
        a = np.array([struct.unpack("<f", v)[0] for f, w, v in fs if f == 4 and w == 5] or
                     np.frombuffer(b"".join(v for f, w, v in fs if f == 4 and w == 2), np.float32), np.float32)
    else:
        a = np.array(_ints(fs, 7), np_t)
    return name, a.reshape(dims).copy()


def _attrs(node_fields):
    out = {}
    for f, w, v in node_fields:
        if f != 5:
            continue
        a = _fields(v)
        name = [x for ff, ww, x in a if ff == 1][0].decode()
        if any(ff == 5 for ff, ww, x in a):
            out[name] = _tensor([x for ff, ww, x in a if ff == 5][0])[1]
        elif any(ff == 8 for ff, ww, x in a):
            out[name] = _ints(a, 8)
        elif any(ff == 3 for ff, ww, x in a):
            out[name] = _ints(a, 3)[0]
        elif any(ff == 2 for ff, ww, x in a):
            out[name] = struct.unpack("<f", [x for ff, ww, x in a if ff == 2][0])[0]
        elif any(ff == 4 for ff, ww, x in a):
            out[name] = [x for ff, ww, x in a if ff == 4][0].decode()
    return out


_erf = np.vectorize(math.erf, otypes=[np.float64])


def _conv1d(x, w, b, a):
    assert a.get("group", 1) == 1 and a.get("dilations", [1]) == [1]
    (stride,), pads = a.get("strides", [1]), a.get("pads", [0, 0])
    x = np.pad(x, ((0, 0), (0, 0), (pads[0], pads[1])))
    k = w.shape[2]
    n_out = (x.shape[2] - k) // stride + 1
    cols = np.stack([x[:, :, j:j + stride * n_out:stride] for j in range(k)], axis=3)         # [B, Cin, n_out, k]
    y = np.einsum("bclk,ock->bol", cols, w, optimize=True).astype(np.float32)
    return y + b[None, :, None] if b is not None else y


def run(path, feeds):
    """-> {graph output name: array}."""
    model = _fields(open(path, "rb").read())
    g = _fields([v for f, w, v in model if f == 7][0])
    env = dict(feeds)
    for f, w, v in g:
        if f == 5:
            name, a = _tensor(v)
            env[name] = a
    outputs = [b"".join(x for ff, ww, x in _fields(v) if ff == 1).decode() for f, w, v in g if f == 12]
    for f, w, v in g:
        if f != 1:
            continue
        n = _fields(v)
        op = [x for ff, ww, x in n if ff == 4][0].decode()
        ins = [env[x.decode()] if x else None for ff, ww, x in n if ff == 1]
        outs = [x.decode() for ff, ww, x in n if ff == 2]
        a = _attrs(n)
        if op == "Constant":
            r = a["value"]
        elif op == "Identity":
            r = ins[0]
        elif op in ("Add", "Sub", "Mul", "Div"):
            r = {"Add": np.add, "Sub": np.subtract, "Mul": np.multiply, "Div": np.divide}[op](ins[0], ins[1])
        elif op == "Pow":
            r = np.power(ins[0], ins[1]).astype(ins[0].dtype)
        elif op == "Sqrt":
            r = np.sqrt(ins[0])
        elif op == "Erf":
            r = _erf(ins[0]).astype(np.float32)
        elif op == "MatMul":
            r = np.matmul(ins[0], ins[1])
        elif op == "ReduceMean":
            r = np.mean(ins[0], axis=tuple(a["axes"]), keepdims=bool(a.get("keepdims", 1)), dtype=np.float32)
        elif op == "Softmax":
            z = ins[0] - np.max(ins[0], axis=a.get("axis", -1), keepdims=True)
            e = np.exp(z)
            r = e / np.sum(e, axis=a.get("axis", -1), keepdims=True)
        elif op == "Transpose":
            r = np.transpose(ins[0], a["perm"])
        elif op == "Reshape":
            shape = [int(s) if s != 0 else ins[0].shape[i] for i, s in enumerate(ins[1])]
            r = ins[0].reshape(shape)
        elif op == "Gather":
            r = np.take(ins[0], ins[1], axis=a.get("axis", 0))
        elif op == "Trilu":
            k = int(ins[1]) if len(ins) > 1 and ins[1] is not None else 0
            r = np.triu(ins[0], k) if a.get("upper", 1) else np.tril(ins[0], k)
        elif op == "Conv":
            r = _conv1d(ins[0], ins[1], ins[2] if len(ins) > 2 else None, a)
        else:
            raise NotImplementedError("ONNX op %s" % op)
        env[outs[0]] = np.asarray(r, np.float32) if np.asarray(r).dtype == np.float64 else np.asarray(r)
    return {o: env[o] for o in outputs}
