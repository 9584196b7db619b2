"""oracle/whisper_ref.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy f32 restatement of the model-execution half of the reference's hot path:
`run_encoder` (/root/reference/src/main.rs:698-707), `greedy_decode_with_past` (:753-829) and
`argmax_last_dim_raw` (:709-735).  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
`--impl reference` legs of bench.py import this; the product (whisper-rust-ort_b200/) never does.

Where the arithmetic lives: the reference executes three ONNX graphs through ONNX Runtime
(`ort =2.0.0-rc.6`, Cargo.toml:22; ORT 1.19.x prebuilt) — a third-party dependency absent from
/root/reference, as are the .onnx files.  The graphs are traced from Hugging Face
`WhisperForConditionalGeneration` (scripts/export_onnx_whisper.py:20-28, optimum 2.1 /
transformers 4.42), so the published algorithm restated here is that module's forward:
pre-LN blocks, q scaled by head_dim**-0.5, k_proj without bias, exact-erf GELU, LN eps 1e-5,
fixed sinusoid encoder positions, learned decoder positions, tied output projection, causal mask
on decoder self-attention only.

Parity pinning: the reference holds no numeric golden vectors for encoder states / token ids
("parity unpinned" by the reference).  This file is pinned against
tests/golden/hf_whisper_*.npz, produced by tests/golden/make_golden.py from the installed
transformers 5.5.0 Whisper implementation on the same seeded weights and inputs.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import erf

F = np.float32


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdims=True, dtype=F)
    xc = x - mu
    var = (xc * xc).mean(-1, keepdims=True, dtype=F)
    return xc / np.sqrt(var + F(eps)) * w + b


def gelu(x):
    return (F(0.5) * x * (F(1.0) + erf(x * F(1.0 / math.sqrt(2.0))))).astype(F)


def linear(x, w, b=None):
    y = x @ w.T
    return y if b is None else y + b


def softmax(s):
    s = s - s.max(-1, keepdims=True)
    e = np.exp(s)
    return e / e.sum(-1, keepdims=True, dtype=F)


def argmax_last_dim_raw(row: np.ndarray, suppress: set[int] | None) -> int:
    """main.rs:709-735 on one logits row: strict '>' so the lowest index wins ties; suppressed
    ids skipped; NaN never wins; all-skipped -> 0."""
    v = np.array(row, dtype=F, copy=True)
    if suppress:
        idx = np.fromiter((i for i in suppress if 0 <= i < v.size), dtype=np.int64)
        v[idx] = -np.inf
    v[np.isnan(v)] = -np.inf
    if not np.any(v > -np.inf):
        return 0
    return int(np.argmax(v))        # numpy argmax returns the first maximum


class WhisperRef:
    def __init__(self, cfg, weights: dict[str, np.ndarray]):
        self.cfg = cfg
        self.w = {k: np.asarray(v, dtype=F) for k, v in weights.items()}
        self.H = cfg.n_heads
        self.hd = cfg.d_model // cfg.n_heads
        self.scale = F(self.hd ** -0.5)

    # ---------------- attention helpers ----------------
    def _split(self, x):            # [B,T,d] -> [B,H,T,hd]
        B, T, _ = x.shape
        return x.reshape(B, T, self.H, self.hd).transpose(0, 2, 1, 3)

    def _merge(self, x):            # [B,H,T,hd] -> [B,T,d]
        B, H, T, hd = x.shape
        return x.transpose(0, 2, 1, 3).reshape(B, T, H * hd)

    def _attend(self, q, k, v, causal_offset=None):
        """q [B,H,Tq,hd], k/v [B,H,Tk,hd]. causal_offset = absolute position of q row 0."""
        s = (q * self.scale) @ k.transpose(0, 1, 3, 2)
        if causal_offset is not None:
            Tq, Tk = s.shape[-2:]
            qi = causal_offset + np.arange(Tq)[:, None]
            kj = np.arange(Tk)[None, :]
            s = np.where(kj <= qi, s, F(-np.inf))
        return softmax(s) @ v

    # ---------------- encoder (run_encoder, main.rs:698-707) ----------------
    def conv_stem(self, mel):
        """mel [B,n_mels,3000] -> [B,1500,d] after conv1+GELU, conv2(stride 2)+GELU, + positions."""
        w = self.w
        e = "model.encoder."
        B, C, T = mel.shape
        xp = np.pad(mel.astype(F), ((0, 0), (0, 0), (1, 1)))
        w1 = w[e + "conv1.weight"]
        h = sum(xp[:, :, k:k + T].transpose(0, 2, 1) @ w1[:, :, k].T for k in range(3))
        h = gelu(h + w[e + "conv1.bias"])                       # [B,3000,d]
        hp = np.pad(h, ((0, 0), (1, 1), (0, 0)))
        w2 = w[e + "conv2.weight"]
        T2 = (T + 2 - 3) // 2 + 1
        h2 = sum(hp[:, k:k + 2 * T2:2, :] @ w2[:, :, k].T for k in range(3))
        h2 = gelu(h2 + w[e + "conv2.bias"])                     # [B,1500,d]
        return (h2 + w[e + "embed_positions.weight"][None, :T2]).astype(F)

    def encoder_layer(self, x, i):
        w = self.w
        p = f"model.encoder.layers.{i}."
        h = layer_norm(x, w[p + "self_attn_layer_norm.weight"], w[p + "self_attn_layer_norm.bias"])
        q = self._split(linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]))
        k = self._split(linear(h, w[p + "self_attn.k_proj.weight"]))
        v = self._split(linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"]))
        a = self._merge(self._attend(q, k, v))
        x = x + linear(a, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
        h = layer_norm(x, w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"])
        h = gelu(linear(h, w[p + "fc1.weight"], w[p + "fc1.bias"]))
        return (x + linear(h, w[p + "fc2.weight"], w[p + "fc2.bias"])).astype(F)

    def encode(self, mel, return_layers=False):
        x = self.conv_stem(mel)
        layers = [x]
        for i in range(self.cfg.enc_layers):
            x = self.encoder_layer(x, i)
            layers.append(x)
        w = self.w
        out = layer_norm(x, w["model.encoder.layer_norm.weight"], w["model.encoder.layer_norm.bias"]).astype(F)
        return (out, layers) if return_layers else out

    # ---------------- decoder ----------------
    def cross_kv(self, enc):
        """The `present.{i}.encoder.{key,value}` outputs of decoder_model.onnx (main.rs:786-787)."""
        out = []
        for i in range(self.cfg.dec_layers):
            p = f"model.decoder.layers.{i}.encoder_attn."
            k = self._split(linear(enc, self.w[p + "k_proj.weight"]))
            v = self._split(linear(enc, self.w[p + "v_proj.weight"], self.w[p + "v_proj.bias"]))
            out.append((k, v))
        return out

    def decoder_forward(self, ids, pos0, self_kv, cross):
        """ids [B,T] at absolute positions pos0..pos0+T-1; self_kv: list of [k,v] (or None) per
        layer, extended in place.  Returns logits of the LAST row only, [B,vocab]."""
        w = self.w
        d = "model.decoder."
        B, T = ids.shape
        x = (w[d + "embed_tokens.weight"][ids] + w[d + "embed_positions.weight"][pos0:pos0 + T][None]).astype(F)
        for i in range(self.cfg.dec_layers):
            p = f"{d}layers.{i}."
            h = layer_norm(x, w[p + "self_attn_layer_norm.weight"], w[p + "self_attn_layer_norm.bias"])
            q = self._split(linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]))
            k = self._split(linear(h, w[p + "self_attn.k_proj.weight"]))
            v = self._split(linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"]))
            if self_kv[i] is not None:
                k = np.concatenate([self_kv[i][0], k], axis=2)
                v = np.concatenate([self_kv[i][1], v], axis=2)
            self_kv[i] = [k, v]
            a = self._merge(self._attend(q, k, v, causal_offset=pos0))
            x = x + linear(a, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
            h = layer_norm(x, w[p + "encoder_attn_layer_norm.weight"], w[p + "encoder_attn_layer_norm.bias"])
            q = self._split(linear(h, w[p + "encoder_attn.q_proj.weight"], w[p + "encoder_attn.q_proj.bias"]))
            a = self._merge(self._attend(q, cross[i][0], cross[i][1]))
            x = x + linear(a, w[p + "encoder_attn.out_proj.weight"], w[p + "encoder_attn.out_proj.bias"])
            h = layer_norm(x, w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"])
            h = gelu(linear(h, w[p + "fc1.weight"], w[p + "fc1.bias"]))
            x = (x + linear(h, w[p + "fc2.weight"], w[p + "fc2.bias"])).astype(F)
        hl = layer_norm(x[:, -1], w[d + "layer_norm.weight"], w[d + "layer_norm.bias"]).astype(F)
        return hl @ w[d + "embed_tokens.weight"].T           # tied proj_out, no bias

    def greedy(self, enc, prompt, max_new_tokens, eot, suppress=(), begin_suppress=(),
               return_logits=False, forced=None):
        """greedy_decode_with_past (main.rs:753-829) for a batch of independent sequences.
        Returns a list of per-sequence token lists = prompt + generated (EOT included if hit).
        `forced` [B, n] teacher-forces the generated ids (logits still returned) — used to
        compare a reduced-precision build step by step."""
        B = enc.shape[0]
        base = set(int(t) for t in suppress)
        first = base | set(int(t) for t in begin_suppress)
        cross = self.cross_kv(enc)
        self_kv = [None] * self.cfg.dec_layers
        toks = [list(map(int, prompt)) for _ in range(B)]
        done = [False] * B
        all_logits = []
        ids = np.tile(np.asarray(prompt, dtype=np.int64)[None], (B, 1))
        logits = self.decoder_forward(ids, 0, self_kv, cross)               # step 0 (:771-779)
        pos = len(prompt)
        step = 0
        while True:
            if return_logits:
                all_logits.append(logits.copy())
            nxt = np.zeros(B, dtype=np.int64)
            for b in range(B):
                t = argmax_last_dim_raw(logits[b], first if step == 0 else base)
                if forced is not None:
                    t = int(forced[b][step])
                nxt[b] = t
                if not done[b]:
                    toks[b].append(t)
                    if t == eot:
                        done[b] = True                                       # :781-783, :820-822
            step += 1
            # `for _ in 1..max_new_tokens` (:793): at most max(1,max_new) generated tokens
            if step >= max_new_tokens or all(done):
                break
            logits = self.decoder_forward(nxt[:, None], pos, self_kv, cross)
            pos += 1
        return (toks, all_logits) if return_logits else toks


def transcribe_tokens(model: WhisperRef, mel_chunks, prompt, max_new_tokens, eot, suppress=(),
                      begin_suppress=()):
    """Per-chunk body of transcribe_longform_chunked (main.rs:895-915): encoder + greedy."""
    enc = model.encode(np.asarray(mel_chunks, dtype=F))
    return model.greedy(enc, prompt, max_new_tokens, eot, suppress, begin_suppress)
