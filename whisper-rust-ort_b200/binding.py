"""ctypes binding over libwhisper_b200.so — the only way Python (tests, bench, smoke) reaches the
product: everything goes through the C ABI of include/whisper_b200.h, exactly as a Rust host
would.  Fails loudly when the library is missing or has no GPU."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .weights import ModelCfg

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwhisper_b200.so")

WB_PREC_FP32, WB_PREC_BF16 = 0, 1
N_FRAMES = 3000


class WbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libwhisper_b200 error {code}: {msg}")
        self.code = code


class wb_model_cfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_mels", "d_model", "n_heads", "ffn_dim", "enc_layers", "dec_layers",
                                         "vocab", "n_audio_ctx", "n_text_ctx", "precision", "max_batch", "max_chunks")]
    _fields_.append(("seed", C.c_uint64))


class wb_timing(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("mel_ms", "encoder_ms", "cross_kv_ms", "decode_ms", "h2d_ms", "d2h_ms")]
    _fields_ += [(n, C.c_int32) for n in ("mel_launches", "encoder_launches", "decode_launches", "decode_steps")]


_lib = None
i64p, i32p, f32p, f64p = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double)


def lib():
    """Load the shared library (no GPU needed to load; compute calls need one)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WbError(-2, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no Python/CPU fallback for the hot path)")
    L = C.CDLL(LIB_PATH)
    vp, cp, ci = C.c_void_p, C.c_char_p, C.c_int
    L.wb_last_error.restype = cp
    sig = {
        "wb_default_cfg": [C.POINTER(wb_model_cfg), cp],
        "wb_create": [C.POINTER(vp), ci, C.POINTER(wb_model_cfg), cp],
        "wb_get_cfg": [vp, C.POINTER(wb_model_cfg)],
        "wb_get_timing": [vp, C.POINTER(wb_timing)],
        "wb_set_debug": [vp, ci],
        "wb_set_load_hint": [vp, ci],
        "wb_selftest_gemm": [vp, ci, ci, ci, ci, ci, ci, f32p, f32p],
        "wb_selftest_attn": [vp, ci, f32p, f32p],
        "wb_mark": [vp, ci],
        "wb_elapsed_ms": [vp, ci, ci, f32p],
        "wb_bench_kernel": [vp, cp, ci, ci, f32p, f64p],
        "wb_device_count": [C.POINTER(ci)],
        "wb_get_tensor": [vp, cp, f32p, C.c_int64],
        "wb_log_mel": [vp, f32p, i64p, ci, C.c_int64, C.c_int64, f32p, i64p, C.POINTER(ci)],
        "wb_upload_pcm": [vp, f32p, i64p, ci, C.c_int64, C.c_int64, C.POINTER(ci)],
        "wb_run_log_mel": [vp],
        "wb_get_chunks": [vp, i32p, i64p, ci],
        "wb_get_chunk_mel": [vp, ci, ci, f32p],
        "wb_encode": [vp, f32p, ci, ci, f32p],
        "wb_get_encoder_debug": [vp, cp, f32p, C.c_int64],
        "wb_greedy_decode": [vp, ci, i64p, ci, ci, C.c_int64, i64p, ci, i64p, ci, i64p, i32p, i64p, f32p],
        "wb_transcribe_batch": [vp, f32p, i64p, ci, i64p, ci, ci, C.c_int64, i64p, ci, i64p, ci, i64p, i32p, i32p, ci, C.POINTER(ci)],
        "wb_transcribe_resident": [vp, i64p, ci, ci, C.c_int64, i64p, ci, i64p, ci, i64p, i32p, ci],
        "wb_pool_create": [C.POINTER(vp), ci, C.POINTER(wb_model_cfg), cp, ci],
        "wb_pool_slots": [vp],
        "wb_pool_submit": [vp, f32p, i64p, ci, i64p, ci, ci, C.c_int64, i64p, ci, i64p, ci, i64p, i32p, i32p, ci],
        "wb_pool_wait": [vp, ci, C.POINTER(ci)],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = ci
    L.wb_destroy.argtypes = [vp]
    L.wb_destroy.restype = None
    L.wb_pool_destroy.argtypes = [vp]
    L.wb_pool_destroy.restype = None
    _lib = L
    return L


def _chk(rc):
    if rc != 0:
        raise WbError(rc, lib().wb_last_error().decode("utf-8", "replace"))


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(f32p)


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, (a.ctypes.data_as(i64p) if a.size else None)


def default_cfg(name: str = "base", precision: int = WB_PREC_FP32, max_batch: int | None = None,
                max_chunks: int | None = None, seed: int = 0) -> wb_model_cfg:
    cfg = wb_model_cfg()
    _chk(lib().wb_default_cfg(C.byref(cfg), name.encode()))
    cfg.precision = precision
    if max_batch is not None:
        cfg.max_batch = max_batch
        cfg.max_chunks = max(cfg.max_chunks, max_batch)
    if max_chunks is not None:
        cfg.max_chunks = max_chunks
    cfg.seed = seed
    return cfg


def load_audio(path: str):
    """load_audio_16k_mono (main.rs:228-316) through the C ABI: RIFF/WAVE or MPEG Layer III file -> (mono f32 PCM at
    16 kHz, duration in seconds).  Host-only: works without a GPU."""
    L = lib()
    f32p = C.POINTER(C.c_float)
    L.wb_host_load_audio_16k_mono.argtypes = [C.c_char_p, C.POINTER(f32p), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.wb_host_free.argtypes = [C.c_void_p]
    buf, n, dur = f32p(), C.c_int64(), C.c_double()
    _chk(L.wb_host_load_audio_16k_mono(os.fspath(path).encode(), C.byref(buf), C.byref(n), C.byref(dur)))
    try:
        return np.ctypeslib.as_array(buf, (max(n.value, 1),))[:n.value].copy(), dur.value
    finally:
        L.wb_host_free(buf)


def model_cfg_of(cfg: wb_model_cfg) -> ModelCfg:
    return ModelCfg(cfg.n_mels, cfg.d_model, cfg.n_heads, cfg.ffn_dim, cfg.enc_layers, cfg.dec_layers, cfg.vocab,
                    cfg.n_audio_ctx, cfg.n_text_ctx)


class Whisper:
    """One context per GPU.  Mirrors the reference's (encoder, decoder, decoder_with_past) session
    triple built at main.rs:1103-1108."""

    def __init__(self, cfg: wb_model_cfg | None = None, device: int = 0, weights_path: str | None = None):
        self.L = lib()
        self.cfg = cfg if cfg is not None else default_cfg()
        self.h = C.c_void_p()
        _chk(self.L.wb_create(C.byref(self.h), device, C.byref(self.cfg), weights_path.encode() if weights_path else None))

    def close(self):
        if self.h:
            self.L.wb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- group 1 ----
    @staticmethod
    def _pack(clips):
        if isinstance(clips, np.ndarray) and clips.ndim == 2:
            n, ln = clips.shape
            return np.ascontiguousarray(clips, np.float32).reshape(-1), np.arange(n + 1, dtype=np.int64) * ln
        offs = np.zeros(len(clips) + 1, np.int64)
        offs[1:] = np.cumsum([len(c) for c in clips])
        flat = np.concatenate([np.asarray(c, np.float32) for c in clips]) if len(clips) else np.zeros(0, np.float32)
        return flat, offs

    def log_mel(self, clips, chunk_len: int = 0, step: int = 0, want_mel: bool = True):
        """-> (list of [n_mels, nf_i] arrays or None, n_chunks).  whisper_log_mel_80 per file (n_mels = 80; 128 for large-v3)."""
        nm = self.cfg.n_mels
        flat, offs = self._pack(clips)
        n_files = len(offs) - 1
        nfr = np.zeros(n_files, np.int64)
        nch = C.c_int(0)
        # frame count is floor(N/160) (>=1): size the host buffer first
        lens = np.diff(offs)
        nf = np.maximum(lens // 160, 1)
        out = np.empty(int(nf.sum()) * nm, np.float32) if want_mel else None
        _chk(self.L.wb_log_mel(self.h, flat.ctypes.data_as(f32p), offs.ctypes.data_as(i64p), n_files, chunk_len, step,
                               out.ctypes.data_as(f32p) if want_mel else None, nfr.ctypes.data_as(i64p), C.byref(nch)))
        assert np.array_equal(nfr, nf), (nfr, nf)
        mels = None
        if want_mel:
            mels, o = [], 0
            for k in nf:
                mels.append(out[o:o + nm * int(k)].reshape(nm, int(k)))
                o += nm * int(k)
        return mels, nch.value

    def upload_pcm(self, clips, chunk_len: int = 0, step: int = 0) -> int:
        flat, offs = self._pack(clips)
        nch = C.c_int(0)
        _chk(self.L.wb_upload_pcm(self.h, flat.ctypes.data_as(f32p), offs.ctypes.data_as(i64p), len(offs) - 1, chunk_len, step, C.byref(nch)))
        return nch.value

    def run_log_mel(self):
        _chk(self.L.wb_run_log_mel(self.h))

    def chunks(self, n):
        fi, sp = np.zeros(n, np.int32), np.zeros(n, np.int64)
        _chk(self.L.wb_get_chunks(self.h, fi.ctypes.data_as(i32p), sp.ctypes.data_as(i64p), n))
        return fi, sp

    def chunk_mel(self, begin, n):
        out = np.empty((n, self.cfg.n_mels, N_FRAMES), np.float32)
        _chk(self.L.wb_get_chunk_mel(self.h, begin, n, out.ctypes.data_as(f32p)))
        return out

    # ---- group 2 ----
    def encode(self, mel=None, chunk_begin: int = 0, B: int | None = None, want_hidden: bool = True):
        c = self.cfg
        if mel is not None:
            mel, mp = _f32(mel)
            B = mel.shape[0]
            assert mel.shape == (B, c.n_mels, N_FRAMES), mel.shape
        else:
            mp = None
        out = np.empty((B, c.n_audio_ctx, c.d_model), np.float32) if want_hidden else None
        _chk(self.L.wb_encode(self.h, mp, chunk_begin, B, out.ctypes.data_as(f32p) if want_hidden else None))
        return out

    def set_debug(self, on=True):
        _chk(self.L.wb_set_debug(self.h, int(on)))

    def set_load_hint(self, batches_in_flight: int):
        """How many batches the caller keeps in flight on this GPU (1 = latency-oriented decode kernels)."""
        _chk(self.L.wb_set_load_hint(self.h, int(batches_in_flight)))

    def encoder_debug(self, what: str, B: int):
        c = self.cfg
        out = np.empty((B, c.n_audio_ctx, c.d_model), np.float32)
        _chk(self.L.wb_get_encoder_debug(self.h, what.encode(), out.ctypes.data_as(f32p), out.size))
        return out

    # ---- group 3 ----
    def greedy_decode(self, B, prompt, max_new_tokens, eot, suppress=(), begin_suppress=(), forced=None,
                      want_logits=False):
        c = self.cfg
        prompt, pp = _i64(prompt)
        sup, sp = _i64(list(suppress))
        bsup, bp = _i64(list(begin_suppress))
        mn = max(1, max_new_tokens)
        toks = np.full((B, len(prompt) + mn), -1, np.int64)
        lens = np.zeros(B, np.int32)
        fp = None
        if forced is not None:
            forced = np.ascontiguousarray(forced, np.int64)
            assert forced.shape == (B, mn)
            fp = forced.ctypes.data_as(i64p)
        logits = np.empty((B, mn, c.vocab), np.float32) if want_logits else None
        _chk(self.L.wb_greedy_decode(self.h, B, pp, len(prompt), max_new_tokens, eot, sp, len(sup), bp, len(bsup),
                                     toks.ctypes.data_as(i64p), lens.ctypes.data_as(i32p), fp,
                                     logits.ctypes.data_as(f32p) if want_logits else None))
        seqs = [toks[b, :lens[b]].tolist() for b in range(B)]
        return (seqs, logits) if want_logits else seqs

    # ---- fused ----
    def transcribe_batch(self, clips, prompt, max_new_tokens, eot, suppress=(), begin_suppress=()):
        """Host PCM in -> per-chunk token lists + file index per chunk (H2D and D2H inside)."""
        flat, offs = self._pack(clips)
        prompt, pp = _i64(prompt)
        sup, sp = _i64(list(suppress))
        bsup, bp = _i64(list(begin_suppress))
        cap = self.cfg.max_chunks
        stride = len(prompt) + max(1, max_new_tokens)
        toks = np.full((cap, stride), -1, np.int64)
        lens, fidx = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        nch = C.c_int(0)
        _chk(self.L.wb_transcribe_batch(self.h, flat.ctypes.data_as(f32p), offs.ctypes.data_as(i64p), len(offs) - 1,
                                        pp, len(prompt), max_new_tokens, eot, sp, len(sup), bp, len(bsup),
                                        toks.ctypes.data_as(i64p), lens.ctypes.data_as(i32p), fidx.ctypes.data_as(i32p),
                                        cap, C.byref(nch)))
        n = nch.value
        return [toks[i, :lens[i]].tolist() for i in range(n)], fidx[:n].copy()

    def transcribe_resident(self, n_chunks, prompt, max_new_tokens, eot, suppress=(), begin_suppress=()):
        prompt, pp = _i64(prompt)
        sup, sp = _i64(list(suppress))
        bsup, bp = _i64(list(begin_suppress))
        stride = len(prompt) + max(1, max_new_tokens)
        toks = np.full((n_chunks, stride), -1, np.int64)
        lens = np.zeros(n_chunks, np.int32)
        _chk(self.L.wb_transcribe_resident(self.h, pp, len(prompt), max_new_tokens, eot, sp, len(sup), bp, len(bsup),
                                           toks.ctypes.data_as(i64p), lens.ctypes.data_as(i32p), n_chunks))
        return [toks[i, :lens[i]].tolist() for i in range(n_chunks)]

    def transcribe_batch_ptr(self, pcm_ptr: int, n_clips: int, clip_len: int, prompt, max_new_tokens, eot,
                             suppress=(), begin_suppress=(), out=None):
        """Same as transcribe_batch for equal-length clips living in caller memory (e.g. a pinned
        host buffer): no Python-side copy of the PCM.  Returns (tokens [n, stride], lens)."""
        offs = np.arange(n_clips + 1, dtype=np.int64) * clip_len
        prompt, pp = _i64(prompt)
        sup, sp = _i64(list(suppress))
        bsup, bp = _i64(list(begin_suppress))
        cap = self.cfg.max_chunks
        stride = len(prompt) + max(1, max_new_tokens)
        toks = np.full((cap, stride), -1, np.int64) if out is None else out
        lens, fidx = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        nch = C.c_int(0)
        _chk(self.L.wb_transcribe_batch(self.h, C.cast(pcm_ptr, f32p), offs.ctypes.data_as(i64p), n_clips,
                                        pp, len(prompt), max_new_tokens, eot, sp, len(sup), bp, len(bsup),
                                        toks.ctypes.data_as(i64p), lens.ctypes.data_as(i32p), fidx.ctypes.data_as(i32p),
                                        cap, C.byref(nch)))
        return toks[:nch.value], lens[:nch.value]

    def selftest_gemm(self, M, N, K, lda=None, batch=1, f32_out=False):
        """tcgen05 kernel vs SIMT kernel on seeded bf16 operands -> (max |diff|, max |value|)."""
        d, a = C.c_float(0), C.c_float(0)
        _chk(self.L.wb_selftest_gemm(self.h, M, N, K, lda if lda else K, batch, int(f32_out), C.byref(d), C.byref(a)))
        return float(d.value), float(a.value)

    def selftest_attn(self, B=1):
        """tcgen05 flash attention vs SIMT attention on seeded bf16 q|k|v -> (max |diff|, max |value|)."""
        d, a = C.c_float(0), C.c_float(0)
        _chk(self.L.wb_selftest_attn(self.h, B, C.byref(d), C.byref(a)))
        return float(d.value), float(a.value)

    def mark(self, slot: int):
        _chk(self.L.wb_mark(self.h, slot))

    def elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float(0)
        _chk(self.L.wb_elapsed_ms(self.h, a, b, C.byref(ms)))
        return float(ms.value)

    def bench_kernel(self, kernel: str, B: int, iters: int = 20):
        """-> (mean launch ms from CUDA events, algorithmic bytes per launch)."""
        ms, by = C.c_float(0), C.c_double(0)
        _chk(self.L.wb_bench_kernel(self.h, kernel.encode(), B, iters, C.byref(ms), C.byref(by)))
        return float(ms.value), float(by.value)

    def timing(self) -> dict:
        t = wb_timing()
        _chk(self.L.wb_get_timing(self.h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in wb_timing._fields_}

    def tensor(self, name: str, shape) -> np.ndarray:
        out = np.empty(shape, np.float32)
        _chk(self.L.wb_get_tensor(self.h, name.encode(), out.ctypes.data_as(f32p), out.size))
        return out


class Pool:
    """wb_pool: one handle, n_slots batches in flight on one GPU, driven from a single host thread.

    submit() returns a ticket at once (it blocks only while n_slots batches are already queued); wait(ticket) returns
    that batch's token lists.  The PCM of a submitted batch must stay alive until its ticket has been collected: the
    pool keeps a reference to whatever array / pointer owner was passed."""

    def __init__(self, cfg: wb_model_cfg, n_slots: int, device: int = 0, weights_path: str | None = None):
        self.L = lib()
        self.cfg = cfg
        self.h = C.c_void_p()
        self._live = {}
        _chk(self.L.wb_pool_create(C.byref(self.h), device, C.byref(self.cfg), weights_path.encode() if weights_path else None, n_slots))

    @property
    def slots(self) -> int:
        return int(self.L.wb_pool_slots(self.h))

    def submit_ptr(self, pcm_ptr: int, n_clips: int, clip_len: int, prompt, max_new_tokens, eot, suppress=(), begin_suppress=(), keep=None) -> int:
        """Equal-length clips in caller memory (e.g. a pinned host buffer); `keep` is held until the ticket is collected."""
        offs = np.arange(n_clips + 1, dtype=np.int64) * clip_len
        prompt, pp = _i64(prompt)
        sup, sp = _i64(list(suppress))
        bsup, bp = _i64(list(begin_suppress))
        cap = self.cfg.max_chunks
        stride = len(prompt) + max(1, max_new_tokens)
        toks = np.full((cap, stride), -1, np.int64)
        lens, fidx = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        t = self.L.wb_pool_submit(self.h, C.cast(pcm_ptr, f32p), offs.ctypes.data_as(i64p), n_clips, pp, len(prompt), max_new_tokens, eot,
                                  sp, len(sup), bp, len(bsup), toks.ctypes.data_as(i64p), lens.ctypes.data_as(i32p),
                                  fidx.ctypes.data_as(i32p), cap)
        if t < 0:
            _chk(t)
        self._live[t] = (keep, offs, toks, lens, fidx)
        return t

    def submit(self, clips, prompt, max_new_tokens, eot, suppress=(), begin_suppress=()) -> int:
        flat = np.ascontiguousarray(np.stack([np.asarray(c, np.float32) for c in clips]))
        return self.submit_ptr(flat.ctypes.data, flat.shape[0], flat.shape[1], prompt, max_new_tokens, eot, suppress, begin_suppress, keep=flat)

    def wait(self, ticket: int):
        """-> (token lists per chunk, file index per chunk)"""
        n = C.c_int(0)
        rc = self.L.wb_pool_wait(self.h, ticket, C.byref(n))
        _, _, toks, lens, fidx = self._live.pop(ticket, (None, None, None, None, None))
        _chk(rc)
        return [toks[i, :lens[i]].tolist() for i in range(n.value)], fidx[:n.value].copy()

    def close(self):
        if self.h:
            self.L.wb_pool_destroy(self.h)
            self.h = C.c_void_p()
            self._live.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
