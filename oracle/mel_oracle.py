"""ctypes wrapper over oracle/libwbref.so (mel_ref.c). TEST INFRASTRUCTURE — see mel_ref.c header."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    so = os.path.join(_HERE, "libwbref.so")
    src = os.path.join(_HERE, "mel_ref.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        fp = C.POINTER(C.c_float)
        L.wbref_n_frames.restype = C.c_long
        L.wbref_n_frames.argtypes = [C.c_long]
        L.wbref_log_mel.restype = C.c_int
        L.wbref_log_mel.argtypes = [fp, C.c_long, fp, C.c_int, fp, fp]
        L.wbref_log_mel_batch.restype = C.c_int
        L.wbref_log_mel_batch.argtypes = [fp, C.c_long, C.c_long, fp, C.c_int]
        L.wbref_log_mel_n.restype = C.c_int
        L.wbref_log_mel_n.argtypes = [fp, C.c_long, fp, C.c_int, C.c_int, fp, fp]
        L.wbref_mel_filterbank.argtypes = [fp]
        L.wbref_mel_filterbank_n.argtypes = [fp, C.c_int]
        L.wbref_hann.argtypes = [fp, C.c_int]
        L.wbref_fft400_f32.argtypes = [fp, fp, fp, fp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def n_frames(n: int) -> int:
    return int(lib().wbref_n_frames(n))


def log_mel(pcm: np.ndarray, dft64: bool = False, return_raw: bool = False, n_mels: int = 80):
    """whisper_log_mel_80 (main.rs:407-509) on one whole file -> [80, floor(N/160)] f32.  n_mels=128: the same function
    with the 128-triangle filterbank of the large-v3 frontend (BASELINE.json configs[4]; the reference itself only has 80)."""
    x = np.ascontiguousarray(pcm, dtype=np.float32)
    if x.size == 0:
        raise ValueError("Empty audio")          # main.rs:414-416
    nf = n_frames(x.size)
    out = np.empty((n_mels, nf), np.float32)
    raw = np.empty((n_mels, nf), np.float32)
    g = C.c_float(0)
    rc = lib().wbref_log_mel_n(_p(x), x.size, _p(out), n_mels, int(dft64), _p(raw), C.byref(g))
    assert rc == 0
    return (out, raw, float(g.value)) if return_raw else out


def log_mel_batch(pcm: np.ndarray, threads: int = 1) -> np.ndarray:
    x = np.ascontiguousarray(pcm, dtype=np.float32)
    b, n = x.shape
    out = np.empty((b, 80, n_frames(n)), np.float32)
    rc = lib().wbref_log_mel_batch(_p(x), b, n, _p(out), threads)
    assert rc == 0
    return out


def filterbank(n_mels: int = 80) -> np.ndarray:
    fb = np.empty((n_mels, 201), np.float32)
    lib().wbref_mel_filterbank_n(_p(fb), n_mels)
    return fb


def hann() -> np.ndarray:
    w = np.empty(400, np.float32)
    lib().wbref_hann(_p(w), 400)
    return w


def fft400(re: np.ndarray, im: np.ndarray | None = None):
    re = np.ascontiguousarray(re, np.float32)
    im = np.zeros(400, np.float32) if im is None else np.ascontiguousarray(im, np.float32)
    ro, io = np.empty(400, np.float32), np.empty(400, np.float32)
    lib().wbref_fft400_f32(_p(re), _p(im), _p(ro), _p(io))
    return ro, io


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


def chunk_mels(mel_full: np.ndarray, n_samples: int, chunk_len: int = 480000, step: int = 400000):
    """Chunk slicing of transcribe_longform_chunked (main.rs:895-905): [n_chunks,80,3000],
    zero (literal 0.0) padded in mel space."""
    total = mel_full.shape[1]
    nm = mel_full.shape[0]
    outs = []
    for pos in chunk_starts(n_samples, chunk_len, step):
        fs = pos // 160
        m = np.zeros((nm, 3000), np.float32)
        if fs < total:
            ae = min(fs + 3000, total)
            m[:, : ae - fs] = mel_full[:, fs:ae]
        outs.append(m)
    return np.stack(outs)
