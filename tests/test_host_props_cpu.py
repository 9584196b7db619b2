"""CPU: randomised (hypothesis) agreement of the C++ host functions with oracle/host_ref.py — the hand-picked
cases of test_host_cpu.py widened to arbitrary word lists, sample counts and rate pairs, plus the structural
properties the reference relies on (chunk windows cover the file, stitching never duplicates an overlap)."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import host_ref as hr
from test_host_cpu import L, stitch            # noqa: F401  (fixture + helper)

WORDS = ["the", "The", "cat", "sat", "on", "mat", "über", "naïve", "日本", "a", "A", "and", "then", "went", "home", "x1", "..."]
words = st.lists(st.sampled_from(WORDS), min_size=0, max_size=24)


def _join(ws, rng_spaces):
    return "".join(w + " " * s for w, s in zip(ws, rng_spaces)) if ws else ""


@settings(max_examples=150, deadline=None)
@given(st.lists(words, min_size=0, max_size=6), st.integers(0, 2**31 - 1))
def test_stitching_matches_oracle_on_random_word_lists(L, chunks_w, seed):
    rng = np.random.default_rng(seed)
    chunks = []
    for i, ws in enumerate(chunks_w):
        if i and chunks_w[i - 1] and rng.random() < 0.6:          # make a real overlap with the previous chunk
            k = int(rng.integers(1, min(len(chunks_w[i - 1]), 18) + 1))
            ws = chunks_w[i - 1][-k:] + ws
        chunks.append(_join(ws, rng.integers(1, 3, len(ws))))
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
