"""CPU: the MPEG Layer III ingest (csrc/host/mp3.cpp) behind wb_host_load_audio_16k_mono — the compressed
container the reference reads through symphonia (main.rs:266-275, Cargo.toml:19).

No MP3 file, encoder or symphonia exists offline, so parity is anchored three ways:
 * an INDEPENDENT conforming decoder — libavcodec's `mp3float`, found inside the opencv wheel of this image
   (tests/libav_ref.py) — must produce the same PCM from well-formed streams drawn by tests/mp3_writer.py
   over MPEG-1 / -2 / -2.5, every sample rate, mono / stereo / MS and intensity stereo, all four window types and mixed
   blocks, every Huffman book, the bit reservoir, CRC;   tolerance 1e-4 of the stream's peak + 1e-6 (two f32 pipelines);
 * a closed-form case: one spectral line is a narrow-band tone at (k + 1/2) * sr / 1152;
 * the constant tables are checked for what the standard guarantees (complete prefix codes, band sums).
The container behaviour the reference's loop can see of symphonia (ID3v2 skipped, Xing/Info frame not decoded, no
gapless trimming, truncated last frame dropped, Layer I/II and free format rejected) is pinned here as well."""
import ctypes as C
import itertools

import numpy as np
import pytest

import host_ref as hr
import libav_ref
import mp3_writer as mw

needs_libav = pytest.mark.skipif(not libav_ref.available(), reason="no libavcodec with an MP3 decoder in this image")
f32p = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def L(wb):
    L = wb.lib()
    L.wb_host_mp3_decode_mono.argtypes = [C.c_char_p, C.c_int64, C.POINTER(f32p), C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_uint32)]
    L.wb_host_load_audio_16k_mono.argtypes = [C.c_char_p, C.POINTER(f32p), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.wb_host_free.argtypes = [C.c_void_p]
    L.wb_last_error.restype = C.c_char_p
    return L


def decode(L, data):
    out, n, ch, sr = f32p(), C.c_int64(), C.c_int(), C.c_uint32()
    if L.wb_host_mp3_decode_mono(data, len(data), C.byref(out), C.byref(n), C.byref(ch), C.byref(sr)) != 0:
        raise RuntimeError(L.wb_last_error().decode())
    a = np.ctypeslib.as_array(out, (max(n.value, 1),))[:n.value].copy()
    L.wb_host_free(out)
    return a, ch.value, sr.value


def load(L, path):
    out, n, dur = f32p(), C.c_int64(), C.c_double()
    if L.wb_host_load_audio_16k_mono(str(path).encode(), C.byref(out), C.byref(n), C.byref(dur)) != 0:
        raise RuntimeError(L.wb_last_error().decode())
    a = np.ctypeslib.as_array(out, (max(n.value, 1),))[:n.value].copy()
    L.wb_host_free(out)
    return a, dur.value


def mono_mix(planes):
    if len(planes) == 1:
        return planes[0]
    return (np.float32(0) + planes[0] + planes[1]) / np.float32(2)             # main.rs:268-273: acc over channels, / channels, all f32


def test_tables_are_what_the_standard_guarantees():
    T = mw.tables()
    for t, ((dim, enc), linbits) in T["sel"].items():
        assert len(enc) == dim * dim and sum(2.0 ** -ln for _, ln in enc.values()) == 1.0, t       # complete prefix code
        codes = sorted((format(c, "0%db" % ln) for c, ln in enc.values()))
        assert all(not b.startswith(a) for a, b in zip(codes, codes[1:])), t
    assert all(sum(r) == 576 for r in T["long"]) and all(sum(r) == 192 for r in T["short"])
    assert sum(2.0 ** -ln for _, ln in T["quad_a"]) == 1.0


RATES = [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (1, 2), (2, 0), (2, 1), (2, 2)]
@needs_libav
def test_committed_tables_are_what_the_generator_writes(tmp_path):
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_mp3_tables", os.path.join(root, "tools", "gen_mp3_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gen.main(str(tmp_path / "t.h"))                         # runs the generator's own checks (Kraft sums, symbol sets, band sums)
    assert (tmp_path / "t.h").read_text() == open(gen.OUT).read()


CASES = [(v, s, ch, ms, False, res) for (v, s), ch, ms, res in itertools.product(RATES, [1, 2], [False, True], [True, False]) if not (ms and ch == 1)]
CASES += [(v, s, 2, ms, True, True) for (v, s), ms in itertools.product(RATES, [False, True])]          # intensity stereo, alone and with MS


@needs_libav
@pytest.mark.parametrize("version,sr_idx,channels,ms,intensity,reservoir", CASES)
def test_same_pcm_as_libavcodec(L, version, sr_idx, channels, ms, intensity, reservoir):
    for seed in range(3 if intensity else 2):
        frames, sr = mw.make_stream(seed=1000 * seed + 100 * channels + 10 * version + sr_idx, version=version, sr_idx=sr_idx,
                                    br_idx=9 if version == 0 else 8, channels=channels, n_frames=7, ms=ms, reservoir=reservoir,
                                    crc=bool(seed), intensity=intensity)
        ref = mono_mix(libav_ref.decode_frames(frames, channels))
        got, ch, rate = decode(L, b"".join(frames))
        assert (ch, rate, len(got)) == (channels, sr, len(ref)) and len(got) == 7 * (1152 if version == 0 else 576)
        peak = float(np.abs(ref).max())
        assert peak > 0 and np.abs(got - ref).max() <= 1e-4 * peak + 1e-6


@needs_libav
@pytest.mark.parametrize("version,sr_idx", [(0, 0), (0, 1), (1, 2), (2, 2)])
def test_file_through_the_loader(wb, L, tmp_path, version, sr_idx):
    """ID3v2 in front, an Info frame first, an ID3v1 tag and half a frame behind: the loader returns the channel mean of the
    audio frames only, resampled like main.rs:207-226."""
    channels = 2
    frames, sr = mw.make_stream(seed=7 + version, version=version, sr_idx=sr_idx, br_idx=9 if version == 0 else 8, channels=channels, n_frames=9)
    side = (32 if version == 0 else 17)
    info = bytearray(frames[0])                                              # same header, body replaced by an Info tag
    info[4:] = bytes(len(info) - 4)
    info[4 + side:4 + side + 4] = b"Info"
    id3 = b"ID3\x04\x00\x00" + bytes([0, 0, 1, 10]) + bytes(138)             # synch-safe size 138
    blob = id3 + bytes(info) + b"".join(frames) + frames[0][:len(frames[0]) // 2]
    p = tmp_path / "clip.mp3"
    p.write_bytes(blob)
    got, dur = load(L, p)
    ref = hr.resample_linear(mono_mix(libav_ref.decode_frames(frames, channels)), sr, 16000)
    assert got.shape == ref.shape and dur == len(ref) / 16000.0
    assert np.abs(got - ref).max() <= 1e-4 * float(np.abs(ref).max()) + 1e-7
    # ... and the same bytes under a .wav name: the container is recognised by content, like symphonia's probe
    q = tmp_path / "misnamed.wav"
    q.write_bytes(blob)
    assert np.array_equal(load(L, q)[0], got)
    pcm, seconds = wb.load_audio(p)                                          # the Python mirror of load_audio_16k_mono
    assert np.array_equal(pcm, got) and seconds == dur


@pytest.mark.parametrize("line", [7, 25, 40, 300])
def test_one_spectral_line_is_a_narrow_band_tone(L, line):
    n_frames = 12
    spectrum = [0] * 576
    spectrum[line] = 9
    spec = dict(global_gain=170, big_values=(line + 2) // 2, table_select=[13, 13, 13], **{"is": spectrum})
    specs = {(f, g, 0): spec for f in range(n_frames) for g in range(2)}
    frames, sr = mw.make_stream(seed=0, version=0, sr_idx=0, br_idx=9, channels=1, n_frames=n_frames, specs=specs, reservoir=False)
    pcm, ch, rate = decode(L, b"".join(frames))
    assert (ch, rate, len(pcm)) == (1, 44100, 1152 * n_frames)
    x = pcm[2304:2304 + 8 * 1152].astype(np.float64)
    power = np.abs(np.fft.rfft(x * np.hanning(len(x)))) ** 2
    f = np.fft.rfftfreq(len(x), 1.0 / sr)
    centre = (line + 0.5) * sr / 1152.0
    assert abs(f[np.argmax(power)] - centre) <= sr / 1152.0
    assert power[np.abs(f - centre) <= 2.5 * sr / 1152.0].sum() >= 0.9 * power.sum()
    assert 9 ** (4 / 3) * 2 ** ((170 - 210) / 4) * 0.2 <= np.abs(x).max() <= 9 ** (4 / 3) * 2 ** ((170 - 210) / 4) * 2.0


def test_container_rules(L, tmp_path):
    frames, sr = mw.make_stream(seed=3, channels=1, n_frames=4)
    whole, _, _ = decode(L, b"".join(frames))
    assert len(whole) == 4 * 1152
    # a frame cut off by the end of the file is dropped (IoError -> break, main.rs:258-262); junk between frames is skipped
    cut, _, _ = decode(L, b"".join(frames)[:-10])
    assert np.array_equal(cut, whole[:3 * 1152])
    junk, _, _ = decode(L, frames[0] + frames[1] + b"\x00\x01\x02" * 11 + frames[2] + frames[3] + b"TAG" + bytes(125))
    assert np.array_equal(junk, whole)
    # Layer II, free format: errors; no MPEG stream at all: the loader's "unsupported audio container"
    l2 = bytearray(frames[0] + frames[0]); l2[1] = (l2[1] & ~0x06) | 0x04; l2[len(frames[0]) + 1] = l2[1]
    with pytest.raises(RuntimeError, match="no MPEG audio stream|unsupported codec"):
        decode(L, bytes(l2))
    with pytest.raises(RuntimeError, match="no MPEG audio stream"):
        decode(L, bytes(4000))
    (tmp_path / "x.mp3").write_bytes(b"ID3" + bytes(3000))
    with pytest.raises(RuntimeError, match="unsupported audio container"):
        load(L, tmp_path / "x.mp3")


def test_damaged_streams_fail_or_decode_but_never_crash(L):
    """Bytes flipped anywhere behind the first header, and truncation: the call returns samples or an error message (the
    same sources are clean under -fsanitize=address,undefined on 400 such streams)."""
    rng = np.random.default_rng(0)
    outcomes = set()
    for it in range(150):
        v, ch = int(rng.integers(0, 3)), int(rng.integers(1, 3))
        frames, _ = mw.make_stream(seed=it % 17, version=v, sr_idx=int(rng.integers(0, 3)), br_idx=9 if v == 0 else 8, channels=ch,
                                   n_frames=3, ms=bool(ch == 2 and it % 2), intensity=bool(ch == 2 and it % 3 == 0))
        data = bytearray(b"".join(frames))
        for _ in range(int(rng.integers(1, 40))):
            data[int(rng.integers(4, len(data)))] = int(rng.integers(0, 256))
        if it % 5 == 0:
            data = data[:int(rng.integers(8, len(data)))]
        try:
            pcm, _, _ = decode(L, bytes(data))
            outcomes.add("decoded")
            assert len(pcm) % 576 == 0
        except RuntimeError as e:
            outcomes.add("error")
            assert str(e)
    assert outcomes == {"decoded", "error"}


def test_loader_is_reentrant(wb, L, tmp_path):
    """The CLI's group loader decodes the files of a group on several host threads: concurrent calls (WAV and MP3 mixed, the
    decoder's tables built on first use) return exactly what the serial calls return."""
    from concurrent.futures import ThreadPoolExecutor
    paths = []
    for i in range(6):
        frames, _ = mw.make_stream(seed=20 + i, version=i % 3, sr_idx=i % 3, br_idx=9 if i % 3 == 0 else 8, channels=1 + i % 2, n_frames=40, ms=bool(i % 2))
        p = tmp_path / f"m{i}.mp3"
        p.write_bytes(b"".join(frames))
        paths.append(p)
    for i in range(4):
        p = tmp_path / f"w{i}.wav"
        wb.synth.write_wav(str(p), wb.synth.clip(i, 2, 1.5), sr=[16000, 22050, 44100, 8000][i], fmt=["s16", "f32", "u8", "s16"][i])
        paths.append(p)
    serial = [load(L, p) for p in paths]
    with ThreadPoolExecutor(8) as ex:
        for _ in range(3):
            for (a, da), (b, db) in zip(ex.map(lambda q: load(L, q), paths * 4), serial * 4):
                assert da == db and np.array_equal(a, b)


def test_random_bytes_are_not_mistaken_for_mpeg_audio(L, tmp_path):
    """The probe wants two consecutive frame headers of one stream (like symphonia's strict first-frame read): noise, text and
    a lone header do not pass; the loader then reports the container as unsupported."""
    rng = np.random.default_rng(5)
    frames, _ = mw.make_stream(seed=2, channels=1, n_frames=2)
    blobs = [rng.integers(0, 256, 4096, dtype=np.uint8).tobytes() for _ in range(100)]
    blobs += [b"just some text, not audio at all\n" * 100, frames[0][:4] + bytes(2000), b"\xff" * 3000]
    for i, blob in enumerate(blobs):
        p = tmp_path / f"n{i}.mp3"
        p.write_bytes(blob)
        with pytest.raises(RuntimeError, match="unsupported audio container"):
            load(L, p)
