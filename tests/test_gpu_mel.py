"""GPU parity of kernel group 1 (log-mel) against the C restatement of main.rs:323-509 and the
committed HF golden vectors.  Tolerance: 1e-4 absolute (BASELINE.json north_star)."""
import numpy as np
import pytest

import mel_oracle as mo

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def model(wb):
    m = wb.Whisper(wb.default_cfg("toy", max_batch=4, max_chunks=64))
    yield m
    m.close()


def test_exact_30s_clips_vs_oracle_and_hf_golden(wb, model, golden_dir):
    x = wb.synth.batch(3, seed=0)
    mels, n_chunks = model.log_mel(x)
    assert n_chunks == 3
    g = np.load(f"{golden_dir}/mel_hf_seed0.npz")
    for i in range(3):
        ref = mo.log_mel(x[i])
        assert mels[i].shape == (80, 3000)
        assert np.abs(mels[i] - ref).max() <= TOL
        assert np.abs(mels[i][:, g["frames"]] - g["mel"][i]).max() <= TOL
        assert np.abs(mels[i][:, -4:] - g["edge"][i]).max() <= TOL
    # the resident chunk batch is the same data (one chunk per exact-30 s clip)
    cm = model.chunk_mel(0, 3)
    assert np.abs(cm - np.stack(mels)).max() == 0.0


@pytest.mark.parametrize("n", [1, 2, 159, 160, 161, 399, 400, 401, 5000, 16000 * 7 + 13, 480001])
def test_ragged_lengths(wb, model, n):
    x = wb.synth.clip(5, seed=3, seconds=31.0)[:n]
    mels, n_chunks = model.log_mel([x])
    ref = mo.log_mel(x)
    assert mels[0].shape == ref.shape
    assert np.abs(mels[0] - ref).max() <= TOL
    assert n_chunks == len(mo.chunk_starts(n))


def test_long_file_global_max_and_chunking(wb, model):
    # 70 s file: quiet first half, loud second half -> the clamp floor of chunk 0 is set by chunk 2
    x = np.concatenate([0.001 * wb.synth.clip(1, 7, 35.0), wb.synth.clip(2, 7, 35.0)])
    y = wb.synth.clip(3, 7, 12.5)
    mels, n_chunks = model.log_mel([x, y])
    rx, ry = mo.log_mel(x), mo.log_mel(y)
    assert np.abs(mels[0] - rx).max() <= TOL and np.abs(mels[1] - ry).max() <= TOL
    cx, cy = mo.chunk_mels(rx, len(x)), mo.chunk_mels(ry, len(y))
    assert n_chunks == len(cx) + len(cy) == 4
    fi, sp = model.chunks(n_chunks)
    assert fi.tolist() == [0, 0, 0, 1] and sp.tolist() == [0, 400000, 800000, 0]
    got = model.chunk_mel(0, n_chunks)
    ref = np.concatenate([cx, cy])
    assert np.abs(got - ref).max() <= TOL
    # zero padding is literal 0.0 in mel space (quirk Q2), not the clamp floor
    assert np.all(got[3][:, 1250:] == 0.0)


def test_empty_audio_is_an_error(wb, model):
    with pytest.raises(wb.WbError, match="Empty audio"):
        model.log_mel([np.zeros(0, np.float32)])


def test_batch_of_many_clips_is_order_independent(wb, model):
    x = wb.synth.batch(6, seed=11, seconds=3.0)
    a, _ = model.log_mel(x)
    b, _ = model.log_mel(x[::-1].copy())
    for i in range(6):
        assert np.array_equal(a[i], b[5 - i])


def test_128_bin_frontend_large_v3(wb, golden_dir):
    """BASELINE.json configs[4]: the same kernels with the 128-triangle filterbank (toy widths behind it, so only the
    frontend differs) against the oracle (<= 1e-4) and the HF large-v3 feature extractor golden (<= 2e-4, see
    tests/test_oracle_cpu.py for the one bin that needs the slack); ragged and chunked files as for 80 bins."""
    cfg = wb.default_cfg("toy", max_batch=4, max_chunks=16)
    cfg.n_mels = 128
    m = wb.Whisper(cfg)
    x = wb.synth.batch(3, seed=0)
    mels, n_chunks = m.log_mel(x)
    assert n_chunks == 3
    g = np.load(f"{golden_dir}/mel_hf128_seed0.npz")
    for i in range(3):
        assert mels[i].shape == (128, 3000)
        assert np.abs(mels[i] - mo.log_mel(x[i], n_mels=128)).max() <= TOL
        assert np.abs(mels[i][:, g["frames"]] - g["mel"][i]).max() <= 2e-4
    assert np.abs(m.chunk_mel(0, 3) - np.stack(mels)).max() == 0.0
    long = np.concatenate([0.001 * wb.synth.clip(1, 7, 35.0), wb.synth.clip(2, 7, 35.0)])
    odd = wb.synth.clip(5, seed=3, seconds=8.0)[:16000 * 7 + 13]
    mels, n_chunks = m.log_mel([long, odd, odd[:1], odd[:401]])
    refs = [mo.log_mel(c, n_mels=128) for c in (long, odd, odd[:1], odd[:401])]
    for a, r in zip(mels, refs):
        assert a.shape == r.shape and np.abs(a - r).max() <= TOL
    ref = np.concatenate([mo.chunk_mels(r, n) for r, n in zip(refs, (len(long), len(odd), 1, 401))])
    assert n_chunks == len(ref) == 6
    assert np.abs(m.chunk_mel(0, n_chunks) - ref).max() <= TOL
    m.close()


@pytest.mark.parametrize("packed", ["0", "1"])
@pytest.mark.parametrize("ctas", ["1", "3"])
def test_kernel_variants_agree(wb, model, monkeypatch, packed, ctas):
    """FADD2 butterflies or scalar ones; a grid of 1 x or 3 x the SM count of persistent CTAs (so a CTA walks several
    tiles, of different files, with the next tile's PCM streaming in under the current one): same log-mel within the
    tolerance.  The second file is 4-byte- but not 16-byte-aligned in the packed PCM buffer (the cp.async 4-byte path)."""
    monkeypatch.setenv("WB_MEL_PACKED", packed)
    monkeypatch.setenv("WB_MEL_CTAS_PER_SM", ctas)
    x = [wb.synth.clip(1, 9, 20.0)[:300001], wb.synth.clip(2, 9, 25.0), wb.synth.clip(3, 9, 1.0)[:7777], wb.synth.clip(4, 9, 30.0)]
    mels, _ = model.log_mel(x)
    for a, c in zip(mels, x):
        assert np.abs(a - mo.log_mel(c)).max() <= TOL
