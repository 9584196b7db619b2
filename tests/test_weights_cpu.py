"""CPU: the seeded weight definition shared by the CUDA library, the oracle and the goldens."""
import numpy as np


def test_spec_matches_whisper_base_parameter_count(wb):
    specs = wb.weights.tensor_specs(wb.weights.WHISPER_BASE)
    n = sum(int(np.prod(s[1])) for s in specs)
    assert n == 72_593_920        # HF whisper-base state_dict incl. the sinusoid table
    names = [s[0] for s in specs]
    assert len(set(names)) == len(names)
    assert not any(n.endswith("k_proj.bias") for n in names)


def test_generator_is_deterministic_and_seeded(wb):
    cfg = wb.weights.WHISPER_TOY
    a, b, c = wb.weights.generate(cfg, 0), wb.weights.generate(cfg, 0), wb.weights.generate(cfg, 1)
    k = "model.decoder.layers.1.fc1.weight"
    assert np.array_equal(a[k], b[k]) and not np.array_equal(a[k], c[k])
    assert abs(float(a[k].std()) - 0.02) < 2e-3 and abs(float(a[k].mean())) < 1e-3
    ln = a["model.encoder.layer_norm.weight"]
    assert abs(float(ln.mean()) - 1.0) < 0.05
    u = wb.weights.hash_u(0, 3, 100000)
    assert u.min() >= -131070 and u.max() <= 131070


def test_sinusoid_table_matches_hf(wb):
    from transformers.models.whisper.modeling_whisper import sinusoids
    ours = wb.weights.sinusoid_table(1500, 512)
    assert np.abs(ours - sinusoids(1500, 512).numpy()).max() < 2e-4   # HF evaluates in f32


def test_blob_roundtrip(wb, tmp_path):
    import json
    import struct
    cfg = wb.weights.WHISPER_TOY
    w = wb.weights.generate(cfg, 5)
    p = tmp_path / "w.wb200"
    wb.weights.save_blob(str(p), cfg, w)
    raw = p.read_bytes()
    assert raw[:8] == b"WB200W01"
    (jl,) = struct.unpack("<Q", raw[8:16])
    meta = json.loads(raw[16:16 + jl])
    base = (16 + jl + 63) // 64 * 64
    for ent in meta["tensors"][:5] + meta["tensors"][-3:]:
        a = np.frombuffer(raw, "<f4", ent["nbytes"] // 4, base + ent["offset"]).reshape(ent["shape"])
        assert np.array_equal(a, w[ent["name"]])
