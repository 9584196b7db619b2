"""GPU parity of kernel groups 2+3 (encoder, KV-cache greedy decode) on the fp32 validation build
against the numpy oracle (oracle/whisper_ref.py) and the committed HF golden vectors.
Tolerances (BASELINE.json north_star): encoder hidden states 1e-4 (fp32 build), identical greedy
token ids."""
import numpy as np
import pytest

import mel_oracle as mo
import whisper_ref as wr

pytestmark = pytest.mark.gpu
EOT = 50257


@pytest.fixture(scope="module")
def base(wb):
    m = wb.Whisper(wb.default_cfg("base", max_batch=4, max_chunks=8))
    m.set_debug(True)
    yield m
    m.close()


@pytest.fixture(scope="module")
def oracle(wb):
    cfg = wb.weights.WHISPER_BASE
    return wr.WhisperRef(cfg, wb.weights.generate(cfg, 0))


@pytest.fixture(scope="module")
def clips(wb):
    x = wb.synth.batch(2, seed=0)
    return x, np.stack([mo.log_mel(c) for c in x])


def test_device_weights_are_bit_identical_to_python_generator(wb, base, oracle):
    for name in ("model.encoder.conv1.weight", "model.encoder.layers.3.fc1.bias",
                 "model.decoder.layers.5.encoder_attn_layer_norm.weight", "model.decoder.embed_positions.weight"):
        w = oracle.w[name]
        assert np.array_equal(base.tensor(name, w.shape), w), name
    pos = oracle.w["model.encoder.embed_positions.weight"]
    assert np.abs(base.tensor("model.encoder.embed_positions.weight", pos.shape) - pos).max() <= 1.2e-7


def test_encoder_vs_oracle_and_hf_golden(base, oracle, clips, golden_dir):
    _, mel = clips
    enc = base.encode(mel)
    ref, layers = oracle.encode(mel, return_layers=True)
    assert np.abs(base.encoder_debug("stem", 2) - layers[0]).max() <= 1e-4
    assert np.abs(base.encoder_debug("layer0", 2) - layers[1]).max() <= 1e-4
    assert np.abs(enc - ref).max() <= 1e-4
    g = np.load(f"{golden_dir}/hf_whisper_base_seed0.npz")
    assert np.abs(enc[:, g["rows"]] - g["enc"]).max() <= 1e-4


def test_encoder_from_resident_chunks_matches_host_mel_path(base, clips):
    x, mel = clips
    a = base.encode(mel)
    _, n = base.log_mel(x, want_mel=False)
    assert n == 2
    b = base.encode(None, 0, 2)
    assert np.abs(a - b).max() <= 1e-4


def test_greedy_tokens_identical_and_logits_close(base, oracle, clips, golden_dir):
    _, mel = clips
    g = np.load(f"{golden_dir}/hf_whisper_base_seed0.npz")
    steps = g["tokens"].shape[1] - len(g["prompt"])
    enc = base.encode(mel)
    toks, logits = base.greedy_decode(2, g["prompt"], steps, EOT, g["suppress"], g["begin_suppress"], want_logits=True)
    ref_toks, ref_logits = oracle.greedy(enc, g["prompt"], steps, EOT, g["suppress"], g["begin_suppress"], return_logits=True)
    assert toks == ref_toks
    assert np.array_equal(np.array(toks), g["tokens"])
    assert np.abs(logits - np.stack(ref_logits, 1)).max() <= 1e-4
    assert np.abs(logits[:, :, g["logit_cols"]] - g["logits"]).max() <= 1e-4


def test_eot_stops_a_sequence_and_max_new_quirks(base, oracle, clips):
    _, mel = clips
    enc = base.encode(mel)
    prompt = [50258, 50259, 50359, 50363]
    free = base.greedy_decode(2, prompt, 6, EOT)
    # declare the 3rd generated token of sequence 0 to be "eot": it must stop there, sequence 1 not
    fake_eot = free[0][len(prompt) + 2]
    got = base.greedy_decode(2, prompt, 6, fake_eot)
    ref = oracle.greedy(enc, prompt, 6, fake_eot)
    assert got == ref
    assert got[0][-1] == fake_eot and len(got[0]) <= len(prompt) + 3
    # `for _ in 1..max_new_tokens` (main.rs:793): max_new 0 and 1 both yield exactly one token
    for mn in (0, 1):
        out = base.greedy_decode(2, prompt, mn, EOT)
        assert [len(s) for s in out] == [5, 5] and out == oracle.greedy(enc, prompt, mn, EOT)


def test_early_exit_when_every_sequence_hit_eot(base, oracle, clips):
    """The device loop stops at a segment boundary once all sequences emitted EOT (the host looks at one
    mapped int per 16 tokens, never per token, and with one segment of look-ahead so the GPU never idles on the
    poll: one surplus segment may run); outputs stay those of main.rs:753-829."""
    _, mel = clips
    enc = base.encode(mel[:1])
    prompt = [50258, 50259, 50359, 50363]
    free = base.greedy_decode(1, prompt, 40, EOT)
    assert base.timing()["decode_steps"] == len(prompt) + 40 - 1
    first = free[0][len(prompt)]
    got = base.greedy_decode(1, prompt, 40, first)                  # "eot" = the first generated token
    assert got == oracle.greedy(enc, prompt, 40, first) == [prompt + [first]]
    assert base.timing()["decode_steps"] == len(prompt) - 1 + 32     # prefix + the segment with the EOT + one of look-ahead
    # a batch only stops when ALL of its sequences are done
    base.encode(mel)
    both = base.greedy_decode(2, prompt, 40, first)
    ref = oracle.greedy(oracle.encode(mel), prompt, 40, first)
    assert both == ref
    if any(len(s) == len(prompt) + 40 for s in ref):
        assert base.timing()["decode_steps"] == len(prompt) + 40 - 1


def test_transcribe_batch_end_to_end(wb, base, oracle):
    x = wb.synth.batch(3, seed=5)
    prompt = [50258, 50259, 50359, 50363]
    toks, fidx = base.transcribe_batch(x, prompt, 8, EOT)
    assert fidx.tolist() == [0, 1, 2]
    mel = np.stack([mo.log_mel(c) for c in x])
    ref = wr.transcribe_tokens(oracle, mel, prompt, 8, EOT)
    assert toks == ref
