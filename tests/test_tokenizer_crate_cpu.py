"""CPU: the C++ tokenizer.json reader / detokeniser against the Hugging Face `tokenizers` package — Python bindings to
the very Rust crate the reference links (tokenizers 0.15.2, /root/reference/Cargo.lock:1410; the image has 0.22.2 of
the same crate).  The three uses the reference makes of it:
  Tokenizer::from_file  /root/reference/src/main.rs:580
  token_to_id           /root/reference/src/main.rs:531
  decode(ids, true)     /root/reference/src/main.rs:640
No tokenizer.json ships offline, so one is TRAINED here with the crate itself (byte-level BPE, the Whisper/GPT-2 layout:
ByteLevel pre-tokenizer + ByteLevel decoder, Whisper's special tokens, non-special added tokens) and saved with the
crate's own serializer; our reader then has to agree with the crate on that file.
"""
import ctypes as C
import json

import numpy as np
import pytest

from test_host_cpu import L, decode            # noqa: F401  (fixture + helper)

tokenizers = pytest.importorskip("tokenizers")

CORPUS = [
    "Hello world! This is a test of the emergency broadcast system.",
    "The quick brown fox jumps over the lazy dog, again and again and again.",
    "Ça va? Très bien, merci. Le café est prêt: naïve façade, œuvre, Ångström.",
    "日本語のテキストも少しだけ入れておきます。漢字とかなとカナ。",
    "Emoji \U0001F600\U0001F680 and astral \U00010400 characters; tabs\tand\nnewlines\r\n too.",
    "numbers 0123456789 3.14159 1,000,000 and symbols #@$%^&*()[]{}<>|\\/~`",
    "हिन्दी में भी कुछ शब्द होने चाहिए।",
] * 4

SPECIALS = ["<|endoftext|>", "<|startoftranscript|>", "<|en|>", "<|hi|>", "<|translate|>", "<|transcribe|>", "<|notimestamps|>"]
PLAIN_ADDED = ["<|0.00|>", "<|0.02|>", " custom phrase", "日本"]


@pytest.fixture(scope="module")
def trained(tmp_path_factory):
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, trainers, AddedToken
    tok = Tokenizer(models.BPE())
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tok.decoder = decoders.ByteLevel()
    trainer = trainers.BpeTrainer(vocab_size=600, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False)
    tok.train_from_iterator(CORPUS, trainer)
    tok.add_special_tokens([AddedToken(s, special=True) for s in SPECIALS])
    tok.add_tokens([AddedToken(s, special=False) for s in PLAIN_ADDED])
    p = tmp_path_factory.mktemp("crate") / "tokenizer.json"
    tok.save(str(p))
    return tok, p


def _load(L, path):
    t = C.c_void_p()
    assert L.wb_tokenizer_load(C.byref(t), str(path).encode()) == 0, L.wb_last_error()
    return t


def test_token_to_id_matches_crate(L, trained):
    tok, p = trained
    t = _load(L, p)
    vocab = tok.get_vocab(with_added_tokens=True)
    assert len(vocab) > 300
    for s, i in vocab.items():
        assert L.wb_tokenizer_token_to_id(t, s.encode("utf-8")) == i == tok.token_to_id(s)
    for s in ["<|xx|>", "", "definitely-not-a-token", "<|EN|>", "Ġ" * 40]:
        want = tok.token_to_id(s)
        assert L.wb_tokenizer_token_to_id(t, s.encode("utf-8")) == (-1 if want is None else want)
    L.wb_tokenizer_free(t)


def test_decode_matches_crate_on_encoded_text(L, trained):
    tok, p = trained
    t = _load(L, p)
    texts = CORPUS[:7] + ["<|startoftranscript|><|en|><|transcribe|><|notimestamps|> Hello world<|endoftext|>",
                          "<|0.00|> custom phrase 日本<|0.02|>", "", " ", "a", "\U0001F600"]
    for s in texts:
        ids = tok.encode(s).ids
        assert decode(L, t, ids) == tok.decode(ids, skip_special_tokens=True)
    L.wb_tokenizer_free(t)


def test_decode_matches_crate_on_random_ids(L, trained):
    """Random id sequences: byte-level pieces that split multi-byte characters (-> U+FFFD via from_utf8_lossy), special and
    non-special added tokens in any position, ids past the vocabulary (dropped by the crate's filter_map)."""
    tok, p = trained
    t = _load(L, p)
    n = tok.get_vocab_size(with_added_tokens=True)
    rng = np.random.default_rng(0)
    for _ in range(400):
        k = int(rng.integers(0, 40))
        ids = [int(x) for x in rng.integers(0, n + 20, k)]
        assert decode(L, t, ids) == tok.decode(ids, skip_special_tokens=True), ids
    # what main.rs:639 drops before the crate sees it: ids outside u32
    ids = [int(x) for x in rng.integers(0, n, 12)]
    assert decode(L, t, [-1] + ids[:6] + [2**40] + ids[6:]) == tok.decode(ids, skip_special_tokens=True)
    L.wb_tokenizer_free(t)


def test_special_tokens_through_crate_file(L, trained):
    tok, p = trained
    t = _load(L, p)
    out = (C.c_int64 * 5)()
    assert L.wb_host_special_tokens(t, b"hi", b"translate", out) == 0
    assert list(out) == [tok.token_to_id(s) for s in ("<|startoftranscript|>", "<|endoftext|>", "<|hi|>", "<|translate|>", "<|notimestamps|>")]
    L.wb_tokenizer_free(t)


def test_no_decoder_joins_with_spaces_like_the_crate(L, trained, tmp_path):
    """Tokenizer::decode without a decoder section joins the surviving tokens with ' ' (tokenizers/src/tokenizer/mod.rs)."""
    from tokenizers import Tokenizer
    tok, p = trained
    d = json.loads(p.read_text())
    d["decoder"] = None
    q = tmp_path / "tokenizer.json"
    q.write_text(json.dumps(d))
    ref = Tokenizer.from_file(str(q))
    t = _load(L, q)
    ids = tok.encode("<|en|> Hello world, again").ids
    assert decode(L, t, ids) == ref.decode(ids, skip_special_tokens=True)
    L.wb_tokenizer_free(t)
