"""Generates tests/golden/*.npz from the INSTALLED Hugging Face transformers Whisper (CPU, fp32).

Run here (container with transformers 5.5.0):  python tests/golden/make_golden.py
The GPU box never runs this; it only reads the committed .npz files.

Why HF: the reference's encoder/decoder arithmetic is ONNX Runtime executing graphs exported
from this very module; neither ORT nor the .onnx files exist offline, so the HF
module is the closest runnable definition of what `Session::run` computes
(/root/reference/src/main.rs:703, 773-776, 814).  The log-mel golden comes from
transformers.WhisperFeatureExtractor, which main.rs:318-322/418/449 says it mirrors (identical
semantics on exact-30 s clips).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "whisper-rust-ort_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import synth      # noqa: E402
import weights    # noqa: E402

PROMPT = [50258, 50259, 50359, 50363]
EOT = 50257


def hf_model(cfg: weights.ModelCfg, tensors):
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    hc = WhisperConfig(vocab_size=cfg.vocab, num_mel_bins=cfg.n_mels, d_model=cfg.d_model,
                       encoder_layers=cfg.enc_layers, decoder_layers=cfg.dec_layers,
                       encoder_attention_heads=cfg.n_heads, decoder_attention_heads=cfg.n_heads,
                       encoder_ffn_dim=cfg.ffn_dim, decoder_ffn_dim=cfg.ffn_dim,
                       max_source_positions=cfg.n_audio_ctx, max_target_positions=cfg.n_text_ctx,
                       pad_token_id=min(50256, cfg.vocab - 1), bos_token_id=min(50256, cfg.vocab - 1),
                       eos_token_id=min(50256, cfg.vocab - 1), decoder_start_token_id=min(50257, cfg.vocab - 1),
                       attn_implementation="eager")
    m = WhisperForConditionalGeneration(hc).eval()
    sd = {k: torch.from_numpy(np.array(v)) for k, v in tensors.items()}
    sd["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("proj_out" in k for k in missing), missing
    return m


@torch.no_grad()
def hf_greedy(m, enc, prompt, max_new, suppress, begin_suppress):
    """Same control flow as main.rs:753-829, logits from HF with its own KV cache."""
    B = enc.shape[0]
    ids = torch.tensor([prompt] * B)
    out = m.model.decoder(input_ids=ids, encoder_hidden_states=enc, use_cache=True)
    past = out.past_key_values
    logits = m.proj_out(out.last_hidden_state[:, -1])
    toks = [list(prompt) for _ in range(B)]
    all_logits = []
    for step in range(max_new):
        all_logits.append(logits.numpy().copy())
        lg = logits.clone()
        sup = list(suppress) + (list(begin_suppress) if step == 0 else [])
        if sup:
            lg[:, sup] = -float("inf")
        nxt = lg.argmax(-1)
        for b in range(B):
            toks[b].append(int(nxt[b]))
        if step == max_new - 1:
            break
        out = m.model.decoder(input_ids=nxt[:, None], encoder_hidden_states=enc,
                              past_key_values=past, use_cache=True)
        past = out.past_key_values
        logits = m.proj_out(out.last_hidden_state[:, -1])
    return toks, all_logits


def main():
    torch.set_num_threads(8)
    from transformers import WhisperFeatureExtractor
    from transformers.models.whisper.configuration_whisper import NON_SPEECH_TOKENS_MULTI
    fe = WhisperFeatureExtractor()

    # ---- log-mel goldens (3 clips, one per synthetic family), subsampled frames ----
    x = synth.batch(3, seed=0)
    hf_mel = np.stack([fe(x[i], sampling_rate=16000, return_tensors="np").input_features[0] for i in range(3)])
    np.savez_compressed(os.path.join(HERE, "mel_hf_seed0.npz"),
                        frames=np.arange(0, 3000, 7), mel=hf_mel[:, :, ::7].astype(np.float32),
                        edge=hf_mel[:, :, -4:].astype(np.float32))

    # the 128-bin frontend of large-v3 (BASELINE.json configs[4]) on the same clips
    fe128 = WhisperFeatureExtractor(feature_size=128)
    hf128 = np.stack([fe128(x[i], sampling_rate=16000, return_tensors="np").input_features[0] for i in range(3)])
    np.savez_compressed(os.path.join(HERE, "mel_hf128_seed0.npz"),
                        frames=np.arange(0, 3000, 7), mel=hf128[:, :, ::7].astype(np.float32),
                        edge=hf128[:, :, -4:].astype(np.float32))
    if "--mel-only" in sys.argv:
        return

    # ---- model goldens: whisper-base dims, seeded weights ----
    for tag, cfg, nclip, steps in (("base", weights.WHISPER_BASE, 2, 24), ("toy", weights.WHISPER_TOY, 2, 12)):
        W = weights.generate(cfg, seed=0)
        m = hf_model(cfg, W)
        mel = torch.from_numpy(hf_mel[:nclip])
        with torch.no_grad():
            eo = m.model.encoder(mel, output_hidden_states=True)
        enc = eo.last_hidden_state
        sup = [t for t in NON_SPEECH_TOKENS_MULTI if t < cfg.vocab]
        bsup = [t for t in (220, EOT) if t < cfg.vocab]
        prompt = PROMPT if cfg.vocab > 50363 else [1, 2, 3, 4]
        toks, logits = hf_greedy(m, enc, prompt, steps, sup, bsup)
        lg = np.stack(logits, 1)                              # [B, steps, V]
        top2 = np.sort(lg, -1)[..., -2:]
        np.savez_compressed(
            os.path.join(HERE, f"hf_whisper_{tag}_seed0.npz"),
            rows=np.arange(0, cfg.n_audio_ctx, 25),
            stem=eo.hidden_states[0][:, ::25].numpy().astype(np.float32),      # after conv stem + pos
            layer0=eo.hidden_states[1][:, ::25].numpy().astype(np.float32),
            enc=enc[:, ::25].numpy().astype(np.float32),
            tokens=np.asarray(toks, dtype=np.int64), prompt=np.asarray(prompt, dtype=np.int64),
            suppress=np.asarray(sup, dtype=np.int64), begin_suppress=np.asarray(bsup, dtype=np.int64),
            logit_cols=np.arange(0, cfg.vocab, 61), logits=lg[:, :, ::61].astype(np.float32),
            margin=(top2[..., 1] - top2[..., 0]).astype(np.float32))
        print(tag, "tokens", toks[0][:12], "min margin", float((top2[..., 1] - top2[..., 0]).min()))

    # ---- whisper-large-v3 WIDTHS (128 mels, d=1280, 20 heads, ffn 5120, vocab 51866) on 2+2 layers: BASELINE.json
    # configs[4] shapes at a depth the CPU finishes in seconds.  No 128-bin frontend exists in the reference, so the
    # input is a seeded random log-mel (the same one tests/test_gpu_wide.py feeds the GPU). ----
    import dataclasses
    cfg = dataclasses.replace(weights.WHISPER_LARGE_V3, enc_layers=2, dec_layers=2)
    W = weights.generate(cfg, seed=0)
    m = hf_model(cfg, W)
    mel = torch.from_numpy(np.random.default_rng(3).normal(0.0, 0.5, (2, 128, 3000)).astype(np.float32)[:1])
    with torch.no_grad():
        eo = m.model.encoder(mel, output_hidden_states=True)
    enc = eo.last_hidden_state
    prompt = [50258, 50259, 50360, 50364]
    toks, logits = hf_greedy(m, enc, prompt, 6, [], [])
    lg = np.stack(logits, 1)
    top2 = np.sort(lg, -1)[..., -2:]
    np.savez_compressed(
        os.path.join(HERE, "hf_whisper_wide_seed0.npz"),
        rows=np.arange(0, cfg.n_audio_ctx, 50),
        stem=eo.hidden_states[0][:, ::50].numpy().astype(np.float32),
        enc=enc[:, ::50].numpy().astype(np.float32),
        tokens=np.asarray(toks, dtype=np.int64), prompt=np.asarray(prompt, dtype=np.int64),
        logit_cols=np.arange(0, cfg.vocab, 97), logits=lg[:, :, ::97].astype(np.float32),
        margin=(top2[..., 1] - top2[..., 0]).astype(np.float32))
    print("wide tokens", toks[0], "min margin", float((top2[..., 1] - top2[..., 0]).min()))


if __name__ == "__main__":
    main()
