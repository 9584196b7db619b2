"""Test helper: a REAL torch.onnx export of HF's WhisperForConditionalGeneration (toy shape, random weights), laid out like
the reference's optimum export (scripts/export_onnx_whisper.py:20-28): `encoder_model.onnx` from `model.get_encoder()`,
`decoder_model.onnx` from the decoder + tied `proj_out` (names `model.decoder.*`).  This is the exporter optimum itself
calls (TorchScript-based `torch.onnx.export`), so the files carry its real behaviour: anonymous transposed
`onnx::MatMul_<n>` Linear weights, the encoder position table folded into `onnx::Add_<n>`, identical initializers
de-duplicated into Identity aliases.  The `onnx` Python package is absent offline; the exporter only needs it for a
post-pass that adds onnxscript functions (none here), which is stubbed out."""
import os
import warnings


SHAPES = {   # wb200.weights.WHISPER_TOY / WHISPER_BASE as HF config fields
    "toy": dict(vocab_size=1031, d_model=128, layers=2, heads=2, ffn=256),
    "base": dict(vocab_size=51865, d_model=512, layers=6, heads=8, ffn=2048),
}


def export(out_dir: str, randomize: bool, seed: int = 0, shape: str = "toy"):
    """-> HF state_dict as {name: float32 ndarray} (the names wb200.weights.tensor_specs uses)."""
    import torch
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
    from transformers import WhisperConfig, WhisperForConditionalGeneration

    onnx_proto_utils._add_onnxscript_fn = lambda proto, custom_opsets: proto
    sh = SHAPES[shape]
    cfg = WhisperConfig(vocab_size=sh["vocab_size"], num_mel_bins=80, d_model=sh["d_model"], encoder_layers=sh["layers"],
                        decoder_layers=sh["layers"], encoder_attention_heads=sh["heads"], decoder_attention_heads=sh["heads"],
                        encoder_ffn_dim=sh["ffn"], decoder_ffn_dim=sh["ffn"], max_source_positions=1500, max_target_positions=448,
                        pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=1, attn_implementation="eager")
    torch.manual_seed(seed)
    m = WhisperForConditionalGeneration(cfg).eval()
    if randomize:                       # HF initialises biases to 0 and LayerNorm weights to 1: make every tensor distinct
        with torch.no_grad():
            for n, p in m.named_parameters():
                if "encoder.embed_positions" in n:
                    continue
                p.copy_(torch.randn_like(p) * 0.05 + (1.0 if "layer_norm.weight" in n else 0.0))

    class Decoder(torch.nn.Module):     # the decoder stack + tied projection, traced layer by layer (HF's mask helper does not trace)
        def __init__(self, model):
            super().__init__()
            self.model, self.proj_out = model.model, model.proj_out

        def forward(self, input_ids, encoder_hidden_states, position_ids):
            d = self.model.decoder
            t = input_ids.shape[1]
            x = d.embed_tokens(input_ids) + torch.nn.functional.embedding(position_ids, d.embed_positions.weight)
            mask = torch.full((1, 1, t, t), float("-inf")).triu(1)
            for layer in d.layers:
                o = layer(x, attention_mask=mask, encoder_hidden_states=encoder_hidden_states)
                x = o[0] if isinstance(o, (tuple, list)) else o
            return self.proj_out(d.layer_norm(x))

    os.makedirs(out_dir, exist_ok=True)
    with warnings.catch_warnings(), torch.no_grad():
        warnings.simplefilter("ignore")
        torch.onnx.export(m.get_encoder(), (torch.randn(1, 80, 3000),), os.path.join(out_dir, "encoder_model.onnx"), dynamo=False,
                          opset_version=14, input_names=["input_features"], output_names=["last_hidden_state"])
        torch.onnx.export(Decoder(m), (torch.tensor([[1, 5, 7, 9]]), torch.randn(1, 1500, sh["d_model"]), torch.tensor([[0, 1, 2, 3]])),
                          os.path.join(out_dir, "decoder_model.onnx"), dynamo=False, opset_version=14,
                          input_names=["input_ids", "encoder_hidden_states", "position_ids"], output_names=["logits"])
    return {k: v.detach().numpy().astype("float32") for k, v in m.state_dict().items()}
