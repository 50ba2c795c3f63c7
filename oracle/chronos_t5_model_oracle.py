"""CPU ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/timesfm_oracle.py header for the import rules).

Chronos-T5 forecast path (BASELINE.json configs[2]; NOT part of the reference, which only wraps Chronos-2): the
tokeniser oracle (oracle/chronos_t5_oracle.py) around the importable ``transformers.T5ForConditionalGeneration``, which
is the very class upstream ``chronos.ChronosModel`` drives (T5Config(vocab_size=4096,
d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_heads=12, feed_forward_proj="relu")).  The three stages mirror
``tsfmx_b200.tsfm.chronos_t5.ChronosT5Adapter``:

  preprocess  : padded steps -> NaN, mean-scale + uniform-bin ids with EOS, ``shared`` embedding lookup
  forward     : ``model.encoder(inputs_embeds=..., attention_mask=...)``
  postprocess : ``model.generate(encoder_outputs=..., do_sample=False, min_new_tokens = max_new_tokens = horizon)``
                (upstream ChronosModel.forward passes min_new_tokens = prediction_length, so EOS cannot end a path
                early), ids -> ``centers[id - 3] * scale``

PARITY PINNING: pinned to the HF T5 implementation installed here (transformers 5.5); "parity unpinned" against the
upstream ``chronos`` package itself (not available offline) — its sampling settings (num_samples 20, top-k 50) are
replaced by greedy decoding so that token ids are comparable.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from . import chronos_t5_oracle as TK


@dataclass
class OraclePreprocess:
    input_embeddings: torch.Tensor
    masks: torch.Tensor
    normalization_stats: dict


class OracleChronosT5Adapter(nn.Module):
    def __init__(self, hf_model) -> None:
        super().__init__()
        self.model = hf_model.eval()
        self.centers, self.boundaries = TK.tables()

    @property
    def model_dims(self) -> int:
        return int(self.model.config.d_model)

    @property
    def patch_len(self) -> int:
        return 1

    @property
    def point_forecast_index(self) -> int:
        return 0

    def preprocess(self, inputs: torch.Tensor, masks: torch.Tensor) -> OraclePreprocess:
        if masks.shape != inputs.shape:
            raise ValueError(f"masks shape {masks.shape} must match inputs shape {inputs.shape}")
        x = torch.where(masks.bool(), torch.full_like(inputs, float("nan")), inputs.float())
        ids, attention_mask, scale = TK.tokenize(x, self.boundaries)
        emb = self.model.shared(ids)
        return OraclePreprocess(emb, ~attention_mask, {"scale": scale, "token_ids": ids})

    def forward(self, input_embeddings: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        out = self.model.encoder(inputs_embeds=input_embeddings, attention_mask=(~masks.bool()).long())
        return out.last_hidden_state

    def decode(self, encoder_states, attention_mask, horizon):
        from transformers.modeling_outputs import BaseModelOutput

        out = self.model.generate(
            encoder_outputs=BaseModelOutput(last_hidden_state=encoder_states), attention_mask=attention_mask.long(),
            do_sample=False, num_beams=1, max_new_tokens=horizon, min_new_tokens=horizon,
        )
        return out[:, 1:]  # drop the decoder start token

    def teacher_forced_logits(self, encoder_states, attention_mask, tokens):
        """Logits of every step when the decoder is fed [start, tokens[:-1]]."""
        from transformers.modeling_outputs import BaseModelOutput

        start = torch.full((tokens.shape[0], 1), self.model.config.decoder_start_token_id, dtype=tokens.dtype)
        dec_in = torch.cat([start, tokens[:, :-1]], 1)
        return self.model(encoder_outputs=BaseModelOutput(last_hidden_state=encoder_states),
                          attention_mask=attention_mask.long(), decoder_input_ids=dec_in).logits

    def postprocess(self, horizon: int, output_embeddings: torch.Tensor, normalization_stats: dict) -> torch.Tensor:
        ids = normalization_stats["token_ids"]
        tokens = self.decode(output_embeddings, ids != TK.PAD_ID, horizon)
        return TK.dequantize(tokens, self.centers, normalization_stats["scale"]).unsqueeze(-1)


def hf_model_from_product(adapter):
    """A transformers T5ForConditionalGeneration carrying the product adapter's weights (same key names)."""
    from transformers import T5Config, T5ForConditionalGeneration

    m = adapter._model
    cfg = T5Config(vocab_size=m.vocab_size, d_model=m.d_model, d_kv=m.d_kv, d_ff=m.d_ff,
                   num_layers=len(m.encoder.block), num_decoder_layers=len(m.decoder.block), num_heads=m.num_heads,
                   feed_forward_proj="relu", tie_word_embeddings=m.tie_word_embeddings, pad_token_id=m.pad_token_id,
                   eos_token_id=m.eos_token_id, decoder_start_token_id=m.decoder_start_token_id, dropout_rate=0.0)
    hf = T5ForConditionalGeneration(cfg).eval()
    state = {k: v.detach().cpu().float() for k, v in m.state_dict().items()}
    missing, unexpected = hf.load_state_dict(state, strict=False)
    allowed = {"encoder.embed_tokens.weight", "decoder.embed_tokens.weight"}
    if m.tie_word_embeddings:
        allowed.add("lm_head.weight")
    assert not unexpected and set(missing) <= allowed, (missing, unexpected)
    if m.tie_word_embeddings:
        hf.tie_weights()
    else:
        # transformers 5.x builds T5ForConditionalGeneration with lm_head.weight aliasing shared.weight even when
        # config.tie_word_embeddings is False (its forward then skips the d_model ** -0.5 rescale, as it should):
        # give the head its own parameter and restore the embedding the load just overwrote.
        hf.lm_head.weight = nn.Parameter(state["lm_head.weight"].clone())
        with torch.no_grad():
            hf.shared.weight.copy_(state["shared.weight"])
            if hf.encoder.embed_tokens.weight.data_ptr() != hf.shared.weight.data_ptr():
                hf.encoder.embed_tokens.weight.copy_(state["shared.weight"])
            if hf.decoder.embed_tokens.weight.data_ptr() != hf.shared.weight.data_ptr():
                hf.decoder.embed_tokens.weight.copy_(state["shared.weight"])
        assert torch.equal(hf.lm_head.weight, state["lm_head.weight"])
    assert torch.equal(hf.shared.weight, state["shared.weight"])
    assert torch.equal(hf.decoder.embed_tokens.weight, state["shared.weight"])
    return hf


def oracle_from_product(decoder):
    """Oracle twin of a product MultimodalDecoder built around a ChronosT5Adapter: the reference's own decoder and
    fusion classes when /root/reference is importable, else the restated ones from timesfm_oracle."""
    from . import timesfm_oracle as O

    adapter = OracleChronosT5Adapter(hf_model_from_product(decoder.adapter))
    fus = decoder.fusion
    dims = [l.weight.shape[1] for l in fus.linears()] + [fus.linears()[-1].weight.shape[0]]
    o = O.OracleDecoder(adapter, dims[0], len(dims) - 1, dims[1:-1])
    with torch.no_grad():
        for src, dst in zip(fus.linears(), [m for m in o.fusion.projection if isinstance(m, nn.Linear)]):
            dst.weight.copy_(src.weight.detach().float().cpu())
    return o.eval()
