"""BLIP-2 (ViT-g/14 + Q-Former, ITC head) dimensions and seeded synthetic weights.

Names follow the HF ``Blip2ForImageTextRetrieval`` state dict (which is also how the
public ``Salesforce/blip2-itm-vit-g`` checkpoint is laid out), so a converted real
checkpoint loads into the engine unchanged.  No checkpoint is available offline: tests
and benchmarks use ``random_state_dict`` (seeded, non-trivial biases and LayerNorm
parameters so that every term of the forward is exercised).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import torch


@dataclass(frozen=True)
class Blip2Dims:
    v_hidden: int = 1408
    v_inter: int = 6144
    v_layers: int = 39
    v_heads: int = 16
    image: int = 224
    patch: int = 14
    v_eps: float = 1e-6
    q_hidden: int = 768
    q_inter: int = 3072
    q_layers: int = 12
    q_heads: int = 12
    q_eps: float = 1e-12
    cross_freq: int = 2
    queries: int = 32
    proj: int = 256
    vocab: int = 30522
    max_pos: int = 512

    @property
    def tokens(self) -> int:
        return (self.image // self.patch) ** 2 + 1

    @property
    def patch_k(self) -> int:
        return 3 * self.patch * self.patch

    @property
    def patch_k_padded(self) -> int:
        return (self.patch_k + 7) // 8 * 8

    def flops_per_image(self) -> float:
        """Algorithmic FLOPs of one image through ViT + Q-Former query path."""
        n, d, f = self.tokens, self.v_hidden, self.v_inter
        vit = self.v_layers * (8 * n * d * d + 4 * n * d * f + 4 * n * n * d) + 2 * (n - 1) * self.patch_k * d
        q, h, fi = self.queries, self.q_hidden, self.q_inter
        ncross = (self.q_layers + self.cross_freq - 1) // self.cross_freq
        qf = self.q_layers * (8 * q * h * h + 4 * q * q * h + 4 * q * h * fi)
        qf += ncross * (4 * n * d * h + 4 * q * h * h + 4 * q * n * h)
        return float(vit + qf + 2 * q * h * self.proj)


TINY = Blip2Dims(v_hidden=64, v_inter=128, v_layers=2, v_heads=4, image=56, patch=14, q_hidden=64, q_inter=128,
                 q_layers=2, q_heads=2, queries=8, proj=16, vocab=100, max_pos=40)
SMALL = Blip2Dims(v_hidden=176, v_inter=384, v_layers=3, v_heads=2, image=224, patch=14, q_hidden=128, q_inter=256,
                  q_layers=4, q_heads=2, queries=32, proj=64, vocab=200, max_pos=64)


def hf_config(d: Blip2Dims):
    """Equivalent HF config (used by the oracle only)."""
    from transformers import Blip2Config
    from transformers.models.blip_2.configuration_blip_2 import Blip2QFormerConfig, Blip2VisionConfig

    vc = Blip2VisionConfig(hidden_size=d.v_hidden, intermediate_size=d.v_inter, num_hidden_layers=d.v_layers,
                           num_attention_heads=d.v_heads, image_size=d.image, patch_size=d.patch,
                           layer_norm_eps=d.v_eps, hidden_act="gelu", qkv_bias=True)
    qc = Blip2QFormerConfig(hidden_size=d.q_hidden, num_hidden_layers=d.q_layers, num_attention_heads=d.q_heads,
                            intermediate_size=d.q_inter, encoder_hidden_size=d.v_hidden, use_qformer_text_input=True,
                            vocab_size=d.vocab, max_position_embeddings=d.max_pos, layer_norm_eps=d.q_eps,
                            cross_attention_frequency=d.cross_freq, hidden_act="gelu")
    return Blip2Config(vision_config=vc.to_dict(), qformer_config=qc.to_dict(), num_query_tokens=d.queries,
                       image_text_hidden_size=d.proj)


def random_state_dict(d: Blip2Dims, seed: int = 0, outliers: bool = False) -> Dict[str, torch.Tensor]:
    """Seeded fp32 CPU weights under HF names.  Linear weights ~ N(0, 1/sqrt(fan_in)) * 0.8
    (keeps activations O(1) through 39 pre-LN blocks), biases ~ N(0, 0.05),
    LayerNorm gamma ~ 1 + N(0, 0.05), beta ~ N(0, 0.05).

    ``outliers=True`` adds what trained ViT / BERT checkpoints show and plain Gaussians do not: a few LayerNorm channels with
    gains several times the rest and offsets of order one, and a handful of residual-stream channels carrying "massive
    activations" (large biases on the block outputs) -- the cases where half-precision operands lose the most.

    The ViT's Linear / Conv weights and biases are fp16-representable values (stored as fp32): the reference keeps them in half
    precision (lavis `create_eva_vit_g(..., precision="fp16")` -> `convert_weights_to_fp16`, which converts exactly the
    Conv / Linear parameters; LayerNorm and the Q-Former stay fp32), so a real checkpoint never carries more mantissa there."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    hot = {}

    def lin(name: str, out_f: int, in_f: int, bias: bool = True) -> None:
        sd[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * (0.8 / in_f**0.5)
        if bias:
            sd[name + ".bias"] = torch.randn(out_f, generator=g) * 0.05

    def ln(name: str, n: int) -> None:
        sd[name + ".weight"] = 1.0 + torch.randn(n, generator=g) * 0.05
        sd[name + ".bias"] = torch.randn(n, generator=g) * 0.05
        if outliers:
            idx = hot.setdefault(n, torch.randperm(n, generator=g)[:6])
            sd[name + ".weight"][idx] *= torch.tensor([6.0, 5.0, 4.0, 3.0, 0.2, 0.1])
            sd[name + ".bias"][idx[:3]] += torch.tensor([1.0, -1.5, 0.7])

    D, F = d.v_hidden, d.v_inter
    sd["query_tokens"] = torch.randn(1, d.queries, d.q_hidden, generator=g) * 0.5
    sd["vision_model.embeddings.class_embedding"] = torch.randn(1, 1, D, generator=g) * 0.5
    sd["vision_model.embeddings.position_embedding"] = torch.randn(1, d.tokens, D, generator=g) * 0.3
    sd["vision_model.embeddings.patch_embedding.weight"] = torch.randn(D, 3, d.patch, d.patch, generator=g) * (0.8 / d.patch_k**0.5)
    sd["vision_model.embeddings.patch_embedding.bias"] = torch.randn(D, generator=g) * 0.05
    for i in range(d.v_layers):
        p = f"vision_model.encoder.layers.{i}."
        lin(p + "self_attn.qkv", 3 * D, D)
        sd[p + "self_attn.qkv.bias"][D : 2 * D] = 0.0  # EVA: no key bias (modeling_blip_2.py Blip2Attention)
        lin(p + "self_attn.projection", D, D)
        ln(p + "layer_norm1", D)
        lin(p + "mlp.fc1", F, D)
        lin(p + "mlp.fc2", D, F)
        ln(p + "layer_norm2", D)
    ln("vision_model.post_layernorm", D)
    H, I = d.q_hidden, d.q_inter
    sd["embeddings.word_embeddings.weight"] = torch.randn(d.vocab, H, generator=g) * 0.5
    sd["embeddings.position_embeddings.weight"] = torch.randn(d.max_pos, H, generator=g) * 0.3
    ln("qformer.layernorm", H)
    for i in range(d.q_layers):
        p = f"qformer.encoder.layer.{i}."
        for blk, kin in (("attention", H),) + ((("crossattention", D),) if i % d.cross_freq == 0 else ()):
            lin(p + blk + ".attention.query", H, H)
            lin(p + blk + ".attention.key", H, kin)
            lin(p + blk + ".attention.value", H, kin)
            lin(p + blk + ".output.dense", H, H)
            ln(p + blk + ".output.LayerNorm", H)
        lin(p + "intermediate.dense", I, H)
        lin(p + "output.dense", H, I)
        ln(p + "output.LayerNorm", H)
        lin(p + "intermediate_query.dense", I, H)
        lin(p + "output_query.dense", H, I)
        ln(p + "output_query.LayerNorm", H)
    if outliers:     # massive activations: two residual channels of the ViT and of the Q-Former get a large constant push mid-network
        for i in (d.v_layers // 3, 2 * d.v_layers // 3):
            b = sd[f"vision_model.encoder.layers.{i}.mlp.fc2.bias"]
            b[hot[D][:2]] += torch.tensor([25.0, -18.0])
        b = sd[f"qformer.encoder.layer.{d.q_layers // 2}.output_query.dense.bias"]
        b[hot[H][:2]] += torch.tensor([6.0, -4.0])
    lin("vision_projection", d.proj, H)
    lin("text_projection", d.proj, H)
    lin("itm_head", 2, H)
    for k in list(sd):  # reference storage precision of the ViT's Conv / Linear parameters (see the docstring)
        if k.startswith("vision_model.") and ("layer_norm" not in k and "layernorm" not in k and "embedding" not in k.split(".")[-1]):
            sd[k] = sd[k].half().float()
    return sd
