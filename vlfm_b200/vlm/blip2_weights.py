"""Checkpoint ingestion for BLIP-2 ITM (vlfm/vlm/blip2itm.py:29-34: ``load_model_and_preprocess(name="blip2_image_text_matching",
model_type="pretrain")``).

lavis stores the model (``Blip2ITM`` = ``Blip2Qformer``) under ITS module names: ``visual_encoder.*`` (EVA ViT-g, fused qkv with
separate ``q_bias`` / ``v_bias`` and NO key bias), ``ln_vision``, ``Qformer.bert.*``, ``query_tokens``, ``vision_proj``,
``text_proj``, ``itm_head``, ``temp``.  The engine consumes the HF ``Blip2ForImageTextRetrieval`` layout (``blip2_config.py``).
This module holds the key-mapping table and a shape check against the engine's dimensions.

Status: no checkpoint exists offline; the table is checked structurally (tests/test_checkpoint_keymaps.py): a synthetic
lavis-named dict converts to exactly the keys / shapes ``random_state_dict`` produces, values bit-identical.  Note the real
pretrain file is split in two by lavis (``eva_vit_g.pth`` for the ViT, ``blip2_pretrained.pth`` for the rest, with the ViT's last
block dropped: 39 of 40): ``load_checkpoint`` accepts either a merged lavis ``model.state_dict()`` dump or an HF-layout dict.
"""
from __future__ import annotations

import re
from typing import Dict, List, Tuple

import torch

RENAMES: List[Tuple[str, str]] = [
    (r"^visual_encoder\.cls_token$", "vision_model.embeddings.class_embedding"),
    (r"^visual_encoder\.pos_embed$", "vision_model.embeddings.position_embedding"),
    (r"^visual_encoder\.patch_embed\.proj\.(weight|bias)$", r"vision_model.embeddings.patch_embedding.\1"),
    (r"^visual_encoder\.blocks\.(\d+)\.norm1\.(weight|bias)$", r"vision_model.encoder.layers.\1.layer_norm1.\2"),
    (r"^visual_encoder\.blocks\.(\d+)\.norm2\.(weight|bias)$", r"vision_model.encoder.layers.\1.layer_norm2.\2"),
    (r"^visual_encoder\.blocks\.(\d+)\.attn\.qkv\.weight$", r"vision_model.encoder.layers.\1.self_attn.qkv.weight"),
    (r"^visual_encoder\.blocks\.(\d+)\.attn\.proj\.(weight|bias)$", r"vision_model.encoder.layers.\1.self_attn.projection.\2"),
    (r"^visual_encoder\.blocks\.(\d+)\.mlp\.(fc1|fc2)\.(weight|bias)$", r"vision_model.encoder.layers.\1.mlp.\2.\3"),
    (r"^ln_vision\.(weight|bias)$", r"vision_model.post_layernorm.\1"),
    (r"^query_tokens$", "query_tokens"),
    (r"^Qformer\.bert\.embeddings\.LayerNorm\.(weight|bias)$", r"qformer.layernorm.\1"),
    (r"^Qformer\.bert\.embeddings\.(word|position)_embeddings\.weight$", r"embeddings.\1_embeddings.weight"),
    (r"^Qformer\.bert\.encoder\.layer\.(\d+)\.(attention|crossattention)\.self\.(query|key|value)\.(weight|bias)$",
     r"qformer.encoder.layer.\1.\2.attention.\3.\4"),
    (r"^Qformer\.bert\.encoder\.layer\.(\d+)\.(attention|crossattention)\.output\.(dense|LayerNorm)\.(weight|bias)$",
     r"qformer.encoder.layer.\1.\2.output.\3.\4"),
    (r"^Qformer\.bert\.encoder\.layer\.(\d+)\.(intermediate|intermediate_query)\.dense\.(weight|bias)$", r"qformer.encoder.layer.\1.\2.dense.\3"),
    (r"^Qformer\.bert\.encoder\.layer\.(\d+)\.(output|output_query)\.(dense|LayerNorm)\.(weight|bias)$", r"qformer.encoder.layer.\1.\2.\3.\4"),
    (r"^vision_proj\.(weight|bias)$", r"vision_projection.\1"),
    (r"^text_proj\.(weight|bias)$", r"text_projection.\1"),
    (r"^itm_head\.(weight|bias)$", r"itm_head.\1"),
]
# not used by the ITC forward: LM head of the Q-Former's BertLMHeadModel, the learned temperature, index buffers, rope/rel-pos leftovers
IGNORED = [r"^Qformer\.cls\..*$", r"^temp$", r"^Qformer\.bert\.embeddings\.position_ids$", r"^visual_encoder\.blocks\.\d+\.attn\.relative_position_index$"]


def is_lavis_layout(sd: Dict[str, torch.Tensor]) -> bool:
    return any(k.startswith(("visual_encoder.", "Qformer.", "ln_vision.")) for k in sd)


def convert_lavis_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """lavis ``Blip2ITM.state_dict()`` names -> HF ``Blip2ForImageTextRetrieval`` names.  Unknown names raise."""
    out: Dict[str, torch.Tensor] = {}
    qb: Dict[str, torch.Tensor] = {}
    vb: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        k = k[7:] if k.startswith("module.") else k
        if any(re.match(p, k) for p in IGNORED):
            continue
        m = re.match(r"^visual_encoder\.blocks\.(\d+)\.attn\.(q|v)_bias$", k)
        if m:
            (qb if m.group(2) == "q" else vb)[m.group(1)] = v
            continue
        for pat, rep in RENAMES:
            m = re.match(pat, k)
            if m:
                out[m.expand(rep)] = v
                break
        else:
            raise KeyError(f"BLIP-2 checkpoint: no mapping for key {k!r} (shape {tuple(v.shape)})")
    for i in qb:   # EVA attention: qkv bias = (q_bias, 0, v_bias) -- the key projection has no bias
        out[f"vision_model.encoder.layers.{i}.self_attn.qkv.bias"] = torch.cat([qb[i], torch.zeros_like(vb[i]), vb[i]])
    return out


def check_state_dict(sd: Dict[str, torch.Tensor], dims) -> None:
    """Every tensor the engine reads exists with the shape ``dims`` implies; extra keys are an error too."""
    want = expected_shapes(dims)
    missing = [k for k in want if k not in sd]
    bad = [(k, tuple(sd[k].shape), tuple(want[k])) for k in want if k in sd and tuple(sd[k].shape) != tuple(want[k])]
    extra = [k for k in sd if k not in want and "position_ids" not in k]
    if missing or bad or extra:
        raise KeyError(f"BLIP-2 checkpoint does not match the engine: {len(missing)} missing (e.g. {missing[:4]}), "
                       f"{len(bad)} wrong shapes (e.g. {bad[:3]}), {len(extra)} unexpected (e.g. {extra[:4]})")


def expected_shapes(d) -> Dict[str, Tuple[int, ...]]:
    """name -> shape of every tensor ``Blip2ITCEngine._load`` reads (HF names)."""
    D, F, H, I = d.v_hidden, d.v_inter, d.q_hidden, d.q_inter
    s: Dict[str, Tuple[int, ...]] = {
        "query_tokens": (1, d.queries, H),
        "vision_model.embeddings.class_embedding": (1, 1, D), "vision_model.embeddings.position_embedding": (1, d.tokens, D),
        "vision_model.embeddings.patch_embedding.weight": (D, 3, d.patch, d.patch), "vision_model.embeddings.patch_embedding.bias": (D,),
        "vision_model.post_layernorm.weight": (D,), "vision_model.post_layernorm.bias": (D,),
        "embeddings.word_embeddings.weight": (d.vocab, H), "embeddings.position_embeddings.weight": (d.max_pos, H),
        "qformer.layernorm.weight": (H,), "qformer.layernorm.bias": (H,),
        "vision_projection.weight": (d.proj, H), "vision_projection.bias": (d.proj,),
        "text_projection.weight": (d.proj, H), "text_projection.bias": (d.proj,),
        "itm_head.weight": (2, H), "itm_head.bias": (2,),
    }

    def lin(n, o, i):
        s[n + ".weight"] = (o, i); s[n + ".bias"] = (o,)

    def ln(n, k):
        s[n + ".weight"] = (k,); s[n + ".bias"] = (k,)

    for i in range(d.v_layers):
        p = f"vision_model.encoder.layers.{i}."
        lin(p + "self_attn.qkv", 3 * D, D); lin(p + "self_attn.projection", D, D); ln(p + "layer_norm1", D)
        lin(p + "mlp.fc1", F, D); lin(p + "mlp.fc2", D, F); ln(p + "layer_norm2", D)
    for i in range(d.q_layers):
        p = f"qformer.encoder.layer.{i}."
        for blk, kin in (("attention", H),) + ((("crossattention", D),) if i % d.cross_freq == 0 else ()):
            lin(p + blk + ".attention.query", H, H); lin(p + blk + ".attention.key", H, kin); lin(p + blk + ".attention.value", H, kin)
            lin(p + blk + ".output.dense", H, H); ln(p + blk + ".output.LayerNorm", H)
        lin(p + "intermediate.dense", I, H); lin(p + "output.dense", H, I); ln(p + "output.LayerNorm", H)
        lin(p + "intermediate_query.dense", I, H); lin(p + "output_query.dense", H, I); ln(p + "output_query.LayerNorm", H)
    return s


def load_checkpoint(path: str, dims) -> Dict[str, torch.Tensor]:
    obj = torch.load(path, map_location="cpu", weights_only=True)
    sd = obj["model"] if isinstance(obj, dict) and "model" in obj and isinstance(obj["model"], dict) else obj
    if is_lavis_layout(sd):
        sd = convert_lavis_state_dict(sd)
    check_state_dict(sd, dims)
    return sd
