"""Checkpoint ingestion for GroundingDINO (vlfm/vlm/grounding_dino.py:33 ``load_model(config_path, weights_path)``).

The reference loads ``data/groundingdino_swint_ogc.pth`` -- a ``{"model": state_dict}`` file with the ORIGINAL module names of
IDEA-Research/GroundingDINO@eeba084 (``backbone.0.*``, ``transformer.*``, ``bert.*``, fused ``qkv`` / ``in_proj`` matrices).  The
engine consumes the HF ``GroundingDinoForObjectDetection`` layout (split query / key / value, ``model.*`` prefixes).  This
module holds the key-mapping table between the two and the q/k/v splits.

Status: no checkpoint exists offline, so the table is checked structurally only (tests/test_checkpoint_keymaps.py): a
synthetic original-named dict built by the inverse rules converts to EXACTLY the key set and shapes of the HF model, values
bit-identical.  The original-side names are restated from the GroundingDINO module tree (``groundingdino/models/GroundingDINO``:
``groundingdino.py``, ``transformer.py``, ``fuse_modules.py``, ``backbone/swin_transformer.py``); a name this table does not
know raises instead of being dropped.
"""
from __future__ import annotations

import re
from typing import Callable, Dict, List, Tuple

import torch

BB = "model.backbone.conv_encoder.model."

# (regex on the original name, HF replacement); first match wins.  q/k/v-fused tensors are listed in SPLITS instead.
RENAMES: List[Tuple[str, str]] = [
    # ---- Swin-T backbone (groundingdino/models/GroundingDINO/backbone/swin_transformer.py)
    (r"^backbone\.0\.patch_embed\.proj\.(weight|bias)$", BB + r"embeddings.patch_embeddings.projection.\1"),
    (r"^backbone\.0\.patch_embed\.norm\.(weight|bias)$", BB + r"embeddings.norm.\1"),
    (r"^backbone\.0\.layers\.(\d+)\.blocks\.(\d+)\.norm1\.(weight|bias)$", BB + r"encoder.layers.\1.blocks.\2.layernorm_before.\3"),
    (r"^backbone\.0\.layers\.(\d+)\.blocks\.(\d+)\.norm2\.(weight|bias)$", BB + r"encoder.layers.\1.blocks.\2.layernorm_after.\3"),
    (r"^backbone\.0\.layers\.(\d+)\.blocks\.(\d+)\.attn\.relative_position_bias_table$", BB + r"encoder.layers.\1.blocks.\2.attention.self.relative_position_bias_table"),
    (r"^backbone\.0\.layers\.(\d+)\.blocks\.(\d+)\.attn\.relative_position_index$", BB + r"encoder.layers.\1.blocks.\2.attention.self.relative_position_index"),
    (r"^backbone\.0\.layers\.(\d+)\.blocks\.(\d+)\.attn\.proj\.(weight|bias)$", BB + r"encoder.layers.\1.blocks.\2.attention.output.dense.\3"),
    (r"^backbone\.0\.layers\.(\d+)\.blocks\.(\d+)\.mlp\.fc1\.(weight|bias)$", BB + r"encoder.layers.\1.blocks.\2.intermediate.dense.\3"),
    (r"^backbone\.0\.layers\.(\d+)\.blocks\.(\d+)\.mlp\.fc2\.(weight|bias)$", BB + r"encoder.layers.\1.blocks.\2.output.dense.\3"),
    (r"^backbone\.0\.layers\.(\d+)\.downsample\.reduction\.weight$", BB + r"encoder.layers.\1.downsample.reduction.weight"),
    (r"^backbone\.0\.layers\.(\d+)\.downsample\.norm\.(weight|bias)$", BB + r"encoder.layers.\1.downsample.norm.\2"),
    (r"^backbone\.0\.norm1\.(weight|bias)$", BB + r"hidden_states_norms.stage2.\1"),      # out_indices (1, 2, 3) -> stage2..4
    (r"^backbone\.0\.norm2\.(weight|bias)$", BB + r"hidden_states_norms.stage3.\1"),
    (r"^backbone\.0\.norm3\.(weight|bias)$", BB + r"hidden_states_norms.stage4.\1"),
    # ---- neck, embeddings, text tower (groundingdino.py)
    (r"^input_proj\.(\d+)\.(\d+)\.(weight|bias)$", r"model.input_proj_vision.\1.\2.\3"),
    (r"^transformer\.level_embed$", r"model.level_embed"),
    (r"^transformer\.tgt_embed\.weight$", r"model.query_position_embeddings.weight"),
    (r"^feat_map\.(weight|bias)$", r"model.text_projection.\1"),
    (r"^bert\.(embeddings\.(?:word|position|token_type)_embeddings\.weight|embeddings\.LayerNorm\.(?:weight|bias)|encoder\..*)$", r"model.text_backbone.\1"),
    # ---- encoder: deformable layers, text enhancer, fusion (transformer.py, fuse_modules.py)
    (r"^transformer\.encoder\.layers\.(\d+)\.self_attn\.(sampling_offsets|attention_weights|value_proj|output_proj)\.(weight|bias)$",
     r"model.encoder.layers.\1.deformable_layer.self_attn.\2.\3"),
    (r"^transformer\.encoder\.layers\.(\d+)\.norm1\.(weight|bias)$", r"model.encoder.layers.\1.deformable_layer.self_attn_layer_norm.\2"),
    (r"^transformer\.encoder\.layers\.(\d+)\.linear1\.(weight|bias)$", r"model.encoder.layers.\1.deformable_layer.fc1.\2"),
    (r"^transformer\.encoder\.layers\.(\d+)\.linear2\.(weight|bias)$", r"model.encoder.layers.\1.deformable_layer.fc2.\2"),
    (r"^transformer\.encoder\.layers\.(\d+)\.norm2\.(weight|bias)$", r"model.encoder.layers.\1.deformable_layer.final_layer_norm.\2"),
    (r"^transformer\.encoder\.text_layers\.(\d+)\.self_attn\.out_proj\.(weight|bias)$", r"model.encoder.layers.\1.text_enhancer_layer.self_attn.out_proj.\2"),
    (r"^transformer\.encoder\.text_layers\.(\d+)\.linear1\.(weight|bias)$", r"model.encoder.layers.\1.text_enhancer_layer.fc1.\2"),
    (r"^transformer\.encoder\.text_layers\.(\d+)\.linear2\.(weight|bias)$", r"model.encoder.layers.\1.text_enhancer_layer.fc2.\2"),
    (r"^transformer\.encoder\.text_layers\.(\d+)\.norm1\.(weight|bias)$", r"model.encoder.layers.\1.text_enhancer_layer.layer_norm_before.\2"),
    (r"^transformer\.encoder\.text_layers\.(\d+)\.norm2\.(weight|bias)$", r"model.encoder.layers.\1.text_enhancer_layer.layer_norm_after.\2"),
    (r"^transformer\.encoder\.fusion_layers\.(\d+)\.gamma_v$", r"model.encoder.layers.\1.fusion_layer.vision_param"),
    (r"^transformer\.encoder\.fusion_layers\.(\d+)\.gamma_l$", r"model.encoder.layers.\1.fusion_layer.text_param"),
    (r"^transformer\.encoder\.fusion_layers\.(\d+)\.layer_norm_v\.(weight|bias)$", r"model.encoder.layers.\1.fusion_layer.layer_norm_vision.\2"),
    (r"^transformer\.encoder\.fusion_layers\.(\d+)\.layer_norm_l\.(weight|bias)$", r"model.encoder.layers.\1.fusion_layer.layer_norm_text.\2"),
    (r"^transformer\.encoder\.fusion_layers\.(\d+)\.attn\.v_proj\.(weight|bias)$", r"model.encoder.layers.\1.fusion_layer.attn.vision_proj.\2"),
    (r"^transformer\.encoder\.fusion_layers\.(\d+)\.attn\.l_proj\.(weight|bias)$", r"model.encoder.layers.\1.fusion_layer.attn.text_proj.\2"),
    (r"^transformer\.encoder\.fusion_layers\.(\d+)\.attn\.values_v_proj\.(weight|bias)$", r"model.encoder.layers.\1.fusion_layer.attn.values_vision_proj.\2"),
    (r"^transformer\.encoder\.fusion_layers\.(\d+)\.attn\.values_l_proj\.(weight|bias)$", r"model.encoder.layers.\1.fusion_layer.attn.values_text_proj.\2"),
    (r"^transformer\.encoder\.fusion_layers\.(\d+)\.attn\.out_v_proj\.(weight|bias)$", r"model.encoder.layers.\1.fusion_layer.attn.out_vision_proj.\2"),
    (r"^transformer\.encoder\.fusion_layers\.(\d+)\.attn\.out_l_proj\.(weight|bias)$", r"model.encoder.layers.\1.fusion_layer.attn.out_text_proj.\2"),
    # ---- two-stage proposal heads
    (r"^transformer\.enc_output\.(weight|bias)$", r"model.enc_output.\1"),
    (r"^transformer\.enc_output_norm\.(weight|bias)$", r"model.enc_output_norm.\1"),
    (r"^transformer\.enc_out_bbox_embed\.layers\.(\d+)\.(weight|bias)$", r"model.encoder_output_bbox_embed.layers.\1.\2"),
    # ---- decoder
    (r"^transformer\.decoder\.layers\.(\d+)\.self_attn\.out_proj\.(weight|bias)$", r"model.decoder.layers.\1.self_attn.out_proj.\2"),
    (r"^transformer\.decoder\.layers\.(\d+)\.norm2\.(weight|bias)$", r"model.decoder.layers.\1.self_attn_layer_norm.\2"),
    (r"^transformer\.decoder\.layers\.(\d+)\.ca_text\.out_proj\.(weight|bias)$", r"model.decoder.layers.\1.encoder_attn_text.out_proj.\2"),
    (r"^transformer\.decoder\.layers\.(\d+)\.catext_norm\.(weight|bias)$", r"model.decoder.layers.\1.encoder_attn_text_layer_norm.\2"),
    (r"^transformer\.decoder\.layers\.(\d+)\.cross_attn\.(sampling_offsets|attention_weights|value_proj|output_proj)\.(weight|bias)$",
     r"model.decoder.layers.\1.encoder_attn.\2.\3"),
    (r"^transformer\.decoder\.layers\.(\d+)\.norm1\.(weight|bias)$", r"model.decoder.layers.\1.encoder_attn_layer_norm.\2"),
    (r"^transformer\.decoder\.layers\.(\d+)\.linear1\.(weight|bias)$", r"model.decoder.layers.\1.fc1.\2"),
    (r"^transformer\.decoder\.layers\.(\d+)\.linear2\.(weight|bias)$", r"model.decoder.layers.\1.fc2.\2"),
    (r"^transformer\.decoder\.layers\.(\d+)\.norm3\.(weight|bias)$", r"model.decoder.layers.\1.final_layer_norm.\2"),
    (r"^transformer\.decoder\.norm\.(weight|bias)$", r"model.decoder.layer_norm.\1"),
    (r"^transformer\.decoder\.ref_point_head\.layers\.(\d+)\.(weight|bias)$", r"model.decoder.reference_points_head.layers.\1.\2"),
    (r"^transformer\.decoder\.bbox_embed\.(\d+)\.layers\.(\d+)\.(weight|bias)$", r"model.decoder.bbox_embed.\1.layers.\2.\3"),
    (r"^bbox_embed\.(\d+)\.layers\.(\d+)\.(weight|bias)$", r"bbox_embed.\1.layers.\2.\3"),
]

# fused [3d, ...] tensors -> query / key / value thirds
SPLITS: List[Tuple[str, str]] = [
    (r"^backbone\.0\.layers\.(\d+)\.blocks\.(\d+)\.attn\.qkv\.(weight|bias)$", BB + r"encoder.layers.\1.blocks.\2.attention.self.{part}.\3"),
    (r"^transformer\.encoder\.text_layers\.(\d+)\.self_attn\.in_proj_(weight|bias)$", r"model.encoder.layers.\1.text_enhancer_layer.self_attn.{part}.\2"),
    (r"^transformer\.decoder\.layers\.(\d+)\.self_attn\.in_proj_(weight|bias)$", r"model.decoder.layers.\1.self_attn.{part}.\2"),
    (r"^transformer\.decoder\.layers\.(\d+)\.ca_text\.in_proj_(weight|bias)$", r"model.decoder.layers.\1.encoder_attn_text.{part}.\2"),
]

# present in the original file, not parameters of the forward the reference runs
IGNORED = [r"^bert\.embeddings\.position_ids$", r"^label_enc\.weight$", r"^bert\.pooler\..*$"]


def is_original_layout(sd: Dict[str, torch.Tensor]) -> bool:
    return any(k.startswith(("backbone.0.", "transformer.", "bert.")) for k in sd)


def convert_groundingdino_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """original GroundingDINO names -> HF GroundingDinoForObjectDetection names.  Unknown names raise."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        k = k[7:] if k.startswith("module.") else k
        if any(re.match(p, k) for p in IGNORED):
            continue
        done = False
        for pat, rep in SPLITS:
            m = re.match(pat, k)
            if m:
                assert v.shape[0] % 3 == 0, (k, tuple(v.shape))
                d = v.shape[0] // 3
                for i, part in enumerate(("query", "key", "value")):
                    out[m.expand(rep).replace("{part}", part)] = v[i * d : (i + 1) * d].clone()
                done = True
                break
        if done:
            continue
        for pat, rep in RENAMES:
            m = re.match(pat, k)
            if m:
                out[m.expand(rep)] = v
                done = True
                break
        if not done:
            raise KeyError(f"GroundingDINO checkpoint: no mapping for key {k!r} (shape {tuple(v.shape)})")
    # the decoder's box heads are the top-level ones (shared modules): mirror whichever side the file carries
    for k in list(out):
        if k.startswith("bbox_embed."):
            out.setdefault("model.decoder." + k, out[k])
        elif k.startswith("model.decoder.bbox_embed."):
            out.setdefault(k[len("model.decoder."):], out[k])
    return out


def load_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    """torch.load a GroundingDINO checkpoint (original ``{"model": ...}`` file or an HF-layout state dict) -> HF layout."""
    obj = torch.load(path, map_location="cpu", weights_only=True)
    sd = obj["model"] if isinstance(obj, dict) and "model" in obj and isinstance(obj["model"], dict) else obj
    return convert_groundingdino_state_dict(sd) if is_original_layout(sd) else sd
