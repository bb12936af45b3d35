"""Checkpoint ingestion (vlfm/vlm/blip2itm.py:29-34, vlfm/vlm/grounding_dino.py:33): the key-mapping tables from the
reference's checkpoint layouts (lavis / GroundingDINO original names) to the layout the engines consume, dry-run on
synthetic dicts: converted key set == model key set, shapes equal, values bit-identical, unknown keys raise.  (No real
checkpoint exists offline; the original-side names are restated from the two packages' module trees.)"""
import re

import pytest
import torch

from vlfm_b200.vlm import blip2_weights as bw
from vlfm_b200.vlm import gdino_weights as gw
from vlfm_b200.vlm.blip2_config import SMALL, TINY, Blip2Dims, random_state_dict


def _to_lavis(sd):
    """HF -> lavis names, written independently of the converter (plain string edits)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("vision_model.encoder.layers."):
            k2 = k.replace("vision_model.encoder.layers.", "visual_encoder.blocks.")
            k2 = k2.replace(".layer_norm1.", ".norm1.").replace(".layer_norm2.", ".norm2.").replace(".self_attn.projection.", ".attn.proj.")
            if k2.endswith(".self_attn.qkv.weight"):
                out[k2.replace(".self_attn.qkv.weight", ".attn.qkv.weight")] = v
            elif k2.endswith(".self_attn.qkv.bias"):
                d = v.shape[0] // 3
                assert float(v[d:2 * d].abs().max()) == 0.0          # EVA: no key bias
                out[k2.replace(".self_attn.qkv.bias", ".attn.q_bias")] = v[:d].clone()
                out[k2.replace(".self_attn.qkv.bias", ".attn.v_bias")] = v[2 * d:].clone()
            else:
                out[k2] = v
        elif k == "vision_model.embeddings.class_embedding":
            out["visual_encoder.cls_token"] = v
        elif k == "vision_model.embeddings.position_embedding":
            out["visual_encoder.pos_embed"] = v
        elif k.startswith("vision_model.embeddings.patch_embedding."):
            out[k.replace("vision_model.embeddings.patch_embedding.", "visual_encoder.patch_embed.proj.")] = v
        elif k.startswith("vision_model.post_layernorm."):
            out[k.replace("vision_model.post_layernorm.", "ln_vision.")] = v
        elif k.startswith("qformer.layernorm."):
            out[k.replace("qformer.layernorm.", "Qformer.bert.embeddings.LayerNorm.")] = v
        elif k.startswith("embeddings."):
            out["Qformer.bert." + k] = v
        elif k.startswith("qformer.encoder."):
            out["Qformer.bert." + k[len("qformer."):].replace("attention.attention.", "attention.self.")] = v
        elif k.startswith("vision_projection."):
            out[k.replace("vision_projection.", "vision_proj.")] = v
        elif k.startswith("text_projection."):
            out[k.replace("text_projection.", "text_proj.")] = v
        else:
            out[k] = v
    out["temp"] = torch.tensor(0.07)
    out["Qformer.cls.predictions.bias"] = torch.zeros(4)
    out["Qformer.bert.embeddings.position_ids"] = torch.arange(8)[None]
    return out


@pytest.mark.parametrize("dims", [TINY, SMALL])
def test_lavis_names_convert_to_the_engine_layout(dims):
    sd = random_state_dict(dims, 5)
    lav = _to_lavis(sd)
    assert bw.is_lavis_layout(lav) and not bw.is_lavis_layout(sd)
    back = bw.convert_lavis_state_dict(lav)
    assert set(back) == set(sd)
    for k in sd:
        assert back[k].shape == sd[k].shape and torch.equal(back[k], sd[k]), k
    bw.check_state_dict(back, dims)
    with pytest.raises(KeyError):
        bw.convert_lavis_state_dict({**lav, "visual_encoder.blocks.0.attn.rel_pos_bias": torch.zeros(1)})
    broken = dict(back)
    broken.pop("qformer.encoder.layer.1.output_query.dense.weight")
    with pytest.raises(KeyError):
        bw.check_state_dict(broken, dims)


def test_expected_shapes_cover_the_full_size_model():
    d = Blip2Dims()
    shapes = bw.expected_shapes(d)
    n = sum(int(torch.tensor(s).prod()) for s in shapes.values())
    assert 1.10e9 < n < 1.25e9                       # 1.17 B parameters
    assert shapes["vision_model.encoder.layers.38.self_attn.qkv.weight"] == (3 * 1408, 1408)
    assert shapes["qformer.encoder.layer.10.crossattention.attention.key.weight"] == (768, 1408)
    assert "qformer.encoder.layer.11.crossattention.attention.key.weight" not in shapes


def _hf_gdino_meta():
    from transformers import GroundingDinoConfig, GroundingDinoForObjectDetection

    with torch.device("meta"):
        m = GroundingDinoForObjectDetection(GroundingDinoConfig())
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


def _to_original(hf_shapes):
    """HF -> original GroundingDINO names (independent of the converter): returns name -> tensor with recognisable values."""
    g = torch.Generator().manual_seed(0)
    hf = {k: torch.randn(s, generator=g) if len(s) else torch.randn((), generator=g) for k, s in hf_shapes.items()}
    out, fused = {}, {}
    BB = "model.backbone.conv_encoder.model."
    for k, v in hf.items():
        if k.startswith("model.decoder.bbox_embed."):
            continue                                   # shared with the top-level bbox_embed in the original file
        o = None
        if k.startswith(BB):
            r = k[len(BB):]
            r = r.replace("embeddings.patch_embeddings.projection.", "patch_embed.proj.").replace("embeddings.norm.", "patch_embed.norm.")
            r = r.replace("encoder.layers.", "layers.").replace(".layernorm_before.", ".norm1.").replace(".layernorm_after.", ".norm2.")
            r = r.replace(".attention.self.relative_position", ".attn.relative_position").replace(".attention.output.dense.", ".attn.proj.")
            r = r.replace(".intermediate.dense.", ".mlp.fc1.").replace(".output.dense.", ".mlp.fc2.")
            m = re.match(r"hidden_states_norms\.stage(\d)\.(weight|bias)", r)
            if m:
                r = f"norm{int(m.group(1)) - 1}.{m.group(2)}"
            m = re.match(r"(layers\.\d+\.blocks\.\d+)\.attention\.self\.(query|key|value)\.(weight|bias)", r)
            if m:
                fused.setdefault(("backbone.0." + m.group(1) + ".attn.qkv." + m.group(3)), {})[m.group(2)] = v
                continue
            o = "backbone.0." + r
        elif k.startswith("model.text_backbone."):
            o = "bert." + k[len("model.text_backbone."):]
        elif k.startswith("model.input_proj_vision."):
            o = "input_proj." + k[len("model.input_proj_vision."):]
        elif k.startswith("model.text_projection."):
            o = "feat_map." + k[len("model.text_projection."):]
        elif k == "model.level_embed":
            o = "transformer.level_embed"
        elif k == "model.query_position_embeddings.weight":
            o = "transformer.tgt_embed.weight"
        elif k.startswith("model.enc_output"):
            o = "transformer." + k[len("model."):]
        elif k.startswith("model.encoder_output_bbox_embed."):
            o = "transformer.enc_out_bbox_embed." + k[len("model.encoder_output_bbox_embed."):]
        elif k.startswith("bbox_embed."):
            o = k
        elif k.startswith("model.encoder.layers."):
            m = re.match(r"model\.encoder\.layers\.(\d+)\.(deformable_layer|text_enhancer_layer|fusion_layer)\.(.*)", k)
            i, kind, r = m.groups()
            if kind == "deformable_layer":
                r = r.replace("self_attn_layer_norm.", "norm1.").replace("final_layer_norm.", "norm2.").replace("fc1.", "linear1.").replace("fc2.", "linear2.")
                o = f"transformer.encoder.layers.{i}.{r}"
            elif kind == "text_enhancer_layer":
                mm = re.match(r"self_attn\.(query|key|value)\.(weight|bias)", r)
                if mm:
                    fused.setdefault(f"transformer.encoder.text_layers.{i}.self_attn.in_proj_{mm.group(2)}", {})[mm.group(1)] = v
                    continue
                r = r.replace("layer_norm_before.", "norm1.").replace("layer_norm_after.", "norm2.").replace("fc1.", "linear1.").replace("fc2.", "linear2.")
                o = f"transformer.encoder.text_layers.{i}.{r}"
            else:
                r = r.replace("vision_param", "gamma_v").replace("text_param", "gamma_l").replace("layer_norm_vision.", "layer_norm_v.").replace("layer_norm_text.", "layer_norm_l.")
                r = r.replace("attn.values_vision_proj.", "attn.values_v_proj.").replace("attn.values_text_proj.", "attn.values_l_proj.")
                r = r.replace("attn.out_vision_proj.", "attn.out_v_proj.").replace("attn.out_text_proj.", "attn.out_l_proj.")
                r = r.replace("attn.vision_proj.", "attn.v_proj.").replace("attn.text_proj.", "attn.l_proj.")
                o = f"transformer.encoder.fusion_layers.{i}.{r}"
        elif k.startswith("model.decoder."):
            r = k[len("model.decoder."):]
            m = re.match(r"layers\.(\d+)\.(self_attn|encoder_attn_text)\.(query|key|value)\.(weight|bias)", r)
            if m:
                mod = "self_attn" if m.group(2) == "self_attn" else "ca_text"
                fused.setdefault(f"transformer.decoder.layers.{m.group(1)}.{mod}.in_proj_{m.group(4)}", {})[m.group(3)] = v
                continue
            r = r.replace("encoder_attn_text_layer_norm.", "catext_norm.").replace("encoder_attn_text.", "ca_text.")
            r = r.replace("self_attn_layer_norm.", "norm2.").replace("encoder_attn_layer_norm.", "norm1.").replace("final_layer_norm.", "norm3.")
            r = r.replace("encoder_attn.", "cross_attn.").replace(".fc1.", ".linear1.").replace(".fc2.", ".linear2.")
            r = r.replace("reference_points_head.", "ref_point_head.")
            r = "norm." + r[len("layer_norm."):] if r.startswith("layer_norm.") else r
            o = "transformer.decoder." + r
        assert o is not None, k
        out[o] = v
    for name, parts in fused.items():
        out[name] = torch.cat([parts["query"], parts["key"], parts["value"]], 0)
    out["bert.embeddings.position_ids"] = torch.arange(512)[None]
    out["label_enc.weight"] = torch.zeros(92, 256)
    return hf, out


def test_groundingdino_original_names_convert_to_the_hf_layout():
    shapes = _hf_gdino_meta()
    hf, orig = _to_original(shapes)
    assert gw.is_original_layout(orig) and not gw.is_original_layout(hf)
    assert any(k.endswith("attn.qkv.weight") for k in orig) and "transformer.decoder.layers.5.ca_text.in_proj_weight" in orig
    conv = gw.convert_groundingdino_state_dict(orig)
    want = {k for k in shapes if "position_ids" not in k}
    assert set(conv) == want, (sorted(set(conv) - want)[:5], sorted(want - set(conv))[:5])
    for k in want:
        assert tuple(conv[k].shape) == shapes[k], k
        src = hf[k[len("model.decoder."):]] if k.startswith("model.decoder.bbox_embed.") else hf[k]
        assert torch.equal(conv[k], src), k
    with pytest.raises(KeyError):
        gw.convert_groundingdino_state_dict({**orig, "transformer.encoder.layers.0.unknown.weight": torch.zeros(1)})
