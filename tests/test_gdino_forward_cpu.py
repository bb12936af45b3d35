"""The model-level GroundingDINO forward of vlfm_b200 (vlm/gdino_forward.py: neck as row GEMM + GroupNorm, cached shape / caption
constants, proposal scoring, top-900 selection, box head on the selected rows only, last-layer heads only) reproduces HF's
``GroundingDinoForObjectDetection`` forward.  Checked on the CPU with a torch implementation of the kernel interface (the
product uses the C-ABI kernels, vlm/gdino_ops.py; their numerics are tested on the GPU)."""
import numpy as np
import pytest
import torch

from oracle.gdino_oracle import GdinoOracle, preprocess
from vlfm_b200.utils.synthetic import make_rgb
from vlfm_b200.vlm.gdino_forward import GdinoForward


class TorchOps:
    """fp32 torch statement of every method of vlm/gdino_ops.py::LibOps (test infrastructure)."""

    def weight(self, w):
        return w.float().contiguous()

    def to_operand(self, x):
        return x.float()

    def linear_operand(self, a, w, bias):
        return torch.nn.functional.linear(a, w, bias)

    def linear(self, x, w, bias, relu=False):
        y = torch.nn.functional.linear(x.float(), w, bias)
        return torch.relu(y) if relu else y

    def layernorm(self, x, g, b, eps):
        return torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, eps)

    def im2col3x3s2(self, rows, B, h, w):
        C = rows.shape[1]
        x = rows.view(B, h, w, C).permute(0, 3, 1, 2)
        col = torch.nn.functional.unfold(x, kernel_size=3, stride=2, padding=1)               # [B, C*9, L], (c, ky, kx) order
        L = col.shape[-1]
        return col.view(B, C, 9, L).permute(0, 3, 2, 1).reshape(B * L, 9 * C)                   # (ky, kx, c) order

    def groupnorm_rows(self, y, B, HW, C, groups, g, b, eps, out, row_off, S):
        x = y.view(B, HW, C).permute(0, 2, 1)
        out[:, row_off:row_off + HW, :] = torch.nn.functional.group_norm(x, groups, g, b, eps).permute(0, 2, 1)

    def mask_rows(self, x, valid):
        return x * valid.view(-1, 1).float()

    def proposal_scores(self, q, text, B, S, T):
        return (q.view(B, S, -1) @ text.view(B, T, -1).transpose(1, 2)).max(-1)[0]

    def topk_rows(self, scores, k):
        return torch.topk(scores, k, dim=1)[1]

    def gather_rows(self, src, idx):
        return torch.gather(src, 1, idx.unsqueeze(-1).repeat(1, 1, src.shape[-1]))

    def decoder_query_pos(self, ref, valid_ratios, dim_t):
        from transformers.models.grounding_dino.modeling_grounding_dino import get_sine_pos_embed

        ref_in = ref[:, :, None] * torch.cat([valid_ratios, valid_ratios], -1)[:, None]
        emb = get_sine_pos_embed(ref_in[:, :, 0, :], num_pos_feats=dim_t.shape[0])
        return ref_in, emb.reshape(-1, emb.shape[-1])

    def box_finish(self, delta, ref):
        return (delta + torch.special.logit(ref, eps=1e-5)).sigmoid()

    def contrastive_sigmoid(self, hs, text, L):
        out = torch.zeros(hs.shape[0], hs.shape[1], L)
        out[..., : text.shape[1]] = (hs @ text.transpose(1, 2)).sigmoid()
        return out


@pytest.mark.parametrize("hw,B", [((256, 320), 1), ((225, 318), 2)])
def test_own_forward_equals_hf_forward(hw, B):
    orc = GdinoOracle(0)
    sd = orc.model.state_dict()
    for k in ("model.decoder.layer_norm.weight", "model.decoder.layer_norm.bias"):       # un-saturated class scores
        sd[k].mul_(0.1)
    model = orc.model
    fw = GdinoForward(model, TorchOps())
    rng = np.random.default_rng(3)
    imgs = [make_rgb(rng, *hw) for _ in range(B)]
    ids = [101, 4010, 1012, 2711, 1012, 3899, 1012, 102]
    px = torch.stack([preprocess(i) for i in imgs])
    cap = {}
    # the module graph's initial reference points come from a hook; the own forward (which sequences the decoder layers itself) keeps its own
    handle = model.model.decoder.register_forward_hook(lambda mod, args, kwargs, out: cap.setdefault("refs", []).append(kwargs["reference_points"]),
                                                       with_kwargs=True)
    with torch.inference_mode():
        ref = model(pixel_values=px, input_ids=torch.tensor([ids] * B), token_type_ids=torch.zeros(B, len(ids), dtype=torch.long),
                    attention_mask=torch.ones(B, len(ids), dtype=torch.long), pixel_mask=torch.ones(B, *hw, dtype=torch.long))
        bb = model.model.backbone.conv_encoder.model(px, return_dict=True).feature_maps                 # NCHW
        feats = [(f.permute(0, 2, 3, 1).reshape(-1, f.shape[1]).contiguous(), f.shape[2], f.shape[3]) for f in bb]
        logits, boxes = fw.forward_features(feats, B, hw[0], hw[1], ids)
    handle.remove()
    ref_l, ref_b = ref.logits.sigmoid(), ref.pred_boxes
    assert logits.shape == ref_l.shape and boxes.shape == ref_b.shape
    # the 900 selected proposals are the same SET; near-tied selection scores (random weights) may swap neighbours between the two
    # computations, and a swapped proposal meets a different learned query: compare the rows whose proposal is in the same place
    r0, r1 = cap["refs"][0], fw.last_reference_points
    d = (r0[:, :, None, :] - r1[:, None, :, :]).abs().sum(-1)
    assert float(d.min(2)[0].max()) <= 1e-5 and float(d.min(1)[0].max()) <= 1e-5, "the selected proposal sets differ"
    same = (r0 - r1).abs().sum(-1) < 1e-6
    assert float(same.float().mean()) >= 0.98
    assert float((logits - ref_l)[same].abs().max()) <= 2e-4, float((logits - ref_l)[same].abs().max())
    assert float((boxes - ref_b)[same].abs().max()) <= 2e-4, float((boxes - ref_b)[same].abs().max())
