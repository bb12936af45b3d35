"""Hand-written kernels under the GroundingDINO feature enhancer / decoder.

The module graph (fusion layers, deformable encoder, decoder, heads) is still HF's
``GroundingDinoForObjectDetection``; this file swaps its two hot primitives for the library's own kernels:

* every ``nn.Linear``  -> ``TcLinear``: fp16 operands on the tcgen05 GEMM (csrc/gemm_tcgen05.cu), fp32 accumulate,
  bias fused, fp32 out (the reference runs these in fp32 SIMT GEMMs: groundingdino ... nn.Linear);
* ``MultiScaleDeformableAttention`` (groundingdino's ms_deform_attn_cuda.cu / HF's grid_sample fallback)
  -> ``vlfm_msda_forward`` (csrc/gdino_ops.cu).
"""
from __future__ import annotations

import ctypes

import torch

from .. import _lib


def cast_f16(x: torch.Tensor) -> torch.Tensor:
    """fp32 contiguous -> fp16 on the library's cast kernel."""
    out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    _lib.check(_lib.load().vlfm_cast_f32_f16(x.data_ptr(), out.data_ptr(), x.numel(), _lib.stream_ptr()), "vlfm_cast_f32_f16")
    return out


class TcLinear(torch.nn.Module):
    def __init__(self, lin: torch.nn.Linear):
        super().__init__()
        self.in_features, self.out_features = lin.in_features, lin.out_features
        self.weight, self.bias = lin.weight, lin.bias                      # state_dict keys stay those of nn.Linear
        self.register_buffer("w16", lin.weight.detach().to(torch.float16).contiguous(), persistent=False)
        self.register_buffer("b32", None if lin.bias is None else lin.bias.detach().float().contiguous(), persistent=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shp = x.shape
        x2 = x.reshape(-1, self.in_features)
        if x2.dtype != torch.float16:
            x2 = cast_f16(x2.float().contiguous())
        elif not x2.is_contiguous():
            x2 = x2.contiguous()
        m = x2.shape[0]
        out = torch.empty((m, self.out_features), dtype=torch.float32, device=x.device)
        if m:
            rc = _lib.load().vlfm_gemm_f16(x2.data_ptr(), self.w16.data_ptr(), _lib.ptr(self.b32), out.data_ptr(), m, self.out_features,
                                          self.in_features, self.in_features, self.in_features, self.out_features, _lib.EPI_BIAS_F32,
                                          _lib.stream_ptr())
            _lib.check(rc, "vlfm_gemm_f16")
        return out.view(*shp[:-1], self.out_features)


class TcMSDA(torch.nn.Module):
    """Same call signature as transformers' ``MultiScaleDeformableAttention.forward``."""

    def forward(self, value, value_spatial_shapes, value_spatial_shapes_list, level_start_index, sampling_locations, attention_weights,
                im2col_step=None):
        b, s, heads, hd = value.shape
        _, q, _, levels, points, _ = sampling_locations.shape
        value = value.contiguous()
        assert value.dtype in (torch.float32, torch.float16)
        loc = sampling_locations.float().contiguous()
        attw = attention_weights.float().contiguous()
        out = torch.empty((b, q, heads * hd), dtype=torch.float32, device=value.device)
        flat = [int(v) for hw in value_spatial_shapes_list for v in hw]
        shapes = (ctypes.c_int32 * len(flat))(*flat)
        rc = _lib.load().vlfm_msda_forward(value.data_ptr(), int(value.dtype == torch.float16), loc.data_ptr(), attw.data_ptr(), out.data_ptr(),
                                          b, s, q, heads, hd, levels, points, ctypes.cast(shapes, ctypes.c_void_p), _lib.stream_ptr())
        _lib.check(rc, "vlfm_msda_forward")
        return out


def accelerate(model: torch.nn.Module, min_out: int = 16) -> dict:
    """Swap the primitives in place (model already on the GPU).  Returns counts for the log / tests."""
    from transformers.models.grounding_dino.modeling_grounding_dino import MultiScaleDeformableAttention

    n_lin = n_msda = n_skip = 0
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            if isinstance(child, torch.nn.Linear):
                if child.in_features % 8 == 0 and child.in_features >= 16 and child.out_features >= min_out and child.out_features % 4 == 0:
                    setattr(parent, name, TcLinear(child)); n_lin += 1
                else:
                    n_skip += 1
            elif isinstance(child, MultiScaleDeformableAttention):
                setattr(parent, name, TcMSDA()); n_msda += 1
    return {"linear": n_lin, "linear_kept": n_skip, "msda": n_msda}
