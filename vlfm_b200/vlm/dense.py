"""Thin torch-tensor wrappers over the dense C-ABI entry points (csrc/gemm_tcgen05.cu ...)."""
from __future__ import annotations

from typing import Optional

import torch

from .. import _lib


def gemm_f16(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], epilogue: int,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = epilogue(a[M,K] @ w[N,K]^T + bias).  For epilogue 2 ``out`` is the fp32
    residual stream updated in place."""
    lib = _lib.load()
    M, K = a.shape
    N = w.shape[0]
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and a.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        assert epilogue != _lib.EPI_BIAS_RESID_F32
        out = torch.empty((M, N), dtype=torch.float32 if epilogue == _lib.EPI_BIAS_F32 else torch.float16, device=a.device)
    rc = lib.vlfm_gemm_f16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), M, N, K, a.stride(0), w.stride(0),
                           out.stride(0), epilogue, _lib.stream_ptr())
    _lib.check(rc, "vlfm_gemm_f16")
    return out
