"""tcgen05 GEMM vs a plain PyTorch fp32 reference of the same op (fp16 inputs, fp32
accumulation): tolerance 2e-3 relative to the output scale."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_gemm(a, w, bias, epi, resid=None):
    from vlfm_b200.vlm.dense import gemm_f16

    return gemm_f16(a, w, bias, epi, resid)


@pytest.mark.parametrize("M,N,K", [(257, 4224, 1408), (257, 1408, 6144), (128, 128, 64), (1, 256, 768),
                                   (300, 1408, 592), (32, 768, 768), (1000, 6144, 1408), (257, 9216, 1408),
                                   (77, 264, 72), (8224, 4224, 1408), (2100, 2312, 520)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_matches_torch(M, N, K, epi):
    if epi in (1, 2) and (M, N, K) not in [(257, 4224, 1408), (300, 1408, 592), (77, 264, 72), (32, 768, 768), (8224, 4224, 1408), (2100, 2312, 520)]:
        pytest.skip("epilogue variants checked on a subset of shapes")
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + epi)
    a = (torch.randn(M, K, generator=g) * 0.5).half().cuda()
    w = (torch.randn(N, K, generator=g) * 0.05).half().cuda()
    bias = torch.randn(N, generator=g).float().cuda()
    ref = a.float() @ w.float().t() + bias
    if epi == 1:
        ref = torch.nn.functional.gelu(ref)
    resid = None
    if epi == 2:
        resid = torch.randn(M, N, generator=g).float().cuda()
        ref = ref + resid
    out = run_gemm(a, w, bias, epi, resid)
    torch.cuda.synchronize()
    got = out.float()
    scale = ref.abs().max().item()
    tol = 2e-3 * scale if epi in (0, 1) else 2e-4 * scale
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= tol, f"max err {(got - ref).abs().max().item()} vs tol {tol}"


@pytest.mark.parametrize("M,N,K", [(257, 4224, 1408), (257, 1408, 6144), (129, 96, 72), (258, 1408, 1408), (386, 264, 200), (257, 6144, 1408)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3, 4])
def test_tail_rows_on_cuda_cores(M, N, K, epi):
    """M = 128*q + r, r <= 2: the r tail rows are computed by the last row tile's CTAs on CUDA cores from the swizzled W stages
    (incl. split-K atomics for the residual epilogue, K and N tails).  Same tolerance as the tensor-core rows; the tail rows
    are also checked on their own."""
    g = torch.Generator(device="cpu").manual_seed(M * 11 + N * 5 + K + epi)
    a = (torch.randn(M, K, generator=g) * 0.5).half().cuda()
    w = (torch.randn(N, K, generator=g) * 0.05).half().cuda()
    bias = torch.randn(N, generator=g).float().cuda()
    ref = a.float() @ w.float().t() + bias
    if epi == 1:
        ref = torch.nn.functional.gelu(ref)
    if epi == 4:
        ref = torch.relu(ref)
    resid = None
    if epi == 2:
        resid = torch.randn(M, N, generator=g).float().cuda()
        ref = ref + resid
    out = run_gemm(a, w, bias, epi, resid)
    torch.cuda.synchronize()
    got = out.float()
    scale = ref.abs().max().item()
    tol = 2e-3 * scale if epi in (0, 1, 4) else 2e-4 * scale
    q = (M // 128) * 128
    assert torch.isfinite(got).all()
    assert (got[q:] - ref[q:]).abs().max().item() <= tol, f"tail rows: max err {(got[q:] - ref[q:]).abs().max().item()} vs tol {tol}"
    assert (got - ref).abs().max().item() <= tol
