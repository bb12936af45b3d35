"""Attention and LayerNorm kernels vs plain PyTorch fp32 references of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,heads,Nq,Nk,hd", [
    (1, 16, 257, 257, 88),    # ViT-g, batch 1: 32-row blocks, split keys
    (8, 16, 257, 257, 88),    # large batch: 64-row blocks
    (4, 12, 32, 257, 64),     # Q-Former cross attention
    (64, 12, 32, 32, 64),     # Q-Former self attention, large batch
    (1, 12, 7, 7, 64),        # text tower, short prompt
    (3, 2, 17, 16, 16),       # tiny dims
])
def test_attention_matches_torch(B, heads, Nq, Nk, hd):
    from vlfm_b200.vlm.dense import attention_f16

    g = torch.Generator(device="cpu").manual_seed(B * 131 + Nq)
    D = heads * hd
    qkv_q = torch.randn(B * Nq, D, generator=g).half().cuda()
    kv = torch.randn(B * Nk, 2 * D, generator=g).half().cuda()      # k and v as strided column slices
    scale = hd ** -0.5
    out = attention_f16(qkv_q, kv[:, :D], kv[:, D:], B, heads, Nq, Nk, hd, scale).float()
    q = qkv_q.float().view(B, Nq, heads, hd).transpose(1, 2)
    k = kv[:, :D].float().reshape(B, Nk, heads, hd).transpose(1, 2)
    v = kv[:, D:].float().reshape(B, Nk, heads, hd).transpose(1, 2)
    ref = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1) @ v
    ref = ref.transpose(1, 2).reshape(B * Nq, D)
    err = (out - ref).abs().max().item()
    assert err <= 4e-3, err       # fp16 P and V operands, fp32 accumulation


@pytest.mark.parametrize("rows,D,eps", [(257, 1408, 1e-6), (32, 768, 1e-12), (19200, 96, 1e-5), (5, 1536, 1e-5), (100, 64, 1e-5)])
def test_layernorm_matches_torch(rows, D, eps):
    from vlfm_b200.vlm.dense import layernorm

    g = torch.Generator(device="cpu").manual_seed(rows + D)
    x = (torch.randn(rows, D, generator=g) * 3 + 0.7).cuda()
    gamma = (1 + 0.1 * torch.randn(D, generator=g)).cuda()
    beta = (0.1 * torch.randn(D, generator=g)).cuda()
    o16, o32 = layernorm(x, gamma, beta, eps, True, True)
    ref = torch.nn.functional.layer_norm(x, (D,), gamma, beta, eps)
    assert (o32 - ref).abs().max().item() <= 2e-5
    assert (o16.float() - ref).abs().max().item() <= 4e-3
