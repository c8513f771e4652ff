"""vcof_vae_attn — the fused single-head d = 384 attention of the VAE's AttentionBlock (wan_vae.py:244-266) — against
fp32 torch math on the GPU and against the unfused three-kernel path it replaces.

Tolerance: relative Frobenius < 6e-3 against fp32 softmax(q k^T / sqrt(384)) v on bf16 inputs (P is rounded to bf16
before the PV product, the output to bf16: the same two roundings as the DiT attention kernel, whose tests use 1.5e-2)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
C = 384


def _ref(qkv):
    q, k, v = qkv[..., :C].float(), qkv[..., C:2 * C].float(), qkv[..., 2 * C:].float()
    return torch.softmax(q @ k.transpose(1, 2) / math.sqrt(C), dim=-1) @ v


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


@pytest.mark.parametrize("T,N", [(1, 64), (2, 200), (3, 1000), (1, 128), (2, 3600)])
def test_fused_vae_attention_matches_fp32(T, N):
    """Ragged token counts (last key block of 8 / 40 keys, last query tile of 72 / 104 rows), several frames per launch."""
    from videocof_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(T * 1000 + N)
    qkv = torch.randn(T, N, 3 * C, generator=g, device="cuda").bfloat16()
    out = ops.vae_attn(qkv, C)
    assert out.shape == (T, N, C) and out.dtype == torch.bfloat16
    assert torch.isfinite(out.float()).all()
    assert _rel(out, _ref(qkv)) < 6e-3


def test_fused_vae_attention_rescale_path_and_peaked_rows():
    """Keys whose norm grows block by block (the lazy rescale fires repeatedly, per row) and query rows scaled so that
    the softmax is nearly one-hot."""
    from videocof_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    T, N = 2, 777
    qkv = torch.randn(T, N, 3 * C, generator=g, device="cuda")
    qkv[..., C:2 * C] *= (1.0 + 0.4 * (torch.arange(N, device="cuda") // 64))[None, :, None]
    qkv[:, ::7, :C] *= 5
    qkv = qkv.bfloat16()
    out = ops.vae_attn(qkv, C)
    assert torch.isfinite(out.float()).all()
    assert _rel(out, _ref(qkv)) < 6e-3


def test_fused_vae_attention_full_size_and_strided_rows():
    """720p latent frame count: 90 x 160 = 14 400 tokens per frame, 2 frames, qkv rows padded to a wider pitch."""
    from videocof_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(9)
    T, N = 2, 14400
    buf = torch.randn(T, N, 3 * C + 64, generator=g, device="cuda").bfloat16()
    qkv = buf[..., :3 * C]
    out = ops.vae_attn(qkv, C)
    ref = torch.cat([_ref(qkv[t:t + 1]) for t in range(T)])
    assert _rel(out, ref) < 6e-3
    # constant-V property over every row: the probabilities of a row sum to one across all 225 key blocks
    qkv2 = qkv.clone()
    qkv2[..., 2 * C:] = 0.75
    out2 = ops.vae_attn(qkv2.contiguous(), C)
    assert float((out2.float() - 0.75).abs().max()) < 2 * 2 ** -8


def test_attention_block_fused_equals_unfused(monkeypatch):
    """The VAE's AttentionBlock through the fused kernel and through the round-1 path (GEMM -> softmax -> GEMM)."""
    from videocof_b200 import vae as V
    torch.manual_seed(3)
    blk = V.AttentionBlock(C).to("cuda", torch.bfloat16).eval()
    with torch.no_grad():
        torch.nn.init.normal_(blk.proj.weight, std=0.05)       # the reference zero-initialises it (:241)
        x = torch.randn(2, 12, 20, C, device="cuda").bfloat16()
        monkeypatch.setenv("VCOF_VAE_ATTN", "unfused")
        a = V.attn_block(x, blk)
        monkeypatch.setenv("VCOF_VAE_ATTN", "fused")
        b = V.attn_block(x, blk)
    assert _rel(b, a) < 4e-3
