"""The hot kernels at BASELINE.json's full sizes (configs[1], "c2": Wan-14B, 81 f x 720p latents = 75 600 tokens, 40 heads,
C = 5120, F = 13 824), next to the same checks at the CPU-runnable config's token count ("c1": 1 280 tokens, 1.3B widths).

The oracle finishes a whole DiT forward only at c1 (tests/test_dit_gpu.py).  At c2 the kernels are checked one by one, on
the exact launch shapes of a c2 denoising step:
  * token-local kernels (the three block GEMMs with their epilogues, LN+modulate, RMSNorm+RoPE): EVERY output element
    against fp32 torch math on the GPU (tools/gpu_probe.run_case — the same statements as the small cases);
  * self-attention (one 100 ms launch, 75 600 x 75 600 x 40 heads): sampled query rows — tile edges, the ragged last
    block, random rows — for three heads against the oracle's attention_ref (oracle/dit_oracle.py), plus a
    size-independent property over every row of every head: with V constant along the keys the softmax weights must sum
    to one, so each output row reproduces that constant to bf16 rounding.
Tolerances are those of the small cases (tools/gpu_probe.py), which the kernels meet with a 5x margin."""
import pytest
import torch

import gpu_probe
from oracle import dit_oracle

pytestmark = pytest.mark.gpu

CONFIGS = {"c1": dict(F=5, H=16, W=16, heads=12, ffn=8960), "c2": dict(F=21, H=45, W=80, heads=40, ffn=13824)}


def _dims(cfg):
    c = CONFIGS[cfg]
    return c["F"] * c["H"] * c["W"], c["heads"] * 128, c["ffn"], c


@pytest.mark.parametrize("cfg", ["c1", "c2"])
@pytest.mark.parametrize("which", ["qkvo", "ffn_up", "ffn_down", "proj_residual"])
def test_block_gemms_every_element(cfg, which):
    L, C, F, _ = _dims(cfg)
    M, N, K, epi = {"qkvo": (L, C, C, "bias"), "ffn_up": (L, F, C, "bias_gelu"),
                    "ffn_down": (L, C, F, "bias_gate_res"), "proj_residual": (L, C, C, "bias_gate_res")}[which]
    res = gpu_probe.run_case(dict(kind="gemm", M=M, N=N, K=K, epi=epi))
    assert res["ok"], res


@pytest.mark.parametrize("cfg", ["c1", "c2"])
@pytest.mark.parametrize("mode", ["mod", "affine"])
def test_ln_modulate_every_element(cfg, mode):
    L, C, _, _ = _dims(cfg)
    res = gpu_probe.run_case(dict(kind="ln", L=L, C=C, mode=mode))
    assert res["ok"], res


@pytest.mark.parametrize("cfg", ["c1", "c2"])
def test_rmsnorm_rope_every_element(cfg):
    _, _, _, c = _dims(cfg)
    res = gpu_probe.run_case(dict(kind="rms", F=c["F"], H=c["H"], W=c["W"], heads=c["heads"], mode="cot"))
    assert res["ok"], res


def _sample_rows(L, g):
    edges = [0, 1, 127, 128, 255, 256, L // 2 - 1, L // 2, L - 257, L - 129, L - 128, L - 2, L - 1]
    rnd = torch.randint(0, L, (48,), generator=g).tolist()
    return torch.tensor(sorted({r for r in edges + rnd if 0 <= r < L}), dtype=torch.long)


@pytest.mark.parametrize("cfg", ["c1", "c2"])
def test_self_attention_sampled_rows_and_normalisation(cfg):
    from videocof_b200 import ops
    L, C, _, c = _dims(cfg)
    n, d = c["heads"], 128
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(L)
    gd = torch.Generator(device="cuda").manual_seed(L)
    q = torch.randn(L, C, device=dev, generator=gd).bfloat16()
    k = torch.randn(L, C, device=dev, generator=gd).bfloat16()
    v = torch.randn(L, C, device=dev, generator=gd).bfloat16()
    out = ops.attention(q, k, v, n)
    assert bool(torch.isfinite(out.float()).all())
    rows = _sample_rows(L, g).to(dev)
    heads = sorted({0, n // 2, n - 1})
    qs = q[rows].view(-1, n, d)[:, heads]
    ref = dit_oracle.attention_ref(qs, k.view(L, n, d)[:, heads], v.view(L, n, d)[:, heads])     # fp32 [rows, 3, d]
    got = out[rows].view(-1, n, d)[:, heads].float()
    rel = float((got - ref).norm() / ref.norm())
    assert rel < 1.5e-2, rel                                   # tools/gpu_probe.py's attention tolerance
    # property over all L x heads rows: V constant along the keys -> every output row is that constant
    const = torch.randn(C, device=dev, generator=gd).bfloat16()
    out_c = ops.attention(q, k, const[None, :].expand(L, C).contiguous(), n).float()
    err = (out_c - const.float()[None, :]).abs()
    bound = const.float().abs()[None, :] * 2.0 ** -6 + 1e-6    # two bf16 ulps of the constant
    assert bool((err <= bound).all()), float((err / (const.float().abs()[None, :] + 1e-6)).max())


@pytest.mark.parametrize("cfg", ["c1", "c2"])
def test_cross_attention_every_element(cfg):
    """Cross-attention launch shape: L queries against the 512-row text context (kv_len = 512), all heads."""
    L, _, _, c = _dims(cfg)
    if cfg == "c2":
        # scores of all 40 heads at once would be 6 GB of fp32: check the launch in three head groups instead
        from videocof_b200 import ops
        n, d = c["heads"], 128
        dev = torch.device("cuda")
        gd = torch.Generator(device="cuda").manual_seed(7)
        q = torch.randn(L, n * d, device=dev, generator=gd).bfloat16()
        k = torch.randn(512, n * d, device=dev, generator=gd).bfloat16()
        v = torch.randn(512, n * d, device=dev, generator=gd).bfloat16()
        out = ops.attention(q, k, v, n).view(L, n, d)
        for h0 in range(0, n, 8):
            hs = slice(h0, h0 + 8)
            ref = dit_oracle.attention_ref(q.view(L, n, d)[:, hs], k.view(512, n, d)[:, hs], v.view(512, n, d)[:, hs])
            got = out[:, hs].float()
            rel = float((got - ref).norm() / ref.norm())
            assert rel < 1.5e-2, (h0, rel)
        return
    res = gpu_probe.run_case(dict(kind="attn", Lq=L, Lk=512, kv=512, heads=c["heads"], vt=0))
    assert res["ok"], res


# ---- VAE convolutions at 720p -------------------------------------------------------------------------------------
# The three layer shapes that carry the decoder's time at full resolution (68 % of the decoder in the
# 96- and 192-channel 3x3x3 convolutions), on 5 frames of the 81 f x 720p clip — every output element against
# torch.nn.functional.conv3d in fp32 on the GPU, rounded where the C-ABI contract rounds (tests/vcof_emulator.py:
# bf16(acc + bias), + residual in fp32, clamp, bf16).  "small" is the same code on a 24 x 40 frame (dry-run on the CPU).
VAE_LAYERS = {
    "96_96_res": dict(cin=96, cout=96, res=True, div=1),         # full-resolution ResidualBlock conv (line-resident kernel)
    "192_192": dict(cin=192, cout=192, res=False, div=2),        # half-resolution stage (tap-streaming kernel)
    "head_96_3": dict(cin=96, cout=3, res=False, div=1, clamp=1.0, n_store=3),   # decoder head, clamp fused
}


@pytest.mark.parametrize("size", ["small", "720p"])
@pytest.mark.parametrize("layer", list(VAE_LAYERS))
@torch.no_grad()
def test_vae_conv_every_element(size, layer):
    import torch.nn.functional as Fn
    from videocof_b200 import vae
    c = VAE_LAYERS[layer]
    H, W = ((24, 40) if size == "small" else (720, 1280))
    T, H, W = 5, H // c["div"], W // c["div"]
    dev = torch.device("cuda")
    torch.manual_seed(3)
    conv = vae.CausalConv3d(c["cin"], c["cout"], (3, 3, 3), padding=(1, 1, 1))
    conv.weight.data = (conv.weight.data * 3).bfloat16().float()
    conv.bias.data = conv.bias.data.bfloat16().float()
    conv = conv.to(dev)
    gd = torch.Generator(device="cuda").manual_seed(T * H)
    x = torch.randn(T, H, W, c["cin"], device=dev, generator=gd).bfloat16()
    ldc = (c["cout"] + 7) // 8 * 8
    res = torch.randn(T, H, W, ldc, device=dev, generator=gd).bfloat16() if c["res"] else None
    got = vae.conv_causal(x, conv, residual=res, clamp=c.get("clamp", 0.0), n_store=c.get("n_store"))
    torch.cuda.synchronize()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        xin = Fn.pad(x.permute(3, 0, 1, 2)[None].float(), (1, 1, 1, 1, 2, 0))      # causal in time, 'same' in space
        ref = Fn.conv3d(xin, conv.weight.float(), conv.bias.float())[0].permute(1, 2, 3, 0)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    ref = ref.bfloat16().float()
    if res is not None:
        ref = ref + res[..., :c["cout"]].float()
    if c.get("clamp"):
        ref = ref.clamp(-c["clamp"], c["clamp"])
    ref = ref.bfloat16().float()
    g = got[..., :c["cout"]].float()
    assert tuple(g.shape) == tuple(ref.shape)
    rel = float((g - ref).norm() / ref.norm())
    assert rel < 4e-3, rel                                     # tests/test_vae_gpu.py's single-op tolerance
