"""CUDA VAE (videocof_b200.vae on libvcof kernels) vs the goldens of the executed reference, and every libvcof
VAE entry point vs its executable C-ABI statement (tests/vcof_emulator.py) on identical inputs.

Tolerances: latents (mu) relative Frobenius error < 2e-2; decoded pixels PSNR >= 40 dB against the fp32
reference (signal range [-1, 1]); single ops relative error < 4e-3 (bf16 output rounding + fp32 accumulation
order)."""
import os

import numpy as np
import pytest
import torch

import vcof_emulator as emu
from gen_golden_vae_impl import VAE_CASES, vae_inputs
from oracle.vae_oracle import VAEConfig, make_vae_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    from videocof_b200.vae import AutoencoderKLWan
    m = AutoencoderKLWan()
    m.load_state_dict(make_vae_params(VAEConfig(), seed=17), strict=True)
    return m.to("cuda", torch.bfloat16).eval()


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def psnr(a, b):
    mse = float(((a.float().cpu() - b.float().cpu()) ** 2).mean())
    return 10 * np.log10(4.0 / max(mse, 1e-20))


# ---- single kernels against the C-ABI statement --------------------------------------------------------
CONV_CASES = {
    "c3x3x3_32_96": dict(cin=32, cout=96, k=(3, 3, 3), T=3, H=12, W=20),
    "c3x3x3_96_96_res": dict(cin=96, cout=96, k=(3, 3, 3), T=2, H=9, W=17, res=True),
    "c3x3x3_384_384": dict(cin=384, cout=384, k=(3, 3, 3), T=2, H=8, W=16),
    "c3x3x3_192_384": dict(cin=192, cout=384, k=(3, 3, 3), T=1, H=5, W=7),
    "c3x3x3_96_96_res_act": dict(cin=96, cout=96, k=(3, 3, 3), T=2, H=9, W=17, res=True, act=True),
    "c3x3x3_192_384_actonly": dict(cin=192, cout=384, k=(3, 3, 3), T=1, H=8, W=16, act=True, raw=False),
    "head_96_3_clamp": dict(cin=96, cout=3, k=(3, 3, 3), T=2, H=16, W=16, clamp=1.0, n_store=3),
    "time_768": dict(cin=384, cout=768, k=(3, 1, 1), T=3, H=4, W=6),
}


@pytest.mark.parametrize("kc", ["32", "64"])
@pytest.mark.parametrize("name", list(CONV_CASES))
def test_conv_igemm_vs_contract(name, kc, monkeypatch):
    """Both K-slice widths of the tap-streaming kernel (64- and 128-byte TMA rows, SW64 / SW128 descriptors)."""
    monkeypatch.setenv("VCOF_CONV_KC", kc)
    monkeypatch.setenv("VCOF_CONV_LINES", "0")
    _conv_case(name, monkeypatch)


def _conv_case(name, monkeypatch):
    from videocof_b200 import vae
    c = CONV_CASES[name]
    torch.manual_seed(0)
    conv = vae.CausalConv3d(c["cin"], c["cout"], c["k"], padding=tuple(k // 2 for k in c["k"]))
    conv.weight.data = (conv.weight.data * 3).bfloat16().float()
    conv.bias.data = conv.bias.data.bfloat16().float()
    x = torch.randn(c["T"], c["H"], c["W"], c["cin"]).bfloat16()
    res = torch.randn(c["T"], c["H"], c["W"], (c["cout"] + 7) // 8 * 8).bfloat16() if c.get("res") else None
    kw = dict(clamp=c.get("clamp", 0.0), n_store=c.get("n_store"))
    norm = None
    if c.get("act"):
        norm = vae.RMS_norm(c["cout"], images=False)
        norm.gamma.data = (torch.rand_like(norm.gamma) + 0.5).bfloat16().float()
        kw.update(act_norm=norm, want_raw=c.get("raw", True))
    from videocof_b200 import ops
    real = ops.conv_igemm, ops.conv_lines, ops.rms_silu_cl
    # contract (CPU)
    ops.conv_igemm, ops.conv_lines, ops.rms_silu_cl = emu.conv_igemm, emu.conv_lines, emu.rms_silu_cl
    try:
        want = vae.conv_causal(x, conv, residual=res, **kw)
    finally:
        ops.conv_igemm, ops.conv_lines, ops.rms_silu_cl = real
    conv_cuda = conv.to("cuda")
    conv_cuda.__dict__.pop("_vcof_pack", None)
    if norm is not None:
        norm.__dict__.pop("_vcof_pack", None)
        kw["act_norm"] = norm.to("cuda")
    got = vae.conv_causal(x.cuda(), conv_cuda, residual=None if res is None else res.cuda(), **kw)
    torch.cuda.synchronize()
    if not isinstance(want, tuple):
        want, got = (want,), (got,)
    for g_, w_ in zip(got, want):
        assert (g_ is None) == (w_ is None)
        if g_ is not None:
            assert rel(g_, w_) < 4e-3, rel(g_, w_)


LINE_CASES = ["c3x3x3_96_96_res", "c3x3x3_384_384", "c3x3x3_192_384", "c3x3x3_96_96_res_act", "c3x3x3_32_96",
              "head_96_3_clamp", "c3x3x3_192_384_actonly"]


@pytest.mark.parametrize("name", LINE_CASES)
def test_conv_lines_vs_contract(name, monkeypatch):
    """vcof_conv_lines (VCOF_CONV_LINES=1) against the same CPU contract as the default kernel."""
    monkeypatch.setenv("VCOF_CONV_LINES", "1")
    _conv_case(name, monkeypatch)


def test_resamplers_and_norm_vs_contract():
    from videocof_b200 import ops, vae
    torch.manual_seed(1)
    names = ("gemm", "conv_igemm", "conv_lines", "rms_silu_cl", "softmax_rows")
    real = {n: getattr(ops, n) for n in names}
    for mode, cin in (("downsample2d", 96), ("downsample3d", 192), ("upsample3d", 384), ("upsample2d", 192)):
        rs = vae.Resample(cin, mode)
        for p in rs.parameters():
            p.data = (p.data * 2).bfloat16().float()
        x = torch.randn(5, 8, 12, cin).bfloat16()
        fn = vae.downsample if mode.startswith("down") else vae.upsample
        for n in names:
            setattr(ops, n, getattr(emu, n))
        try:
            want = fn(x, rs)
        finally:
            for n in names:
                setattr(ops, n, real[n])
        rs_c = rs.to("cuda")
        for m in rs_c.modules():
            m.__dict__.pop("_vcof_pack", None)
        got = fn(x.cuda(), rs_c)
        torch.cuda.synchronize()
        assert got.shape == want.shape, (mode, got.shape, want.shape)
        assert rel(got, want) < 4e-3, (mode, rel(got, want))
    g = torch.rand(192) + 0.5
    x = torch.randn(3, 5, 7, 192).bfloat16()
    for silu in (True, False):
        want = emu.rms_silu_cl(x, g, silu)
        got = ops.rms_silu_cl(x.cuda(), g.cuda(), silu)
        assert rel(got, want) < 3e-3, rel(got, want)
    s = torch.randn(37, 533) * 3
    assert rel(ops.softmax_rows(s.cuda(), 0.7), emu.softmax_rows(s, 0.7)) < 3e-3


# ---- whole encode / decode vs the reference goldens -------------------------------------------------------
@pytest.mark.parametrize("name", list(VAE_CASES))
def test_vae_encode_decode_vs_reference_golden(name, golden_dir, model):
    T, H, W = VAE_CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    video, z = vae_inputs(T, H, W)
    with torch.no_grad():
        post = model.encode(video[None].cuda().bfloat16())[0]
        dec = model.decode(z[None].cuda().bfloat16()).sample[0]
    mu = post.mode()[0]
    ref_mu, ref_dec = torch.from_numpy(gold["mu"]), torch.from_numpy(gold["dec"])
    assert tuple(mu.shape) == tuple(ref_mu.shape) and tuple(dec.shape) == tuple(ref_dec.shape)
    assert rel(mu, ref_mu) < 2e-2, rel(mu, ref_mu)
    assert psnr(dec, ref_dec) >= 40.0, psnr(dec, ref_dec)
    assert float(dec.abs().max()) <= 1.0


def test_vae_rejects_cpu_tensors(model):
    from videocof_b200._lib import VcofError
    with pytest.raises(VcofError):
        model.decode(torch.zeros(1, 16, 1, 4, 4, dtype=torch.bfloat16))
