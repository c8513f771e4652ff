"""umT5 text encoder on the GPU: the new libvcof entry points against plain torch fp32 math,
and videocof_b200.text_encoder.WanT5EncoderModel against the golden outputs of the UNMODIFIED reference
(tests/golden/t5_*.npz) and against the CPU oracle with bf16 rounding emulated.

Tolerances (bf16 compute): single kernels 4e-3 relative Frobenius (one or two bf16 roundings of O(1) values); the
attention 6e-3 (bf16 probabilities); the whole encoder 3e-2 against the fp32 reference (the contract emulator on CPU
sits at 0.9e-2 .. 1.4e-2 on these cases: ~20 bf16 roundings per layer at widths of 64 .. 256) and 1.2e-2 against the
oracle that rounds where the kernels do (two bf16 pipelines that differ only in fp32 summation order sit at
4e-3 .. 6e-3 at the real width: a flipped rounding moves a value by a whole bf16 ulp)."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_fro(got, ref):
    return float((got.float() - ref.float()).norm() / (ref.float().norm() + 1e-30))


@pytest.mark.parametrize("M,N,K", [(192, 256, 256), (512, 4096, 4096), (333, 10240, 512), (200, 104, 64)])
@pytest.mark.parametrize("epi", ["mul", "add"])
def test_gemm_bf16_rmw_epilogues(M, N, K, epi):
    from videocof_b200 import ops
    torch.manual_seed(7)
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    o0 = torch.randn(M, N, device="cuda").bfloat16()
    got = o0.clone()
    ops.gemm(a, w, None, epi, out=got)
    lin = (a.float() @ w.float().t()).bfloat16().float()
    ref = o0.float() * lin if epi == "mul" else o0.float() + lin
    assert rel_fro(got, ref) < 4e-3


@pytest.mark.parametrize("rows,C", [(1, 64), (77, 4096), (1024, 256)])
def test_t5_rmsnorm(rows, C):
    from videocof_b200 import ops
    torch.manual_seed(3)
    x = (torch.randn(rows, C, device="cuda") * 3).bfloat16()
    w = (1 + 0.1 * torch.randn(C, device="cuda")).bfloat16()
    got = ops.t5_rmsnorm(x, w, 1e-6)
    xf = x.float()
    ref = w.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16().float()
    assert rel_fro(got, ref) < 4e-3


def test_embed_rows():
    from videocof_b200 import ops
    torch.manual_seed(5)
    table = torch.randn(1000, 256, device="cuda").bfloat16()
    ids = torch.randint(0, 1000, (777,), device="cuda")
    assert torch.equal(ops.embed_rows(ids, table), table[ids])


@pytest.mark.parametrize("B,L,heads,d,lens", [(1, 512, 8, 64, None), (2, 512, 4, 64, [512, 77]), (2, 96, 4, 64, [96, 1]),
                                              (1, 1, 2, 64, None), (3, 160, 4, 16, [160, 13, 100]),
                                              (1, 300, 2, 128, None), (1, 200, 3, 32, [150])])
def test_t5_attention(B, L, heads, d, lens):
    from videocof_b200 import ops
    torch.manual_seed(11)
    C = heads * d
    q = (torch.randn(B * L, C, device="cuda") * 0.7).bfloat16()
    k = torch.randn(B * L, C, device="cuda").bfloat16()
    v = torch.randn(B * L, C, device="cuda").bfloat16()
    bias_rel = torch.randn(heads, 2 * L - 1, device="cuda")
    mask = None
    if lens is not None:
        mask = torch.zeros(B, L, dtype=torch.int32, device="cuda")
        for b, n in enumerate(lens):
            mask[b, :n] = 1
    got = ops.t5_attention(q, k, v, bias_rel, B, L, heads, key_mask=mask)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().view(B, L, heads, d) for t in (q, k, v))
    idx = (torch.arange(L, device="cuda")[None, :] - torch.arange(L, device="cuda")[:, None]) + L - 1
    bias = bias_rel[:, idx].unsqueeze(0).expand(B, -1, -1, -1).clone()
    if mask is not None:
        bias.masked_fill_((mask == 0).view(B, 1, 1, L), torch.finfo(torch.bfloat16).min)
    p = torch.softmax(torch.einsum("binc,bjnc->bnij", qf, kf) + bias, dim=-1)
    ref = torch.einsum("bnij,bjnc->binc", p, vf).reshape(B * L, C)
    assert not torch.isnan(got).any()
    assert rel_fro(got, ref) < 6e-3


def _model(cfg, params):
    from videocof_b200.text_encoder import WanT5EncoderModel
    m = WanT5EncoderModel(**cfg.to_kwargs())
    m.load_state_dict(params, strict=True)
    return m.to("cuda", torch.bfloat16).eval()


@pytest.mark.parametrize("name", ["t5_tiny", "t5_tiny_shared", "t5_d64"])
def test_encoder_matches_reference_golden(name, golden_dir):
    from gen_golden_t5 import T5_CASES, t5_inputs
    from oracle.t5_oracle import T5Config, make_t5_params, t5_forward
    ckw, B, L, lens = T5_CASES[name]
    cfg = T5Config(**ckw)
    params = make_t5_params(cfg, seed=19)
    ids, mask = t5_inputs(cfg.vocab, B, L, lens)
    gold = torch.from_numpy(np.load(os.path.join(golden_dir, name + ".npz"))["out"])
    out = _model(cfg, params)(ids.cuda(), attention_mask=None if mask is None else mask.cuda())[0]
    torch.cuda.synchronize()
    assert out.dtype == torch.bfloat16 and tuple(out.shape) == tuple(gold.shape)
    assert rel_fro(out.cpu(), gold) < 3e-2
    emu = t5_forward(params, cfg, ids, mask, emulate_bf16=True)
    assert rel_fro(out.cpu(), emu) < 1.2e-2


def test_encoder_umt5_width_two_layers():
    """The real layer shape (dim 4096, 64 heads x 64, ffn 10240, 512 tokens, a 77-token prompt and an empty one)
    on two layers with a small vocabulary, against the CPU oracle."""
    from oracle.t5_oracle import T5Config, make_t5_params, t5_forward
    cfg = T5Config(vocab=512, num_layers=2)
    params = make_t5_params(cfg, seed=23)
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(1, cfg.vocab, (2, 512), generator=g)
    mask = torch.zeros(2, 512, dtype=torch.long)
    mask[0, :77] = 1
    mask[1, :1] = 1
    ids[mask == 0] = 0
    out = _model(cfg, params)(ids.cuda(), attention_mask=mask.cuda())[0]
    torch.cuda.synchronize()
    ref = t5_forward(params, cfg, ids, mask, emulate_bf16=True)
    assert rel_fro(out[0, :77].cpu(), ref[0, :77]) < 1.2e-2
    assert rel_fro(out.cpu(), ref) < 1.2e-2
    assert rel_fro(out.cpu(), t5_forward(params, cfg, ids, mask)) < 3e-2


def test_encoder_rejects_cpu_and_fp32():
    from oracle.t5_oracle import T5Config, make_t5_params
    from videocof_b200._lib import VcofError
    from videocof_b200.text_encoder import WanT5EncoderModel
    cfg = T5Config(vocab=50, dim=64, dim_attn=64, dim_ffn=128, num_heads=4, num_layers=1)
    m = WanT5EncoderModel(**cfg.to_kwargs())
    m.load_state_dict(make_t5_params(cfg, seed=1))
    ids = torch.ones(1, 8, dtype=torch.long)
    with pytest.raises(VcofError):
        m.eval()(ids)                                           # CPU weights
    with pytest.raises(VcofError):
        m.to("cuda").eval()(ids.cuda())                         # fp32 weights
