"""Whole-forward parity of the CUDA DiT (videocof_b200.dit, every op a libvcof kernel) against the
CPU oracle and against golden outputs of the executed reference.

Tolerances (stated per north_star: floating-point path, bf16 compute vs fp32 reference):
  * vs the oracle with bf16 rounding emulated at the reference's CUDA rounding points:
      relative Frobenius error < 1.5e-2
  * vs the fp32 gold / the reference goldens: relative Frobenius error < 4e-2
"""
import os

import numpy as np
import pytest
import torch

from gen_golden import DIT_CASES, ROPE_MODES, dit_inputs
from oracle.dit_oracle import DiTConfig, dit_forward, make_dit_params

pytestmark = pytest.mark.gpu


def build_cuda_model(cfg, params):
    from videocof_b200.dit import WanTransformer3DModel
    m = WanTransformer3DModel(**cfg.to_kwargs())
    m.load_state_dict(params, strict=True)
    return m.to("cuda", torch.bfloat16).eval()


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("name", ["dit_tiny", "dit_tiny_b2"])
@pytest.mark.parametrize("mode", list(ROPE_MODES))
def test_dit_forward_vs_oracle_and_golden(name, mode, golden_dir):
    ckw, shape, n_ctx, B = DIT_CASES[name]
    cfg = DiTConfig(**ckw)
    params = make_dit_params(cfg, seed=11)
    x, ctx, t = dit_inputs(shape, n_ctx, cfg.text_dim, B, seed=23)
    f = shape[1]
    seq_len = f * (shape[2] // 2) * (shape[3] // 2)
    kw = ROPE_MODES[mode](f, B)
    model = build_cuda_model(cfg, params)
    with torch.no_grad():
        y = model(x=x.cuda().bfloat16(), t=t.cuda(), context=[c.cuda().bfloat16() for c in ctx],
                  seq_len=seq_len, **kw)
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == (B,) + shape
    xr = x.bfloat16().float()
    cr = [c.bfloat16().float() for c in ctx]
    emu = dit_forward(params, cfg, xr, t, cr, seq_len, emulate_bf16=True, **kw)
    assert rel(y, emu) < 1.5e-2, ("vs bf16-emulating oracle", rel(y, emu))
    gold = torch.from_numpy(np.load(os.path.join(golden_dir, name + ".npz"))["out_" + mode])
    assert rel(y, gold) < 4e-2, ("vs reference golden", rel(y, gold))


def test_dit_c1_shape_two_layers_vs_golden(golden_dir):
    """1.3B widths (C=1536, F=8960, 12 heads) at the C1 token count (L=1280), chain-of-frames RoPE."""
    name = "dit_c1_2layer"
    ckw, shape, n_ctx, B = DIT_CASES[name]
    cfg = DiTConfig(**ckw)
    params = make_dit_params(cfg, seed=11)
    x, ctx, t = dit_inputs(shape, n_ctx, cfg.text_dim, B, seed=23)
    f = shape[1]
    seq_len = f * (shape[2] // 2) * (shape[3] // 2)
    kw = ROPE_MODES["cot"](f, B)
    model = build_cuda_model(cfg, params)
    with torch.no_grad():
        y = model(x=x.cuda().bfloat16(), t=t.cuda(), context=[c.cuda().bfloat16() for c in ctx],
                  seq_len=seq_len, **kw)
    gold = torch.from_numpy(np.load(os.path.join(golden_dir, name + ".npz"))["out_cot"])
    assert rel(y, gold) < 4e-2, rel(y, gold)


def test_padded_sequence_matches_unpadded():
    """seq_len > L (the SP padding rule, reference :904-910): padded rows must not leak into real tokens."""
    ckw, shape, n_ctx, B = DIT_CASES["dit_tiny"]
    cfg = DiTConfig(**ckw)
    params = make_dit_params(cfg, seed=11)
    x, ctx, t = dit_inputs(shape, n_ctx, cfg.text_dim, B, seed=23)
    L = shape[1] * (shape[2] // 2) * (shape[3] // 2)
    model = build_cuda_model(cfg, params)
    args = dict(x=x.cuda().bfloat16(), t=t.cuda(), context=[c.cuda().bfloat16() for c in ctx])
    with torch.no_grad():
        a = model(seq_len=L, **args)
        b = model(seq_len=L + 37, **args)
    assert torch.equal(a, b)


def test_forward_rejects_cpu_and_fp32():
    from videocof_b200._lib import VcofError
    from videocof_b200.dit import WanTransformer3DModel
    ckw, shape, n_ctx, B = DIT_CASES["dit_tiny"]
    cfg = DiTConfig(**ckw)
    m = WanTransformer3DModel(**cfg.to_kwargs())
    x, ctx, t = dit_inputs(shape, n_ctx, cfg.text_dim, B, seed=23)
    with pytest.raises(VcofError):
        m(x=x, t=t, context=ctx, seq_len=240)           # CPU tensors: no fallback
    m = m.to("cuda")                                      # fp32 weights: refused, not silently cast
    with pytest.raises(VcofError):
        m(x=x.cuda().bfloat16(), t=t.cuda(), context=[c.cuda().bfloat16() for c in ctx], seq_len=240)


def test_teacache_skips_the_same_steps_as_the_reference(golden_dir):
    """TeaCache gate (reference wan_transformer3d.py:956-1031): same skip decisions and outputs over 4 steps."""
    from gen_golden import TEACACHE, TEACACHE_T
    ckw, shape, n_ctx, B = DIT_CASES["dit_tiny"]
    cfg = DiTConfig(**ckw)
    params = make_dit_params(cfg, seed=11)
    gold = np.load(os.path.join(golden_dir, "dit_tiny_teacache.npz"))
    x, ctx, _ = dit_inputs(shape, n_ctx, cfg.text_dim, B, seed=23)
    f = shape[1]
    seq_len = f * (shape[2] // 2) * (shape[3] // 2)
    model = build_cuda_model(cfg, params)
    model.enable_teacache(**TEACACHE)
    calc = []
    for i, tv in enumerate(TEACACHE_T):
        with torch.no_grad():
            y = model(x=x.cuda().bfloat16(), t=torch.tensor([tv]).cuda(), context=[c.cuda().bfloat16() for c in ctx],
                      seq_len=seq_len, **ROPE_MODES["cot"](f, B))
        calc.append(bool(model.should_calc))
        assert rel(y, torch.from_numpy(gold["outs"][i])) < 4e-2, (i, rel(y, torch.from_numpy(gold["outs"][i])))
    assert calc == [bool(v) for v in gold["should_calc"]]
    assert not all(calc), "fixture must contain at least one skipped step"


def test_cfg_skip_drops_the_unconditional_half():
    """cfg_skip (reference utils/cfg_optimization.py:5-38): late steps run the cond half only and duplicate it."""
    ckw, shape, n_ctx, _ = DIT_CASES["dit_tiny"]
    cfg = DiTConfig(**ckw)
    params = make_dit_params(cfg, seed=11)
    x, ctx, t = dit_inputs(shape, n_ctx, cfg.text_dim, 2, seed=23)
    f = shape[1]
    seq_len = f * (shape[2] // 2) * (shape[3] // 2)
    model = build_cuda_model(cfg, params)
    args = dict(x=x.cuda().bfloat16(), t=t.cuda(), context=[c.cuda().bfloat16() for c in ctx], seq_len=seq_len)
    with torch.no_grad():
        full = model(**args)
        model.enable_cfg_skip(0.5, 4)
        model.current_steps = 3
        skipped = model(**args)
    assert torch.equal(skipped[0], skipped[1])
    assert torch.equal(skipped[1], full[1])


def test_batched_cfg_forward_is_bit_identical_to_the_per_sample_loop(monkeypatch):
    """Batch 2 (classifier-free guidance): tokens of both samples stacked along M (weights stream once per step) must
    reproduce the per-sample loop bit for bit — the GEMM tiles of a row do not depend on how many rows follow it."""
    ckw, shape, n_ctx, B = DIT_CASES["dit_tiny_b2"]
    cfg = DiTConfig(**ckw)
    params = make_dit_params(cfg, seed=11)
    x, ctx, t = dit_inputs(shape, n_ctx, cfg.text_dim, B, seed=23)
    f = shape[1]
    seq_len = f * (shape[2] // 2) * (shape[3] // 2)
    model = build_cuda_model(cfg, params)
    args = dict(x=x.cuda().bfloat16(), t=t.cuda(), context=[c.cuda().bfloat16() for c in ctx], seq_len=seq_len + 3,
                **ROPE_MODES["cot"](f, B))
    with torch.no_grad():
        monkeypatch.setenv("VCOF_DIT_BATCHED", "0")
        loop = model(**args)
        monkeypatch.setenv("VCOF_DIT_BATCHED", "1")
        batched = model(**args)
    assert torch.equal(loop, batched)


def test_context_cache_forward_is_bit_identical(monkeypatch):
    """SURVEY §8a a4 / a11 (K3 / K10): the text embedding and the cross-attention K / V of it are step-invariant;
    with `enable_context_cache` they are computed at the first forward only, and two consecutive cached forwards
    (different latents and timesteps, same prompt embeddings) equal the uncached ones bit for bit — per-sample loop
    and batched CFG path — with fewer launches on the second."""
    from videocof_b200 import ops
    ckw, shape, n_ctx, B = DIT_CASES["dit_tiny_b2"]
    cfg = DiTConfig(**ckw)
    params = make_dit_params(cfg, seed=11)
    x, ctx, t = dit_inputs(shape, n_ctx, cfg.text_dim, B, seed=23)
    f = shape[1]
    seq_len = f * (shape[2] // 2) * (shape[3] // 2)
    model = build_cuda_model(cfg, params)
    ctx = [c.cuda().bfloat16() for c in ctx]
    kw = dict(context=ctx, seq_len=seq_len, **ROPE_MODES["cot"](f, B))
    steps = [(x.cuda().bfloat16(), t.cuda()), ((x * 0.5).cuda().bfloat16(), (t * 0.5).cuda())]
    for batched in ("1", "0"):
        monkeypatch.setenv("VCOF_DIT_BATCHED", batched)
        with torch.no_grad():
            model.disable_context_cache()
            ref, n_ref = [], []
            for xx, tt in steps:
                ops.reset_launches()
                ref.append(model(x=xx, t=tt, **kw))
                n_ref.append(ops.launches())
            model.enable_context_cache()
            got, n_got = [], []
            for xx, tt in steps:
                ops.reset_launches()
                got.append(model(x=xx, t=tt, **kw))
                n_got.append(ops.launches())
            model.disable_context_cache()
        for a, b in zip(got, ref):
            assert torch.equal(a, b)
        layers = len(model.blocks)
        assert n_ref[0] == n_ref[1]
        # second cached forward: no text embedding (2 GEMMs per sample) and no K / norm / V launches in any block
        saved = B * 2 + 3 * layers * (1 if batched == "1" else B)
        assert n_got[1] == n_ref[1] - saved, (n_got, n_ref, saved)
