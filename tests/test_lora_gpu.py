"""LoRA merge through libvcof (videocof_b200.lora -> vcof_gemm_bf16 with the accumulate epilogue) against
oracle/lora_oracle.py (itself pinned bit-exactly to the reference, tests/test_lora_oracle.py).

Tolerance (written here, floating point): the merged weight is a bf16 rounding of fp32(W) + s * up @ down; the kernel
accumulates the hi/lo-split bf16 factors in fp32 in a different order than torch.mm, so a value that lands within
~2^-16 of a rounding boundary may round the other way: at most one bf16 ulp *at the scale of the operands*
(|got - want| <= 2^-7 * max(|W_before|, |W_after|) — where W and the update cancel, the result is tiny and its own
ulp is not the right yardstick — plus the split's own floor 2^-16 * |s| * ||up_row|| * ||down_col||: the dropped
lo*lo terms scale with the factors, not with their possibly cancelling sum), on at most 0.2 % of the elements."""
import types

import pytest
import torch

from oracle import lora_oracle
from oracle.dit_oracle import DiTConfig, make_dit_params

pytestmark = pytest.mark.gpu
CFG = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=32)


def _floors(sd, names, multiplier):
    """name -> 2^-16 * |s| * outer(||up rows||, ||down cols||) for every merged layer."""
    out = {}
    for layer, e in lora_oracle.normalise_keys(sd).items():
        name = lora_oracle.resolve(layer, names)
        if name is None or "lora_up.weight" not in e or "lora_down.weight" not in e:
            continue
        up, down = e["lora_up.weight"].float(), e["lora_down.weight"].float()
        s = multiplier * (float(e["alpha"]) / up.shape[1] if "alpha" in e else 1.0)
        out[name] = 2.0 ** -16 * abs(s) * torch.outer(up.norm(dim=1), down.norm(dim=0))
    return out


def _check(got, want, before, name, floor):
    got, want, before = got.detach().cpu().float(), want.float(), before.float()
    err = (got - want).abs()
    bound = 2.0 ** -7 * torch.maximum(before.abs(), want.abs()) + floor
    assert bool((err <= bound).all()), (name, float((err - bound).max()))
    frac = float((err > 0).float().mean())
    assert frac < 2e-3, (name, frac)


@pytest.mark.parametrize("rank", [8, 4, 20])
def test_merge_unmerge_vs_oracle(rank, tmp_path):
    from safetensors.torch import save_file
    from videocof_b200 import lora
    from videocof_b200.dit import WanTransformer3DModel
    cfg = DiTConfig(**CFG)
    params = make_dit_params(cfg, seed=11)
    model = WanTransformer3DModel(**cfg.to_kwargs())
    model.load_state_dict(params, strict=True)
    model = model.to(torch.bfloat16).cuda()
    shapes = {k[:-7]: tuple(v.shape) for k, v in params.items()
              if k.endswith(".weight") and v.dim() == 2 and ".norm" not in k and k.startswith("blocks.")}
    sd = lora_oracle.make_lora_state(shapes, rank=rank, seed=5)
    weights = {k[:-7]: v.to(torch.bfloat16) for k, v in params.items() if k.endswith(".weight")}
    touched = lora_oracle.merge(weights, sd, 0.8)
    floors = _floors(sd, list(weights), 0.8)
    pipe = types.SimpleNamespace(transformer=model)
    lora.merge_lora(pipe, None, 0.8, device="cuda", state_dict=dict(sd), transformer_only=True)
    torch.cuda.synchronize()
    state = model.state_dict()
    for name in touched:
        _check(state[name + ".weight"], weights[name], params[name + ".weight"], name, floors[name])
    for k, v in state.items():                                  # everything else untouched, bit for bit
        if k.endswith(".weight") and k[:-7] not in touched and v.dtype == torch.bfloat16:
            assert torch.equal(v.cpu(), params[k].to(torch.bfloat16)), k
    path = str(tmp_path / "lora.safetensors")
    sd_file = {k: v.contiguous() for k, v in sd.items() if not k.startswith("lora_te")}
    save_file(sd_file, path)
    # unmerge starts from the kernel's merged weights: compare with the oracle applied to exactly those
    merged = {k[:-7]: v.detach().cpu().clone() for k, v in state.items() if k.endswith(".weight")}
    before = {k: v.clone() for k, v in merged.items()}
    lora_oracle.merge(merged, sd_file, 0.8, sign=-1.0)
    lora.unmerge_lora(pipe, path, 0.8, device="cuda")
    torch.cuda.synchronize()
    state = model.state_dict()
    for name in touched:
        _check(state[name + ".weight"], merged[name], before[name], name + " (unmerge)", floors[name])


def test_merge_14b_sized_layer_and_cpu_resident_weight():
    """One 5120 x 5120 projection at rank 64 (the 14B shape), with the nn.Linear left on the CPU: the reference moves
    each layer to `device` for the update and back (lora_utils.py:471-497) — so does this path, through the GPU."""
    from videocof_b200 import lora
    g = torch.Generator().manual_seed(3)
    lin = torch.nn.Linear(5120, 5120, bias=False).to(torch.bfloat16)
    lin.weight.data = (torch.randn(5120, 5120, generator=g) * 0.02).to(torch.bfloat16)
    w0 = lin.weight.data.clone()
    sd = {"diffusion_model.proj.lora_down.weight": torch.randn(64, 5120, generator=g) / 70,
          "diffusion_model.proj.lora_up.weight": torch.randn(5120, 64, generator=g) * 0.05,
          "diffusion_model.proj.alpha": torch.tensor(32.0)}
    root = torch.nn.Module()
    root.proj = lin
    pipe = types.SimpleNamespace(transformer=root)
    lora.merge_lora(pipe, None, 1.0, device="cuda", state_dict=sd, transformer_only=True)
    torch.cuda.synchronize()
    assert not lin.weight.is_cuda
    want = {"proj": w0}
    assert lora_oracle.merge(want, sd, 1.0) == ["proj"]
    _check(lin.weight.data, want["proj"], w0, "proj", _floors(sd, ["proj"], 1.0)["proj"])
    assert float((lin.weight.data.float() - w0.float()).abs().mean()) > 1e-4      # the update is not a no-op
