"""On-device LoRA merge for the DiT (and the umT5 text encoder): `merge_lora` / `unmerge_lora` with the reference's signatures
(videox_fun/utils/lora_utils.py:371 and :503; called by fast_infer.py:371-385, :449).

The reference lifts every target Linear to fp32 on `device`, adds `multiplier * alpha/rank * up @ down` and casts
back to bf16, one layer at a time through ATen.  Here each layer is ONE tcgen05 GEMM whose epilogue reads the bf16
weight, adds the fp32-accumulated product scaled per column and writes the bf16 weight back in place
(`VCOF_EPI_GATE_ACCUM_BF16`, include/vcof.h) — no fp32 copy of the 14B weights ever exists.  The LoRA factors are
fp32 in the reference's arithmetic; they are fed to the bf16 tensor cores as a hi/lo split
(up = up_hi + up_lo, down = down_hi + down_lo; the product keeps the three leading terms in a 3x-rank GEMM), which
reproduces the fp32 product to ~2^-16 relative — far below the final bf16 rounding of the weight.

No packed copies of the Linear weights exist in videocof_b200 (the GEMM reads nn.Linear's own [N, K] storage), so
nothing has to be rebuilt after a merge; small cached vectors are keyed on `_version` elsewhere.
"""
from collections import defaultdict

import torch

from . import ops
from ._lib import VcofError

_PREFIX_DIT = "lora_unet"
_PREFIX_TE = "lora_te"


def _group_keys(state_dict):
    """{layer token: {"lora_down.weight", "lora_up.weight", "alpha"}} with the reference's two checkpoint dialects
    folded onto kohya tokens (lora_utils.py:379-394)."""
    groups = defaultdict(dict)
    for key, value in state_dict.items():
        if "diffusion_model" in key:
            key = key.replace("diffusion_model.", _PREFIX_DIT + "__")
            for a, b in (("blocks.", "blocks_"), (".self_attn.", "_self_attn_"), (".cross_attn.", "_cross_attn_"),
                         (".ffn.", "_ffn_")):
                key = key.replace(a, b)
        if "lora_A" in key or "lora_B" in key:
            key = _PREFIX_DIT + "__" + key
            for a, b in (("blocks.", "blocks_"), (".self_attn.", "_self_attn_"), (".cross_attn.", "_cross_attn_"),
                         (".ffn.", "_ffn_"), (".lora_A.default.", ".lora_down."), (".lora_B.default.", ".lora_up.")):
                key = key.replace(a, b)
        layer, elem = key.split(".", 1)
        groups[layer][elem] = value
    return groups


def _module_index(root):
    """underscore-joined module path -> module, for every sub-module: what the reference's attribute search
    (lora_utils.py:403-466) reaches for a `blocks_7_self_attn_q`-style token."""
    return {name.replace(".", "_"): mod for name, mod in root.named_modules() if name}


def _split_hi_lo(t):
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return hi, lo


def _apply(weight, up, down, scale, device):
    """weight (bf16 [N, K], any device) += scale * up[N, r] @ down[r, K] on `device` through libvcof."""
    if weight.dtype != torch.bfloat16:
        raise VcofError(f"merge_lora: target weight is {weight.dtype}; the B200 path holds DiT weights in bf16")
    dev = weight.device if weight.is_cuda else torch.device(device)
    if dev.type != "cuda":
        raise VcofError("merge_lora: no CUDA device to run on (weights on CPU and device='cpu'); "
                        "videocof_b200 has no CPU path")
    w = weight if weight.is_cuda else weight.to(dev)
    if not w.is_contiguous():
        raise VcofError("merge_lora: target weight must be contiguous")
    up = up.to(dev, torch.float32)
    down = down.to(dev, torch.float32)
    r = up.shape[1]
    rp = (3 * r + 7) // 8 * 8
    uh, ul = _split_hi_lo(up)
    dh, dl = _split_hi_lo(down)
    a = torch.zeros((up.shape[0], rp), dtype=torch.bfloat16, device=dev)        # [N, 3r]: up_hi | up_lo | up_hi
    b = torch.zeros((down.shape[1], rp), dtype=torch.bfloat16, device=dev)      # [K, 3r]: dn_hi | dn_hi | dn_lo
    a[:, :r], a[:, r:2 * r], a[:, 2 * r:3 * r] = uh, ul, uh
    b[:, :r], b[:, r:2 * r], b[:, 2 * r:3 * r] = dh.t(), dh.t(), dl.t()
    gate = torch.full((down.shape[1],), float(scale), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ops.gemm(a, b, None, "gate_accum", out=w, gate=gate)
    if w is not weight:
        weight.copy_(w)
    else:
        # the kernel wrote through the raw pointer: bump the tensor's version counter so that anything cached on
        # `weight._version` (the small fp32 vector caches of dit.py / vae.py, packed conv weights) is rebuilt
        weight.mul_(1)          # x * 1 keeps every bit (x + 0 would turn -0.0 into +0.0)


def _run(pipeline, state_dict, multiplier, device, dtype, transformer_only, sub_transformer_name, sign):
    if dtype != torch.float32:
        raise NotImplementedError("merge_lora: only the reference's default fp32 accumulation is implemented "
                                  f"(got dtype={dtype})")
    root = getattr(pipeline, sub_transformer_name)
    index = _module_index(root)
    te = getattr(pipeline, "text_encoder", None)
    te_index = None
    merged = 0
    with torch.no_grad():
        for layer, elems in _group_keys(state_dict).items():
            if _PREFIX_TE in layer:
                # `lora_te_…` tokens resolve under pipeline.text_encoder (lora_utils.py:406-411)
                if transformer_only or te is None:
                    continue
                te_index = _module_index(te) if te_index is None else te_index
                mod = te_index.get(layer.split(_PREFIX_TE + "_")[-1].lstrip("_"))
            else:
                mod = index.get(layer.split(_PREFIX_DIT + "_")[-1].lstrip("_"))
            if mod is None:
                print(f"Error loading layer: {layer}")            # the reference logs and moves on (:425-459)
                continue
            if not hasattr(mod, "weight"):
                continue
            if "lora_up.weight" not in elems or "lora_down.weight" not in elems:
                continue
            up, down = elems["lora_up.weight"], elems["lora_down.weight"]
            if up.dim() == 4:
                up, down = up.squeeze(3).squeeze(2), down.squeeze(3).squeeze(2)
            w = mod.weight.data
            if w.dim() != 2 or up.dim() != 2:
                raise NotImplementedError(f"merge_lora: {layer} targets a {w.dim()}-D weight; only Linear layers "
                                          "are LoRA targets on the VideoCoF path")
            alpha = float(elems["alpha"].item()) / up.shape[1] if "alpha" in elems else 1.0
            _apply(w, up, down, sign * multiplier * alpha, device)
            merged += 1
    return merged


def merge_lora(pipeline, lora_path, multiplier, device="cpu", dtype=torch.float32, state_dict=None,
               transformer_only=False, sub_transformer_name="transformer"):
    """reference lora_utils.py:371-500 (same signature; returns the pipeline)."""
    if state_dict is None:
        from safetensors.torch import load_file
        state_dict = load_file(lora_path)
    _run(pipeline, state_dict, multiplier, device, dtype, transformer_only, sub_transformer_name, 1.0)
    return pipeline


def unmerge_lora(pipeline, lora_path, multiplier=1, device="cpu", dtype=torch.float32,
                 sub_transformer_name="transformer"):
    """reference lora_utils.py:503-618."""
    from safetensors.torch import load_file
    _run(pipeline, load_file(lora_path), multiplier, device, dtype, False, sub_transformer_name, -1.0)
    return pipeline
