"""CPU restatement of the reference's LoRA merge for the DiT (test infrastructure only — the product never imports
this; see tests/test_boundary_cpu.py).

Follows /root/reference/videox_fun/utils/lora_utils.py:
  * key normalisation                      :379-394  (kohya `diffusion_model.` keys and PEFT `lora_A/lora_B` keys)
  * layer resolution                       :403-466  (the `blocks_7_self_attn_q` front/back search; for every name the
                                                      search can resolve it lands on the module with that dotted path)
  * skip rules                             :468-481  (no `.weight`; missing up or down)
  * the update itself                      :482-496  W <- dtype(W) + multiplier * (alpha / rank) * up @ down, computed
                                                      in `dtype` (fp32 by default, fast_infer.py:371-385), then cast back
  * unmerge                                :503-618  same with a minus sign.

Pinned against the executed reference by tests/golden/lora_tiny.npz (tools/gen_golden_lora.py).
"""
import math
from collections import defaultdict

import torch


def normalise_keys(state_dict):
    """:376-394 -> {layer_token: {"lora_down.weight" | "lora_up.weight" | "alpha": tensor}}."""
    updates = defaultdict(dict)
    for key, value in state_dict.items():
        if "diffusion_model" in key:
            key = key.replace("diffusion_model.", "lora_unet__")
            key = key.replace("blocks.", "blocks_")
            key = key.replace(".self_attn.", "_self_attn_")
            key = key.replace(".cross_attn.", "_cross_attn_")
            key = key.replace(".ffn.", "_ffn_")
        if "lora_A" in key or "lora_B" in key:
            key = "lora_unet__" + key
            key = key.replace("blocks.", "blocks_")
            key = key.replace(".self_attn.", "_self_attn_")
            key = key.replace(".cross_attn.", "_cross_attn_")
            key = key.replace(".ffn.", "_ffn_")
            key = key.replace(".lora_A.default.", ".lora_down.")
            key = key.replace(".lora_B.default.", ".lora_up.")
        layer, elem = key.split(".", 1)
        updates[layer][elem] = value
    return updates


def resolve(layer_token, weight_names):
    """Dotted parameter prefix the reference's attribute search (:403-466) reaches for `layer_token`, or None.

    The search walks `_`-separated pieces greedily from the left, joining pieces until an attribute exists; on the
    DiT every module name is a single piece except `self_attn`, `cross_attn`, `norm_q`, `norm_k`, `text_embedding`,
    `time_embedding`, `time_projection`, `patch_embedding` — so the result is the unique dotted path whose
    underscore-joined form equals the token."""
    if "lora_te" in layer_token:
        return None
    return _by_flat_name(layer_token.split("lora_unet_")[-1].lstrip("_"), weight_names)


def resolve_text_encoder(layer_token, weight_names):
    """The same search rooted at pipeline.text_encoder for `lora_te_…` tokens (:406-411): on the umT5 encoder
    (wan_text_encoder.py) `lora_te_blocks_3_ffn_gate_0` lands on blocks.3.ffn.gate.0."""
    if "lora_te" not in layer_token:
        return None
    return _by_flat_name(layer_token.split("lora_te_")[-1].lstrip("_"), weight_names)


def _by_flat_name(flat, weight_names):
    for name in weight_names:
        if name.replace(".", "_") == flat:
            return name
    return None


def merge(weights, state_dict, multiplier, dtype=torch.float32, sign=1.0, text_encoder=False):
    """weights: {module path: weight tensor (e.g. bf16)} for every module that owns a `.weight`.  Returns the names it
    updated; tensors are replaced by new ones of the original dtype (:482-497).  With text_encoder=True `weights` are
    the text encoder's and only `lora_te_…` entries apply."""
    touched = []
    for layer, elems in normalise_keys(state_dict).items():
        name = (resolve_text_encoder if text_encoder else resolve)(layer, list(weights))
        if name is None:
            continue
        if "lora_up.weight" not in elems or "lora_down.weight" not in elems:
            continue
        w = weights[name]
        up = elems["lora_up.weight"].to(dtype)
        down = elems["lora_down.weight"].to(dtype)
        alpha = float(elems["alpha"].item()) / up.shape[1] if "alpha" in elems else 1.0
        if up.dim() == 4:
            delta = torch.mm(up.squeeze(3).squeeze(2), down.squeeze(3).squeeze(2)).unsqueeze(2).unsqueeze(3)
        else:
            delta = torch.mm(up, down)
        weights[name] = (w.to(dtype) + sign * (multiplier * alpha * delta)).to(w.dtype)
        touched.append(name)
    return touched


def make_lora_state(weight_shapes, rank=8, seed=0, dtype=torch.float32):
    """Deterministic synthetic LoRA checkpoint over a DiT's Linear layers, exercising every branch of the key
    normalisation: kohya-style keys with alpha for self-attention, PEFT-style keys (alpha absent -> 1.0) for
    cross-attention and FFN, one norm-only entry (skipped: no up/down pair), one unresolvable entry and one
    text-encoder entry (skipped for the DiT)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def rnd(*shape, std):
        return (torch.randn(*shape, generator=g) * std).to(dtype)

    for name, (n_out, k_in) in weight_shapes.items():
        if ".self_attn." in name:
            base = "diffusion_model." + name
            sd[base + ".lora_down.weight"] = rnd(rank, k_in, std=1.0 / math.sqrt(k_in))
            sd[base + ".lora_up.weight"] = rnd(n_out, rank, std=0.05)
            sd[base + ".alpha"] = torch.tensor(float(rank) / 2)
        elif ".cross_attn." in name or ".ffn." in name:
            sd[name + ".lora_A.default.weight"] = rnd(rank, k_in, std=1.0 / math.sqrt(k_in))
            sd[name + ".lora_B.default.weight"] = rnd(n_out, rank, std=0.05)
    first = next(iter(weight_shapes))
    blk = first.split(".self_attn")[0].split(".cross_attn")[0].split(".ffn")[0]
    sd["diffusion_model." + blk + ".norm3.lora_down.weight"] = rnd(rank, 4, std=1.0)          # no lora_up: skipped
    sd["diffusion_model.blocks.99.self_attn.q.lora_down.weight"] = rnd(rank, 4, std=1.0)       # unresolvable
    sd["diffusion_model.blocks.99.self_attn.q.lora_up.weight"] = rnd(4, rank, std=1.0)
    sd["lora_te_encoder_block_0_layer_0_SelfAttention_q.lora_down.weight"] = rnd(rank, 4, std=1.0)
    sd["lora_te_encoder_block_0_layer_0_SelfAttention_q.lora_up.weight"] = rnd(4, rank, std=1.0)
    return sd


def make_te_lora_state(weight_shapes, rank=4, seed=0, dtype=torch.float32):
    """kohya-style `lora_te_<underscored module path>` entries (with alpha) for every Linear of a umT5 encoder."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, (n_out, k_in) in weight_shapes.items():
        base = "lora_te_" + name.replace(".", "_")
        sd[base + ".lora_down.weight"] = (torch.randn(rank, k_in, generator=g) / math.sqrt(k_in)).to(dtype)
        sd[base + ".lora_up.weight"] = (torch.randn(n_out, rank, generator=g) * 0.05).to(dtype)
        sd[base + ".alpha"] = torch.tensor(float(rank) / 2)
    return sd
