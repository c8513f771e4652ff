"""Golden for the LoRA merge: runs the UNMODIFIED reference merge_lora / unmerge_lora
(/root/reference/videox_fun/utils/lora_utils.py:371-618, loaded through tools/ref_loader.py) on the tiny DiT with
bf16 weights and a synthetic LoRA checkpoint (oracle.lora_oracle.make_lora_state), and stores the bit patterns of
every weight after the merge and after the unmerge.

    python tools/gen_golden_lora.py
"""
import os
import sys
import tempfile
import types
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
warnings.filterwarnings("ignore")

import ref_loader  # noqa: E402
from gen_golden import DIT_CASES, checksum  # noqa: E402

MULT = 0.8
RANK = 8


def linear_shapes(params):
    return {k[:-len(".weight")]: tuple(v.shape) for k, v in params.items()
            if k.endswith(".weight") and v.dim() == 2 and ".norm" not in k and k.startswith("blocks.")}


def bits(t):
    return t.detach().contiguous().view(torch.int16).numpy().copy()


def main():
    from oracle.dit_oracle import DiTConfig, make_dit_params
    from oracle.lora_oracle import make_lora_state
    from safetensors.torch import save_file
    ns = ref_loader.load_reference()
    ref_loader._mod("diffusers.models.lora", LoRACompatibleConv=type("LoRACompatibleConv", (), {}),
                    LoRACompatibleLinear=type("LoRACompatibleLinear", (), {}))
    lora = ref_loader._load("videox_fun.utils.lora_utils", "videox_fun/utils/lora_utils.py")
    ckw, _, _, _ = DIT_CASES["dit_tiny"]
    cfg = DiTConfig(**ckw)
    params = make_dit_params(cfg, seed=11)
    model = ns.dit.WanTransformer3DModel(**cfg.to_kwargs()).eval()
    model.load_state_dict(params, strict=True)
    model = model.to(torch.bfloat16)
    sd = make_lora_state(linear_shapes(params), rank=RANK, seed=5)
    pipe = types.SimpleNamespace(transformer=model)
    lora.merge_lora(pipe, None, MULT, device="cpu", dtype=torch.float32, state_dict=dict(sd), transformer_only=True)
    merged = {k: bits(v) for k, v in model.state_dict().items() if v.dtype == torch.bfloat16}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "lora.safetensors")
        # the text-encoder entries would need pipeline.text_encoder in unmerge_lora (no transformer_only switch there)
        save_file({k: v.contiguous() for k, v in sd.items() if not k.startswith("lora_te")}, path)
        lora.unmerge_lora(pipe, path, MULT, device="cpu", dtype=torch.float32)
    unmerged = {k: bits(v) for k, v in model.state_dict().items() if v.dtype == torch.bfloat16}
    changed = sorted(k for k in merged if not np.array_equal(merged[k], bits(params[k].to(torch.bfloat16))))
    import zlib
    out = {"param_checksum": np.float64(checksum(params)), "changed": np.array(changed),
           "crc_merged": np.array([zlib.crc32(merged[k].tobytes()) for k in changed], dtype=np.uint32),
           "crc_unmerged": np.array([zlib.crc32(unmerged[k].tobytes()) for k in changed], dtype=np.uint32)}
    # full bit patterns for one layer of each kind (fixture size); every changed weight is pinned by its CRC
    for k in ("blocks.0.self_attn.q.weight", "blocks.1.cross_attn.k.weight", "blocks.0.ffn.0.weight",
              "blocks.1.ffn.2.weight"):
        out["merged/" + k] = merged[k]
        out["unmerged/" + k] = unmerged[k]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lora_tiny.npz"), **out)
    print("wrote lora_tiny.npz:", len(changed), "weights changed, e.g.", changed[:3])


if __name__ == "__main__":
    main()
