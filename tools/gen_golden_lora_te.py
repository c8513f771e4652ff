"""Golden for LoRA entries that target the text encoder: runs the UNMODIFIED reference merge_lora / unmerge_lora
(videox_fun/utils/lora_utils.py:371-618) on a pipeline whose text_encoder is the reference's own WanT5EncoderModel
(tiny, bf16) with a synthetic `lora_te_…` checkpoint (oracle.lora_oracle.make_te_lora_state); stores the CRC of every
weight after the merge and after the unmerge.

    python tools/gen_golden_lora_te.py
"""
import os
import sys
import tempfile
import types
import warnings
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
warnings.filterwarnings("ignore")

import ref_loader  # noqa: E402
from gen_golden_lora import bits  # noqa: E402
from gen_golden_pipeline import DIT_KW, T5_KW  # noqa: E402

MULT, RANK = 0.6, 4


def te_linear_shapes(params):
    return {k[:-len(".weight")]: tuple(v.shape) for k, v in params.items()
            if k.endswith(".weight") and v.dim() == 2 and "embedding" not in k}


def main():
    from oracle.dit_oracle import DiTConfig, make_dit_params
    from oracle.lora_oracle import make_te_lora_state
    from oracle.t5_oracle import T5Config, make_t5_params
    from safetensors.torch import save_file
    ns = ref_loader.load_reference_pipeline()
    ref_loader._mod("diffusers.models.lora", LoRACompatibleConv=type("LoRACompatibleConv", (), {}),
                    LoRACompatibleLinear=type("LoRACompatibleLinear", (), {}))
    lora = ref_loader._load("videox_fun.utils.lora_utils", "videox_fun/utils/lora_utils.py")
    dcfg, tcfg = DiTConfig(**DIT_KW), T5Config(**T5_KW)
    dit = ns.dit.WanTransformer3DModel(**dcfg.to_kwargs()).eval()
    dit.load_state_dict(make_dit_params(dcfg, seed=11), strict=True)
    params = make_t5_params(tcfg, seed=19)
    t5 = ns.t5.WanT5EncoderModel(**tcfg.to_kwargs()).eval()
    t5.load_state_dict(params, strict=True)
    dit, t5 = dit.to(torch.bfloat16), t5.to(torch.bfloat16)
    sd = make_te_lora_state(te_linear_shapes(params), rank=RANK, seed=9)
    pipe = types.SimpleNamespace(transformer=dit, text_encoder=t5)
    lora.merge_lora(pipe, None, MULT, device="cpu", dtype=torch.float32, state_dict=dict(sd))
    merged = {k: bits(v) for k, v in t5.state_dict().items()}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "lora.safetensors")
        save_file({k: v.contiguous() for k, v in sd.items()}, path)
        lora.unmerge_lora(pipe, path, MULT, device="cpu", dtype=torch.float32)
    unmerged = {k: bits(v) for k, v in t5.state_dict().items()}
    changed = sorted(k for k in merged if not np.array_equal(merged[k], bits(params[k].to(torch.bfloat16))))
    out = {"changed": np.array(changed),
           "crc_merged": np.array([zlib.crc32(merged[k].tobytes()) for k in changed], dtype=np.uint32),
           "crc_unmerged": np.array([zlib.crc32(unmerged[k].tobytes()) for k in changed], dtype=np.uint32)}
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lora_te_tiny.npz"), **out)
    print("wrote lora_te_tiny.npz:", len(changed), "weights changed:", changed)


if __name__ == "__main__":
    main()
