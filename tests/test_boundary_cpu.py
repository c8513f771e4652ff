"""Drop-in boundary checks that need no GPU: state-dict ABI, C-ABI exports, loud failure modes."""
import ctypes
import json
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_keys_match_reference(golden_dir):
    """Keys/shapes recorded from the executed reference constructor (tools/gen_golden.py)."""
    from videocof_b200.dit import WanTransformer3DModel
    spec = json.load(open(os.path.join(golden_dir, "dit_state_keys.json")))
    m = WanTransformer3DModel(**spec["config"])
    ours = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert ours == spec["state_dict"]
    # attributes the pipeline / CLIs touch
    assert m.config.in_channels == 16 and tuple(m.config.patch_size) == (1, 2, 2)
    assert m.freqs.dtype == torch.complex128 and tuple(m.freqs.shape) == (1024, 64)
    assert isinstance(m.blocks, torch.nn.ModuleList)
    for name in ("enable_teacache", "disable_teacache", "share_teacache", "enable_cfg_skip", "disable_cfg_skip",
                 "share_cfg_skip", "enable_riflex", "disable_riflex", "enable_multi_gpus_inference"):
        assert callable(getattr(m, name))


def test_lora_style_weight_mutation_targets_exist(golden_dir):
    """merge_lora walks getattr chains like blocks_0_self_attn_q and mutates .weight.data in place
    (reference utils/lora_utils.py:412-416, 490-495)."""
    from videocof_b200.dit import WanTransformer3DModel
    spec = json.load(open(os.path.join(golden_dir, "dit_state_keys.json")))
    m = WanTransformer3DModel(**spec["config"])
    for path in ("blocks.0.self_attn.q", "blocks.1.cross_attn.v", "blocks.0.ffn.0", "blocks.1.ffn.2"):
        layer = m.get_submodule(path)
        v0 = layer.weight._version
        layer.weight.data += 0.5
        assert layer.weight._version >= v0


def header_functions():
    src = open(os.path.join(ROOT, "include", "vcof.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vcof_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from videocof_b200 import _lib
    names = header_functions()
    assert "vcof_gemm_bf16" in names and "vcof_attn_fwd" in names
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vcof.h but not exported by libvcof.so"
    bound = set(_lib.SIGNATURES) | {"vcof_last_error", "vcof_abi_version"}
    assert set(names) == bound, "include/vcof.h and videocof_b200/_lib.py disagree"
    assert _lib.load().vcof_abi_version() == 1


def test_ops_fail_loudly_without_cuda_tensors():
    from videocof_b200 import ops
    from videocof_b200._lib import VcofError
    a = torch.zeros(128, 64, dtype=torch.bfloat16)
    with pytest.raises(VcofError):
        ops.gemm(a, a)
    with pytest.raises(VcofError):
        ops.ln_modulate(torch.zeros(4, 64))
    with pytest.raises(VcofError):
        ops.attention(a, a, a, 1)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under videocof_b200/ or videox_fun/ may import it."""
    bad = []
    for pkg in ("videocof_b200", "videox_fun"):
        for dp, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith(".py"):
                    s = open(os.path.join(dp, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_c_abi_rejects_bad_arguments_with_a_message():
    """Error behaviour of the C ABI (include/vcof.h): a malformed call returns a negative code BEFORE anything is
    launched and leaves a message naming the entry point in vcof_last_error(); the Python shim turns that into
    VcofError (the reference's style is an assert / exception at the call site, e.g. attention_utils.py:72-73).
    Argument validation touches no device, so this runs without a GPU."""
    from videocof_b200 import _lib
    lib = _lib.load()
    buf = (ctypes.c_char * 4096)()
    p = ctypes.addressof(buf)
    p = (p + 255) // 256 * 256                      # an aligned host address: never dereferenced by validation
    ptrs = (ctypes.c_void_p * 17)(*([p] * 17))      # host array of (fake) device addresses
    odd = (ctypes.c_void_p * 2)(p, p + 8)
    bad = [
        ("vcof_attn_fwd", (p, 64, p, 64, p, 64, p, 64, 128, 128, 128, 1, 64, 0.125, 0, None), "head_dim"),
        ("vcof_attn_fwd", (p, 128, p, 128, p, 128, p, 128, 128, 128, 200, 1, 128, 0.09, 0, None), "kv_len"),
        ("vcof_attn_fwd", (p, 128, p, 128, p, 128, p, 128, 0, 128, 128, 1, 128, 0.09, 0, None), "empty"),
        ("vcof_gemm_bf16", (p, 64, p, 64, None, None, p, 64, 0, 64, 64, 0, None), "empty"),
        ("vcof_gemm_bf16", (p, 60, p, 64, None, None, p, 64, 128, 64, 64, 0, None), "vcof_gemm_bf16"),
        ("vcof_gemm_bf16", (p, 64, p, 64, None, None, p, 300, 128, 300, 64, 0, None), "ldo"),
        ("vcof_gemm_bf16", (p, 64, p, 64, p, None, p, 64, 128, 64, 64, 5, None), "vcof_gemm_bf16"),
        ("vcof_ln_modulate", (p, 64, None, None, None, None, p, 64, 0, 64, 1e-6, None), "empty"),
        ("vcof_ln_modulate", (p, 66, None, None, None, None, p, 66, 4, 66, 1e-6, None), "vcof_ln_modulate"),
        ("vcof_rmsnorm_rope", (p, 256, p, 1e-6, 4, 256, 100, None, None, 1, 1, 1, 0, 0, 0, None), "head_dim"),
        ("vcof_copy_blocked", (p, 64, p, 64, 4, 60, 8, 1, None), "vcof_copy_blocked"),
        ("vcof_patchify", (p, p, 16, 2, 5, 8, None), "vcof_patchify"),
        ("vcof_unpatchify", (p, 64, p, 16, 2, 8, 7, None), "vcof_unpatchify"),
        ("vcof_linear_f32", (p, p, None, p, 1, 8, 12, 0, 0, None), "vcof_linear_f32"),
        ("vcof_rmsnorm_rope_scatter", (p, 256, None, 2, p, 1e-6, 4, 256, 128, None, None, 1, 1, 1, 0, 0, 0, None),
         "destination pointers"),
        ("vcof_rmsnorm_rope_scatter", (p, 256, ptrs, 17, p, 1e-6, 4, 256, 128, None, None, 1, 1, 1, 0, 0, 0, None),
         "destination pointers"),
        ("vcof_rmsnorm_rope_scatter", (p, 256, ptrs, 3, p, 1e-6, 4, 256, 128, None, None, 1, 1, 1, 0, 0, 0, None),
         "blocks"),
        ("vcof_attn_fwd_scatter", (p, 128, p, 128, p, 128, None, 2, 64, 128, 128, 128, 128, 1, 128, 0.09, None),
         "output chunks"),
        ("vcof_attn_fwd_scatter", (p, 128, p, 128, p, 128, ptrs, 2, 32, 128, 128, 128, 128, 1, 128, 0.09, None),
         "do not cover"),
        ("vcof_attn_fwd_scatter", (p, 128, p, 128, p, 128, ptrs, 2, 64, 128, 128, 128, 128, 1, 64, 0.09, None),
         "head_dim"),
        ("vcof_copy_scatter", (p, 60, ptrs, 2, 4, 64, None), "vcof_copy_scatter"),
        ("vcof_copy_scatter", (p, 64, odd, 2, 4, 64, None), "aligned"),
        ("vcof_copy_rows_scatter", (p, 64, ptrs, 2, 0, 64, None), "vcof_copy_rows_scatter"),
        ("vcof_cl_to_u8", (p, 8, p, 0, 3, None), "vcof_cl_to_u8"),
        ("vcof_cl_to_u8", (p, 2, p, 16, 3, None), "vcof_cl_to_u8"),
        ("vcof_cl_to_u8", (None, 8, p, 16, 3, None), "null"),
        ("vcof_u8_to_cl", (p, p, 16, 3, 2, None), "vcof_u8_to_cl"),
        ("vcof_u8_to_cl", (p, None, 16, 3, 32, None), "null"),
        ("vcof_embed_rows", (p, p, 64, 100, p, 64, 0, 64, None), "empty"),
        ("vcof_embed_rows", (p, p, 64, 100, p, 64, 5, 60, None), "multiples of 8"),
        ("vcof_t5_rmsnorm", (p, 64, p, p, 64, 4, 60, 1e-6, None), "vcof_t5_rmsnorm"),
        ("vcof_t5_attn", (p, 64, p, 64, p, 64, p, 64, p, 1023, None, 1, 513, 1, 64, None), "512"),
        ("vcof_t5_attn", (p, 64, p, 64, p, 64, p, 64, None, 0, None, 1, 16, 1, 64, None), "bias"),
    ]
    for name, args, needle in bad:
        rc = getattr(lib, name)(*args)
        msg = lib.vcof_last_error().decode()
        assert rc < 0, (name, args, rc)
        base = "vcof_attn_fwd" if name == "vcof_attn_fwd_scatter" else name      # shared implementation, shared prefix
        assert needle in msg and base in msg, (name, needle, msg)
        with pytest.raises(_lib.VcofError, match=name):
            _lib.call(name, *args)
