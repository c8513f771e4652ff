"""Host logic of videocof_b200.lora (key dialects, module resolution, alpha / multiplier / sign, skip rules) against
the golden produced by the executed reference (tests/golden/lora_tiny.npz).  The per-layer update is replaced by the
fp32 statement of the C-ABI contract (no GPU here); the tcgen05 path itself is tests/test_lora_gpu.py."""
import os
import types
import zlib

import numpy as np
import pytest
import torch

from oracle.dit_oracle import DiTConfig, make_dit_params
from oracle.lora_oracle import make_lora_state

GOLD = os.path.join(os.path.dirname(__file__), "golden", "lora_tiny.npz")
CFG = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=32)
MULT, RANK = 0.8, 8


def _contract_apply(weight, up, down, scale, device):
    weight.copy_((weight.float() + scale * (up.float() @ down.float())).to(torch.bfloat16))


def _bits(t):
    return t.detach().contiguous().view(torch.int16).numpy()


def _model_and_state():
    from videocof_b200.dit import WanTransformer3DModel
    cfg = DiTConfig(**CFG)
    params = make_dit_params(cfg, seed=11)
    model = WanTransformer3DModel(**cfg.to_kwargs())
    model.load_state_dict(params, strict=True)
    model = model.to(torch.bfloat16)
    shapes = {k[:-7]: tuple(v.shape) for k, v in params.items()
              if k.endswith(".weight") and v.dim() == 2 and ".norm" not in k and k.startswith("blocks.")}
    return model, params, make_lora_state(shapes, rank=RANK, seed=5)


def test_merge_unmerge_host_logic_matches_reference_bits(monkeypatch, tmp_path, capsys):
    from safetensors.torch import save_file
    from videocof_b200 import lora
    monkeypatch.setattr(lora, "_apply", _contract_apply)
    model, params, sd = _model_and_state()
    pipe = types.SimpleNamespace(transformer=model)
    assert lora.merge_lora(pipe, None, MULT, device="cpu", state_dict=dict(sd), transformer_only=True) is pipe
    assert "blocks_99" in capsys.readouterr().out            # unresolvable layer is reported and skipped
    g = np.load(GOLD)
    state = model.state_dict()
    changed = [str(k) for k in g["changed"]]
    for k, crc in zip(changed, g["crc_merged"]):
        assert zlib.crc32(_bits(state[k]).tobytes()) == int(crc), k
    untouched = [k for k, v in state.items() if k not in changed and v.dtype == torch.bfloat16]
    for k in untouched:
        assert torch.equal(state[k], params[k].to(torch.bfloat16)), k
    path = str(tmp_path / "lora.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items() if not k.startswith("lora_te")}, path)
    lora.unmerge_lora(pipe, path, MULT, device="cpu")
    state = model.state_dict()
    for k, crc in zip(changed, g["crc_unmerged"]):
        assert zlib.crc32(_bits(state[k]).tobytes()) == int(crc), k


def test_merge_fails_loudly_without_gpu_and_rejects_other_dtypes():
    from videocof_b200 import lora
    from videocof_b200._lib import VcofError
    model, _, sd = _model_and_state()
    pipe = types.SimpleNamespace(transformer=model)
    if not torch.cuda.is_available():
        with pytest.raises(VcofError):
            lora.merge_lora(pipe, None, 1.0, device="cpu", state_dict=dict(sd), transformer_only=True)
    with pytest.raises(NotImplementedError):
        lora.merge_lora(pipe, None, 1.0, dtype=torch.bfloat16, state_dict=dict(sd), transformer_only=True)


def test_overlay_exports_merge_entry_points():
    import videox_fun.utils.lora_utils as m
    from videocof_b200 import lora
    assert m.merge_lora is lora.merge_lora and m.unmerge_lora is lora.unmerge_lora


def test_text_encoder_entries_merge_into_our_encoder(monkeypatch, tmp_path):
    """`lora_te_…` entries land on videocof_b200.text_encoder.WanT5EncoderModel's Linears exactly as the reference
    lands them on its own encoder (bit-exact CRCs from the executed reference, tools/gen_golden_lora_te.py);
    transformer_only=True skips them."""
    from safetensors.torch import save_file
    from gen_golden_lora_te import MULT as M_TE, RANK as R_TE, te_linear_shapes
    from gen_golden_pipeline import T5_KW
    from oracle.lora_oracle import make_te_lora_state
    from oracle.t5_oracle import T5Config, make_t5_params
    from videocof_b200 import lora
    from videocof_b200.text_encoder import WanT5EncoderModel
    monkeypatch.setattr(lora, "_apply", _contract_apply)
    tcfg = T5Config(**T5_KW)
    params = make_t5_params(tcfg, seed=19)
    t5 = WanT5EncoderModel(**tcfg.to_kwargs())
    t5.load_state_dict(params, strict=True)
    t5 = t5.to(torch.bfloat16)
    dit, _, _ = _model_and_state()
    sd = make_te_lora_state(te_linear_shapes(params), rank=R_TE, seed=9)
    pipe = types.SimpleNamespace(transformer=dit, text_encoder=t5)
    before = {k: v.clone() for k, v in t5.state_dict().items()}
    lora.merge_lora(pipe, None, M_TE, device="cpu", state_dict=dict(sd), transformer_only=True)
    assert all(torch.equal(v, before[k]) for k, v in t5.state_dict().items())
    lora.merge_lora(pipe, None, M_TE, device="cpu", state_dict=dict(sd))
    g = np.load(os.path.join(os.path.dirname(GOLD), "lora_te_tiny.npz"))
    changed = [str(k) for k in g["changed"]]
    state = t5.state_dict()
    for k, crc in zip(changed, g["crc_merged"]):
        assert zlib.crc32(_bits(state[k]).tobytes()) == int(crc), k
    assert all(torch.equal(v, before[k]) for k, v in state.items() if k not in changed)
    path = str(tmp_path / "te.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    lora.unmerge_lora(pipe, path, M_TE, device="cpu")
    state = t5.state_dict()
    for k, crc in zip(changed, g["crc_unmerged"]):
        assert zlib.crc32(_bits(state[k]).tobytes()) == int(crc), k
