"""Host logic of videocof_b200.text_encoder on CPU: the module drives tests/vcof_emulator.py (an executable statement
of the C-ABI contracts) instead of libvcof, and must reproduce the golden outputs of the UNMODIFIED reference text
encoder (tools/gen_golden_t5.py) within the bf16 tolerance — state-dict keys, launch order, position-bias tables,
masks, batching."""
import os

import numpy as np
import pytest
import torch

import vcof_emulator
from gen_golden_t5 import T5_CASES, t5_inputs
from oracle.t5_oracle import T5Config, make_t5_params, position_bias, t5_forward
from videocof_b200.text_encoder import T5RelativeEmbedding, WanT5EncoderModel


def build(name):
    ckw, B, L, lens = T5_CASES[name]
    cfg = T5Config(**ckw)
    params = make_t5_params(cfg, seed=19)
    model = WanT5EncoderModel(**cfg.to_kwargs())
    model.load_state_dict(params, strict=True)                       # same keys as the reference module
    return cfg, params, model.to(torch.bfloat16).eval(), t5_inputs(cfg.vocab, B, L, lens)


@pytest.mark.parametrize("name", list(T5_CASES))
def test_host_module_matches_reference_golden(name, golden_dir, monkeypatch):
    vcof_emulator.install_t5(monkeypatch)
    cfg, params, model, (ids, mask) = build(name)
    gold = torch.from_numpy(np.load(os.path.join(golden_dir, name + ".npz"))["out"])
    out = model(ids, attention_mask=mask)[0]
    assert out.dtype == torch.bfloat16 and tuple(out.shape) == tuple(gold.shape)
    rel = float((out.float() - gold).norm() / gold.norm())
    assert rel < 2e-2, rel                                           # bf16 compute vs the fp32 reference
    emu = t5_forward(params, cfg, ids, mask, emulate_bf16=True)      # same rounding points: much tighter
    rel_emu = float((out.float() - emu).norm() / emu.norm())
    assert rel_emu < 6e-3, rel_emu


def test_state_dict_keys_are_the_reference_ones():
    cfg = T5Config(vocab=50, dim=32, dim_attn=32, dim_ffn=64, num_heads=2, num_layers=2, shared_pos=False)
    model = WanT5EncoderModel(**cfg.to_kwargs())
    assert set(model.state_dict()) == set(make_t5_params(cfg))
    shared = T5Config(vocab=50, dim=32, dim_attn=32, dim_ffn=64, num_heads=2, num_layers=1, shared_pos=True)
    assert set(WanT5EncoderModel(**shared.to_kwargs()).state_dict()) == set(make_t5_params(shared))


def test_bias_table_equals_the_dense_bias():
    emb = T5RelativeEmbedding(32, 4, bidirectional=True)
    L = 150
    tab = emb.table(L)                                               # [heads, 2L-1]
    dense = position_bias(emb.embedding.weight.detach(), L, L)       # [heads, L, L] as the reference builds it
    idx = (torch.arange(L)[None, :] - torch.arange(L)[:, None]) + L - 1
    assert torch.equal(tab[:, idx], dense)
    assert torch.equal(emb(L, L)[0], dense)
    with torch.no_grad():
        emb.embedding.weight.mul_(2.0)                               # in-place mutation invalidates the cached table
    assert torch.equal(emb.table(L)[:, idx], dense * 2)


def test_rejects_cpu_weights_bad_ids_and_3d_masks(monkeypatch):
    cfg, params, model, (ids, mask) = build("t5_tiny_shared")
    with pytest.raises(Exception, match="no CPU path"):
        model(ids)
    vcof_emulator.install_t5(monkeypatch)
    with pytest.raises(IndexError):
        model(torch.full_like(ids, cfg.vocab))
    with pytest.raises(NotImplementedError):
        model(ids, attention_mask=torch.ones(1, ids.shape[1], ids.shape[1]))


def test_from_pretrained_filters_kwargs_and_loads(tmp_path, monkeypatch):
    """fast_infer.py:321-326 passes the whole `text_encoder_kwargs` YAML block (config/wan2.1/wan_civitai.yaml:14-26:
    sub-paths, text_length, … next to the constructor arguments) and a single .pth / .safetensors state dict."""
    from safetensors.torch import save_file
    cfg, params, _, (ids, mask) = build("t5_tiny")
    kw = dict(cfg.to_kwargs(), text_encoder_subpath="x.pth", tokenizer_subpath="google/umt5-xxl", text_length=512)
    pth, st = str(tmp_path / "t5.pth"), str(tmp_path / "t5.safetensors")
    torch.save(params, pth)
    save_file({k: v.contiguous() for k, v in params.items()}, st)
    for path in (pth, st):
        m = WanT5EncoderModel.from_pretrained(path, additional_kwargs=kw, low_cpu_mem_usage=True,
                                              torch_dtype=torch.bfloat16)
        assert m.dtype == torch.bfloat16 and m.device.type == "cpu" and not m.training
        sd = m.state_dict()
        assert set(sd) == set(params)
        assert all(torch.equal(sd[k], params[k].to(torch.bfloat16)) for k in params)
    vcof_emulator.install_t5(monkeypatch)
    ref = t5_forward(params, cfg, ids, mask, emulate_bf16=True)
    out = m(ids, attention_mask=mask)[0]
    assert float((out.float() - ref).norm() / ref.norm()) < 6e-3
