"""Host API of WanTransformer3DModel that the CLIs touch before the first forward (fast_infer.py:281-351):
`from_pretrained` (config.json + sharded safetensors / .bin, `dict_mapping`, subfolder — reference
wan_transformer3d.py:1157-1299), the RIFLEx / TeaCache / cfg-skip switches (:731-800).  No libvcof call is involved:
everything here runs on the CPU.  RIFLEx tables are pinned to goldens of the executed reference
(tools/gen_golden_riflex.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.dit_oracle import DiTConfig, make_dit_params
from videocof_b200.dit import WanTransformer3DModel

CFG = DiTConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=32)


def _checkpoint(tmp_path, shards=2, fmt="safetensors", drop=None, sub="transformer"):
    from safetensors.torch import save_file
    d = tmp_path / sub if sub else tmp_path
    d.mkdir(parents=True, exist_ok=True)
    params = {k: v.to(torch.bfloat16).contiguous() for k, v in make_dit_params(CFG, seed=4).items()}
    saved = {k: v for k, v in params.items() if k != drop}
    cfg = dict(CFG.to_kwargs(), _class_name="WanTransformer3DModel", _diffusers_version="0.31.0", some_future_key=1)
    (d / "config.json").write_text(json.dumps(cfg))
    if fmt == "bin":
        torch.save(saved, d / "diffusion_pytorch_model.bin")
    else:
        keys = sorted(saved)
        for i in range(shards):
            save_file({k: saved[k] for k in keys[i::shards]},
                      str(d / f"diffusion_pytorch_model-{i + 1:05d}-of-{shards:05d}.safetensors"))
    return params


@pytest.mark.parametrize("fmt,shards", [("safetensors", 1), ("safetensors", 3), ("bin", 1)])
def test_from_pretrained_roundtrip(tmp_path, fmt, shards, capsys):
    params = _checkpoint(tmp_path, shards=shards, fmt=fmt)
    kw = {"transformer_subpath": "transformer", "dict_mapping": {"in_dim": "in_channels", "dim": "hidden_size"}}
    m = WanTransformer3DModel.from_pretrained(str(tmp_path), subfolder="transformer", transformer_additional_kwargs=kw,
                                              low_cpu_mem_usage=True, torch_dtype=torch.bfloat16)
    assert "missing keys: 0" in capsys.readouterr().out
    sd = m.state_dict()
    assert set(sd) == set(params)
    for k, v in params.items():
        assert sd[k].dtype == torch.bfloat16 and sd[k].device.type == "cpu" and torch.equal(sd[k], v), k
    # what the CLI and the pipeline read off the loaded model (fast_infer.py:350, pipeline_wan.py:634, 689)
    assert m.config.in_channels == CFG.to_kwargs()["in_dim"] and m.config.hidden_size == CFG.dim
    assert tuple(m.config.patch_size) == (1, 2, 2)
    assert m.freqs.dtype == torch.complex128 and tuple(m.freqs.shape) == (1024, 64) and m.freqs.device.type == "cpu"
    assert all(not p.requires_grad for p in m.parameters())
    assert kw["dict_mapping"] == {"in_dim": "in_channels", "dim": "hidden_size"}     # caller's dict left intact


def test_from_pretrained_reports_missing_and_unexpected(tmp_path, capsys):
    from safetensors.torch import save_file
    _checkpoint(tmp_path, drop="blocks.1.ffn.2.bias", sub=None)
    save_file({"not.a.parameter": torch.zeros(3)}, str(tmp_path / "zz_extra.safetensors"))
    m = WanTransformer3DModel.from_pretrained(str(tmp_path))
    out = capsys.readouterr().out
    assert "missing keys: 1" in out and "unexpected keys: 1" in out
    assert not m.state_dict()["blocks.1.ffn.2.bias"].any()


def test_from_pretrained_errors(tmp_path):
    with pytest.raises(RuntimeError, match="config.json"):
        WanTransformer3DModel.from_pretrained(str(tmp_path))
    (tmp_path / "config.json").write_text(json.dumps(CFG.to_kwargs()))
    with pytest.raises(RuntimeError, match="no weights"):
        WanTransformer3DModel.from_pretrained(str(tmp_path))
    (tmp_path / "config.json").write_text(json.dumps(dict(CFG.to_kwargs(), model_type="i2v")))
    with pytest.raises(NotImplementedError):
        WanTransformer3DModel.from_pretrained(str(tmp_path))


def test_riflex_tables_match_the_executed_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "riflex.npz"))
    rows = torch.from_numpy(g["rows"])
    m = WanTransformer3DModel(**CFG.to_kwargs())

    def check(name):
        f = m.freqs[rows]
        assert m.freqs.dtype == torch.complex128 and tuple(m.freqs.shape) == (1024, 64)
        assert np.array_equal(f.real.numpy(), g[name + "_re"]) and np.array_equal(f.imag.numpy(), g[name + "_im"]), name

    check("plain")
    m.enable_riflex()
    check("default")
    m.enable_riflex(k=4, L_test=30, L_test_scale=None)
    check("k4")
    m.enable_riflex(k=1, L_test=120, L_test_scale=2.0)
    check("k1_scaled")
    m.disable_riflex()
    check("plain")


def test_teacache_and_cfg_skip_switches():
    """Attribute protocol of :731-775 (the pipeline reads num_inference_steps / current_steps, pipeline_wan.py:692-695)."""
    a, b = WanTransformer3DModel(**CFG.to_kwargs()), WanTransformer3DModel(**CFG.to_kwargs())
    a.enable_teacache([1.0, 0.0], num_steps=10, rel_l1_thresh=0.1, num_skip_start_steps=2, offload=False)
    assert a.teacache is not None and a.teacache.num_steps == 10 and a.teacache.num_skip_start_steps == 2
    b.share_teacache(a)
    assert b.teacache is a.teacache
    a.disable_teacache()
    assert a.teacache is None and b.teacache is not None
    a.enable_cfg_skip(0.25, 50)
    assert (a.cfg_skip_ratio, a.current_steps, a.num_inference_steps) == (0.25, 0, 50)
    b.share_cfg_skip(a)
    assert (b.cfg_skip_ratio, b.current_steps, b.num_inference_steps) == (0.25, 0, 50)
    a.enable_cfg_skip(0, 50)                       # ratio 0 switches it off (:751-758)
    assert (a.cfg_skip_ratio, a.current_steps, a.num_inference_steps) == (None, 0, None)
    b.disable_cfg_skip()
    assert (b.cfg_skip_ratio, b.current_steps, b.num_inference_steps) == (None, 0, None)


@pytest.mark.parametrize("fmt", ["pth", "safetensors"])
def test_vae_from_pretrained_adds_the_model_prefix(tmp_path, fmt, capsys):
    """reference wan_vae.py:684-705 as the CLI calls it (fast_infer.py:300-303): one state-dict file whose keys lack the
    `model.` prefix, `additional_kwargs` = the whole vae_kwargs section of the yaml (config/wan2.1/wan_civitai.yaml)."""
    from oracle.vae_oracle import VAEConfig, make_vae_params
    from videocof_b200.vae import AutoencoderKLWan
    params = make_vae_params(VAEConfig(), seed=2)
    bare = {k[len("model."):]: v.contiguous() for k, v in params.items()}
    assert len(bare) == 194                                              # 194 tensors
    path = str(tmp_path / ("Wan2.1_VAE." + fmt))
    if fmt == "pth":
        torch.save(bare, path)
    else:
        from safetensors.torch import save_file
        save_file(bare, path)
    kw = {"vae_subpath": "Wan2.1_VAE.pth", "temporal_compression_ratio": 4, "spatial_compression_ratio": 8}
    vae = AutoencoderKLWan.from_pretrained(path, additional_kwargs=kw).to(torch.bfloat16)
    out = capsys.readouterr().out
    assert "missing keys: 0" in out and "unexpected keys: 0" in out
    sd = vae.state_dict()
    assert set(sd) == set(params)
    for k, v in params.items():
        assert sd[k].dtype == torch.bfloat16 and torch.equal(sd[k], v.to(torch.bfloat16)), k
    assert (vae.latent_channels, vae.temporal_compression_ratio, vae.spatial_compression_ratio) == (16, 4, 8)
    assert vae.config.latent_channels == 16 and vae.dtype == torch.bfloat16
