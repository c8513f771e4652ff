"""The call sequence of the reference CLI (fast_infer.py:281-449 — load three checkpoints, build the scheduler from the
yaml, WanPipeline, the default "sequential_cpu_offload" memory mode, merge a LoRA, pipeline(...).videos, slice the edit
segment, save, unmerge) replayed step by step against the overlay package, with tiny checkpoints on disk and the libvcof
entry points replaced by their contract statements (tests/vcof_emulator.py).  Each model class is pinned on its own
elsewhere; this test is about the seams between them as the CLI exercises them — that `modulation` survives
`replace_parameters_by_name`, that a merged LoRA reaches the forward, that the result has the layout the CLI slices."""
import json
import sys
import types

import numpy as np
import pytest
import torch

import vcof_emulator
from oracle.dit_oracle import DiTConfig, make_dit_params
from oracle.lora_oracle import make_lora_state
from oracle.t5_oracle import T5Config, make_t5_params
from oracle.vae_oracle import VAEConfig, make_vae_params

DIT = DiTConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=32, text_len=64)
T5 = T5Config(vocab=64, dim=32, dim_attn=32, dim_ffn=64, num_heads=2, num_layers=2, shared_pos=False)
YAML = {      # config/wan2.1/wan_civitai.yaml with the tiny widths substituted
    "transformer_additional_kwargs": {"transformer_subpath": "./", "dict_mapping": {"in_dim": "in_channels", "dim": "hidden_size"}},
    "vae_kwargs": {"vae_subpath": "Wan2.1_VAE.pth", "temporal_compression_ratio": 4, "spatial_compression_ratio": 8},
    "text_encoder_kwargs": dict(T5.to_kwargs(), text_encoder_subpath="t5.pth", tokenizer_subpath="tok", text_length=16),
    "scheduler_kwargs": {"scheduler_subpath": None, "num_train_timesteps": 1000, "shift": 5.0, "use_dynamic_shifting": False,
                         "base_shift": 0.5, "max_shift": 1.15, "base_image_seq_len": 256, "max_image_seq_len": 4096},
}


class ToyTokenizer:
    """Stands in for AutoTokenizer.from_pretrained(google/umt5-xxl): same call signature and result fields."""

    def __call__(self, prompt, padding=None, max_length=None, truncation=None, add_special_tokens=None, return_tensors=None):
        ids = torch.zeros(len(prompt), max_length, dtype=torch.long)
        mask = torch.zeros(len(prompt), max_length, dtype=torch.long)
        for i, text in enumerate(prompt):
            toks = [2 + (sum(map(ord, w)) % 60) for w in text.split()][:max_length - 1] + [1]
            ids[i, :len(toks)] = torch.tensor(toks)
            mask[i, :len(toks)] = 1
        return types.SimpleNamespace(input_ids=ids, attention_mask=mask)


def replace_parameters_by_name(module, name_keywords, device):
    """videox_fun/utils/fp8_optimization.py:8-17 restated (the overlay forwards that module to the reference checkout)."""
    for name, param in list(module.named_parameters(recurse=False)):
        if any(k in name for k in name_keywords) and isinstance(param, torch.nn.Parameter):
            tensor = param.data
            delattr(module, name)
            setattr(module, name, tensor.to(device=device))
    for child in module.children():
        replace_parameters_by_name(child, name_keywords, device)


def _contract_apply(weight, up, down, scale, device):
    weight.copy_((weight.float() + scale * (up.float() @ down.float())).to(torch.bfloat16))


@pytest.fixture
def model_dir(tmp_path):
    from safetensors.torch import save_file
    dit = {k: v.to(torch.bfloat16).contiguous() for k, v in make_dit_params(DIT, seed=21).items()}
    keys = sorted(dit)
    for i in range(2):
        save_file({k: dit[k] for k in keys[i::2]}, str(tmp_path / f"diffusion_pytorch_model-0000{i + 1}-of-00002.safetensors"))
    (tmp_path / "config.json").write_text(json.dumps(dict(DIT.to_kwargs(), _class_name="WanModel")))
    torch.save({k[len("model."):]: v for k, v in make_vae_params(VAEConfig(), seed=17).items()}, tmp_path / "Wan2.1_VAE.pth")
    torch.save(make_t5_params(T5, seed=19), tmp_path / "t5.pth")
    shapes = {k[:-7]: tuple(v.shape) for k, v in dit.items()
              if k.endswith(".weight") and v.dim() == 2 and ".norm" not in k and k.startswith("blocks.")}
    lora = {k: v.contiguous() for k, v in make_lora_state(shapes, rank=4, seed=8).items() if not k.startswith("lora_te")}
    (tmp_path / "loras").mkdir()
    save_file(lora, str(tmp_path / "loras" / "videocof.safetensors"))
    return tmp_path


def test_fast_infer_sequence(model_dir, monkeypatch, capsys):
    import os
    # fast_infer.py:24-40 — the CLI's own import lines, answered by the overlay
    from videox_fun.models import AutoencoderKLWan, WanT5EncoderModel, WanTransformer3DModel
    from videox_fun.pipeline import WanPipeline
    from videox_fun.utils.fm_solvers_unipc import FlowUniPCMultistepScheduler
    from videox_fun.utils.lora_utils import merge_lora, unmerge_lora
    from videox_fun.utils.utils import filter_kwargs
    from videocof_b200 import lora as lora_mod
    vcof_emulator.install(monkeypatch)
    vcof_emulator.install_dit(monkeypatch)
    vcof_emulator.install_t5(monkeypatch)
    monkeypatch.setattr(lora_mod, "_apply", _contract_apply)
    device, weight_dtype, model_name = torch.device("cpu"), torch.bfloat16, str(model_dir)
    config = YAML

    transformer = WanTransformer3DModel.from_pretrained(                                     # :281-286
        os.path.join(model_name, config["transformer_additional_kwargs"].get("transformer_subpath", "transformer")),
        transformer_additional_kwargs=dict(config["transformer_additional_kwargs"]), low_cpu_mem_usage=True,
        torch_dtype=weight_dtype)
    vae = AutoencoderKLWan.from_pretrained(                                                  # :300-303
        os.path.join(model_name, config["vae_kwargs"].get("vae_subpath", "vae")),
        additional_kwargs=dict(config["vae_kwargs"])).to(weight_dtype)
    tokenizer = ToyTokenizer()                                                               # :317-319
    text_encoder = WanT5EncoderModel.from_pretrained(                                        # :321-326
        os.path.join(model_name, config["text_encoder_kwargs"].get("text_encoder_subpath", "text_encoder")),
        additional_kwargs=dict(config["text_encoder_kwargs"]), low_cpu_mem_usage=True, torch_dtype=weight_dtype)
    sk = dict(config["scheduler_kwargs"], shift=1)                                           # :333-334
    scheduler = FlowUniPCMultistepScheduler(**filter_kwargs(FlowUniPCMultistepScheduler, sk))  # :335-337
    pipeline = WanPipeline(transformer=transformer, vae=vae, tokenizer=tokenizer, text_encoder=text_encoder,
                           scheduler=scheduler)                                              # :339-345
    replace_parameters_by_name(transformer, ["modulation"], device=device)                   # :349
    assert not isinstance(transformer.blocks[0].modulation, torch.nn.Parameter)
    assert all("modulation" not in n for n, _ in transformer.named_parameters())
    transformer.freqs = transformer.freqs.to(device=device)                                  # :350
    pipeline.enable_sequential_cpu_offload(device=device)                                    # :351

    g = torch.Generator().manual_seed(4)
    frames = torch.randint(0, 256, (9, 32, 48, 3), generator=g, dtype=torch.uint8)
    input_video = (frames.permute(3, 0, 1, 2).float() * (2.0 / 255.0) - 1.0).unsqueeze(0)    # load_video_frames, :86-91
    prompt = "A video sequence showing three parts: first the original scene, then grounded the cup, and finally " \
             "the same scene but remove the cup"                                             # :409-412

    def run():
        with torch.no_grad():
            return pipeline(video=input_video, prompt=prompt, num_frames=17, source_frames=9, reasoning_frames=4,
                            negative_prompt="blurred details, static", height=32, width=48,
                            generator=torch.Generator(device=device).manual_seed(0), guidance_scale=1.0,
                            num_inference_steps=4, shift=3, repeat_rope=True, cot=True).videos   # :420-435

    plain = run()
    before = {k: v.clone() for k, v in transformer.state_dict().items()}
    pipeline = merge_lora(pipeline, str(model_dir / "loras" / "videocof.safetensors"), 1.0, device=device)   # :384-385
    assert isinstance(pipeline, WanPipeline)
    changed = [k for k, v in transformer.state_dict().items() if not torch.equal(v, before[k])]
    assert len(changed) >= 20 and all(k.endswith(".weight") for k in changed)
    sample = run()

    # what the CLI does with the result (:437-446)
    assert isinstance(sample, torch.Tensor) and sample.dtype == torch.float32
    assert tuple(sample.shape) == (1, 3, 1 + 9, 32, 48)                    # grounding frame + edit segment
    assert 0.0 <= float(sample.min()) and float(sample.max()) <= 1.0
    edit_video = sample[:, :, -9:, :, :]
    assert tuple(edit_video.shape) == (1, 3, 9, 32, 48)
    assert not torch.equal(sample, plain)                                   # the merged LoRA reached the forward
    arr = (edit_video[0].numpy() * 255).astype(np.uint8)                    # save_results' conversion
    assert arr.shape == (3, 9, 32, 48)

    pipeline = unmerge_lora(pipeline, str(model_dir / "loras" / "videocof.safetensors"), 1.0, device=device)   # :448-449
    after = transformer.state_dict()
    worst = max(float((after[k].float() - before[k].float()).abs().max()) for k in changed)
    assert worst <= 2.0 ** -7 * max(float(before[k].float().abs().max()) for k in changed)   # back to within a bf16 ulp
    assert "videox_fun" in sys.modules
