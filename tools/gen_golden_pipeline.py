"""tests/golden/pipeline_*.npz: the UNMODIFIED reference WanPipeline.__call__ (videox_fun/pipeline/pipeline_wan.py:518-799,
loaded by tools/ref_loader.load_reference_pipeline) driving the reference's own DiT, VAE, umT5 encoder and UniPC
scheduler — tiny, fp32, CPU, deterministic parameters from the oracle's generators — with a toy tokenizer.

    python tools/gen_golden_pipeline.py

Pins the glue of videocof_b200/pipeline.py (tests/test_pipeline_golden.py): prompt encoding and trimming, chain-of-frames
/ paired latent assembly, the noise draw, CFG batching and combination, the frozen source frames, the scheduler loop and
the split ground / edit decode."""
import os
import sys
import types
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
GOLD = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore")

DIT_KW = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=32)
T5_KW = dict(vocab=97, dim=64, dim_attn=64, dim_ffn=128, num_heads=4, num_layers=2, shared_pos=False)
MAX_LEN = 32
VIDEO = (1, 3, 9, 32, 48)                      # 9 source frames -> 3 latent frames

# name: kwargs of WanPipeline.__call__ on top of COMMON (prompt_embeds cases bypass the tokenizer / text encoder)
CASES = {
    "pipeline_cot": dict(prompt="make the cat blue", guidance_scale=1.0, cot=True),
    "pipeline_cot_cfg": dict(prompt="remove the lamp on the left of the sofa, keep everything else unchanged",
                             negative_prompt="blurry", guidance_scale=5.0, cot=True),
    "pipeline_paired_embeds": dict(prompt_embeds="seeded", guidance_scale=1.0, cot=False),
}
COMMON = dict(height=32, width=48, source_frames=9, reasoning_frames=4, num_inference_steps=4, shift=3,
              repeat_rope=True, max_sequence_length=MAX_LEN)


class ToyTokenizer:
    """Stands in for the umT5 sentencepiece tokenizer (no tokenizer files exist offline): one id per character,
    EOS = 1, pad = 0, with the HF call signature the pipeline uses (pipeline_wan.py:154-163)."""

    def __init__(self, vocab):
        self.vocab = vocab

    def __call__(self, prompt, padding="longest", max_length=None, truncation=False, add_special_tokens=True,
                 return_tensors="pt"):
        rows = [[2 + (ord(c) % (self.vocab - 2)) for c in s] + [1] for s in prompt]
        if truncation and max_length is not None:
            rows = [r[:max_length - 1] + [1] if len(r) > max_length else r for r in rows]
        width = max_length if padding == "max_length" else max(len(r) for r in rows)
        ids = torch.zeros(len(rows), width, dtype=torch.long)
        mask = torch.zeros(len(rows), width, dtype=torch.long)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = torch.tensor(r)
            mask[i, :len(r)] = 1
        return types.SimpleNamespace(input_ids=ids, attention_mask=mask)

    def batch_decode(self, ids):
        return ["".join(chr(int(t)) for t in row) for row in ids]


def pipeline_inputs():
    g = torch.Generator().manual_seed(41)
    video = torch.rand(*VIDEO, generator=g) * 2 - 1
    embeds = torch.randn(1, 11, DIT_KW["text_dim"], generator=g)    # the reference reads prompt_embeds.shape[0] (:584)
    return video, embeds


def run_case(pipe, kw):
    """-> dict of arrays: per-step latents (callback), videos / ground / edit."""
    video, embeds = pipeline_inputs()
    kw = dict(COMMON, **kw)
    if kw.get("prompt_embeds") == "seeded":
        kw["prompt_embeds"] = embeds
    steps = []

    def cb(_p, i, t, tensors):
        steps.append(tensors["latents"].detach().float().cpu().numpy().copy())
        return {}
    out = pipe(video=video, generator=torch.Generator().manual_seed(7), callback_on_step_end=cb, **kw)
    res = dict(latents=np.stack(steps), videos=np.asarray(out.videos))
    if out.ground_videos is not None:
        res["ground_videos"] = np.asarray(out.ground_videos)
    if out.edit_videos is not None:
        res["edit_videos"] = np.asarray(out.edit_videos)
    return res


def main():
    import ref_loader
    from oracle.dit_oracle import DiTConfig, make_dit_params
    from oracle.t5_oracle import T5Config, make_t5_params
    from oracle.vae_oracle import VAEConfig, make_vae_params
    ns = ref_loader.load_reference_pipeline()
    dcfg, tcfg = DiTConfig(**DIT_KW), T5Config(**T5_KW)
    dit = ns.dit.WanTransformer3DModel(**dcfg.to_kwargs()).eval()
    dit.load_state_dict(make_dit_params(dcfg, seed=11), strict=True)
    vae = ns.vae.AutoencoderKLWan().eval()
    vae.load_state_dict(make_vae_params(VAEConfig(), seed=17), strict=True)
    t5 = ns.t5.WanT5EncoderModel(**tcfg.to_kwargs()).eval()
    t5.load_state_dict(make_t5_params(tcfg, seed=19), strict=True)
    for name, kw in CASES.items():
        # the CLIs construct the scheduler with shift=1 and pass the real shift at call time (fast_infer.py:333-334)
        sched = ns.unipc.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, solver_order=2)
        pipe = ns.pipeline.WanPipeline(ToyTokenizer(tcfg.vocab), t5, vae, dit, sched)
        res = run_case(pipe, kw)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **res)
        print("wrote", name, {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    main()
