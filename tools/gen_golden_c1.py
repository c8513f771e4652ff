"""tests/golden/c1_full.npz: BASELINE.json configs[0] ("C1") END TO END from the UNMODIFIED reference, at full depth.

    python tools/gen_golden_c1.py            # ~2-3 minutes of CPU, 16 GB of RAM; needs /root/reference

The reference's own `WanPipeline.__call__` (videox_fun/pipeline/pipeline_wan.py:518-799, loaded by
tools/ref_loader.load_reference_pipeline) drives the reference's own 1.3B DiT (dim 1536, ffn 8960, 12 heads,
**30 layers**), the reference's 3D causal VAE and the reference's UniPC scheduler in fp32 on the CPU (SDPA backend):
5 source frames x 256 x 256 -> chain-of-frames latents [16, 5, 32, 32] (2 | 1 | 2) -> 4 steps (shift 3, guidance 1.0)
-> split decode (1 ground frame + 5 edit frames).  Parameters come from the oracle's seeded generators
(bf16-representable values, so the bf16 CUDA model holds exactly the same weights) and are regenerated from the seeds
by the tests; only inputs' seeds and OUTPUTS are stored:

    latents     [4, 1, 16, 5, 32, 32] fp32   latents after each scheduler step (callback_on_step_end)
    velocity    [4, 1, 16, 5, 32, 32] fp32   the DiT's output at each step (forward hook on the transformer)
    src_latent  [1, 16, 2, 32, 32]    fp32   VAE-encoded source clip (the first fs latent frames), before rounding
    init_latents[1, 16, 5, 32, 32]    fp32   [src | noise] rounded to bf16: what both runs start from (`latents=`)
    videos      [1, 3, 6, 256, 256]   fp16   decoded ground + edit frames in [0, 1]

Consumed by tests/test_widen_z_c1_depth_gpu.py (error growth over 30 layers x 4 steps; PSNR >= 40 dB) and
tests/test_oracle_golden.py::test_c1_full_first_forward (CPU, oracle vs the step-0 velocity).
"""
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
GOLD = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore")

DIT_SEED, VAE_SEED, INPUT_SEED, NOISE_SEED = 23, 17, 43, 7
SOURCE_FRAMES, H, W, N_CTX = 5, 256, 256, 77
CALL = dict(height=H, width=W, source_frames=SOURCE_FRAMES, reasoning_frames=4, num_inference_steps=4,
            guidance_scale=1.0, shift=3, repeat_rope=True, cot=True)


def c1_inputs():
    """Source clip in [-1, 1] and prompt embeddings, bf16-representable (identical on both sides of the comparison)."""
    g = torch.Generator().manual_seed(INPUT_SEED)
    video = (torch.rand(1, 3, SOURCE_FRAMES, H, W, generator=g) * 2 - 1).to(torch.bfloat16).float()
    embeds = torch.randn(1, N_CTX, 4096, generator=g).to(torch.bfloat16).float()
    return video, embeds


def main():
    import ref_loader
    from oracle.dit_oracle import DiTConfig, make_dit_params
    from oracle.vae_oracle import VAEConfig, make_vae_params
    t0 = time.time()
    ns = ref_loader.load_reference_pipeline()
    dcfg = DiTConfig.wan_1_3b()
    dit = ns.dit.WanTransformer3DModel(**dcfg.to_kwargs()).eval()
    dit.load_state_dict(make_dit_params(dcfg, seed=DIT_SEED), strict=True)
    vae = ns.vae.AutoencoderKLWan().eval()
    vae.load_state_dict(make_vae_params(VAEConfig(), seed=VAE_SEED), strict=True)
    sched = ns.unipc.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, solver_order=2)
    # prompt_embeds bypass the tokenizer / umT5 (pipeline_wan.py:230, 583-584); the call only reads text_encoder.dtype (:587)
    import types
    pipe = ns.pipeline.WanPipeline(None, types.SimpleNamespace(dtype=torch.float32), vae, dit, sched)
    print(f"models ready in {time.time() - t0:.0f} s")
    video, embeds = c1_inputs()
    lat, vel = [], []
    hook = dit.register_forward_hook(lambda _m, _a, out: vel.append(out.detach().float().numpy().copy()))

    def cb(_p, i, t, tensors):
        lat.append(tensors["latents"].detach().float().numpy().copy())
        return {}
    t0 = time.time()
    with torch.no_grad():
        # the initial latents [src | noise] exactly as prepare_cot_video_latents builds them (:397-418), then rounded to
        # bf16 and handed back through the pipeline's own `latents=` argument (:398-399) so that the fp32 reference run
        # and the bf16 CUDA run start from bit-identical latents (the CUDA path holds bf16 latents anyway)
        src = vae.encode(video)[0].mode()
        noise = ns.pipeline.randn_tensor((1, 16, src.shape[2] + 1) + tuple(src.shape[3:]),
                                         generator=torch.Generator().manual_seed(NOISE_SEED), device="cpu",
                                         dtype=torch.float32)
        init = torch.cat([src, noise], dim=2).to(torch.bfloat16).float()
        out = pipe(video=video, prompt_embeds=embeds, latents=init, callback_on_step_end=cb, **CALL)
    hook.remove()
    print(f"pipeline ran in {time.time() - t0:.0f} s")
    res = dict(latents=np.stack(lat), velocity=np.stack(vel), src_latent=src.float().numpy(), init_latents=init.numpy(),
               videos=np.asarray(out.videos).astype(np.float16),
               seeds=np.array([DIT_SEED, VAE_SEED, INPUT_SEED, NOISE_SEED]))
    np.savez_compressed(os.path.join(GOLD, "c1_full.npz"), **res)
    print("wrote c1_full", {k: v.shape for k, v in res.items()},
          os.path.getsize(os.path.join(GOLD, "c1_full.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
