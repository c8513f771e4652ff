"""The 4-step fast_infer.py path end to end on the GPU (WanPipeline: VAE encode -> chain-of-frames latents -> 4 UniPC
steps of the DiT -> split VAE decode) against the same path assembled from the CPU oracles in fp32.

north_star parity bar: pixel PSNR >= 40 dB on the 4-step path; latent error reported as relative Frobenius.
Weights are random (tiny DiT widths, the real VAE architecture); the noise is drawn from a CPU generator so both
sides see identical latents."""
import numpy as np
import pytest
import torch

from oracle.dit_oracle import DiTConfig, dit_forward, make_dit_params
from oracle.vae_oracle import VAEConfig, make_vae_params, vae_decode, vae_encode

pytestmark = pytest.mark.gpu


def oracle_pipeline(dit_p, dit_cfg, vae_p, video, ctx, steps, shift, seed):
    """pipeline_wan.py:381-419, 613-637, 689-740, 757-777 restated on the oracles (guidance 1.0, cot=True)."""
    from videocof_b200.scheduler import FlowUniPCMultistepScheduler      # host scheduler is pinned bit-exact
    vcfg = VAEConfig()
    src, _ = vae_encode(vae_p, vcfg, video)
    fs = src.shape[1]
    g = torch.Generator().manual_seed(seed)
    noise = torch.randn((1, 16, fs + 1) + tuple(src.shape[2:]), generator=g, dtype=torch.bfloat16).float()
    lat = torch.cat([src[None], noise], dim=2)
    sched = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, solver_order=2)
    sched.set_timesteps(steps, device="cpu", shift=shift)
    f, h, w = lat.shape[2:]
    seq_len = f * (h // 2) * (w // 2)
    for t in sched.timesteps:
        v = dit_forward(dit_p, dit_cfg, lat, t.expand(1).float(), ctx, seq_len, frame_split_indices=[fs],
                        ground_frame_indices=[(fs, fs + 1)])
        v[:, :, :fs] = 0
        lat = sched.step(v, t, lat, return_dict=False)[0]
    ground = vae_decode(vae_p, vcfg, lat[0, :, fs:fs + 1])
    edit = vae_decode(vae_p, vcfg, lat[0, :, fs + 1:])
    return lat, (torch.cat([ground, edit], dim=1) / 2 + 0.5).clamp(0, 1)


def test_four_step_cot_pipeline_psnr():
    from videocof_b200.dit import WanTransformer3DModel
    from videocof_b200.pipeline import WanPipeline
    from videocof_b200.scheduler import FlowUniPCMultistepScheduler
    from videocof_b200.vae import AutoencoderKLWan
    dcfg = DiTConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=32)
    dp = make_dit_params(dcfg, seed=5)
    vp = make_vae_params(VAEConfig(), seed=17)
    g = torch.Generator().manual_seed(2)
    video = (torch.rand(3, 9, 32, 48, generator=g) * 2 - 1).bfloat16().float()
    ctx = [torch.randn(6, 64, generator=g).bfloat16().float()]

    want_lat, want_pix = oracle_pipeline(dp, dcfg, vp, video, ctx, steps=4, shift=3.0, seed=11)

    dit = WanTransformer3DModel(**dcfg.to_kwargs())
    dit.load_state_dict(dp, strict=True)
    vae = AutoencoderKLWan()
    vae.load_state_dict(vp, strict=True)
    dit, vae = dit.to("cuda", torch.bfloat16).eval(), vae.to("cuda", torch.bfloat16).eval()
    pipe = WanPipeline(None, None, vae, dit, FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1))
    lat_seen = {}

    def cb(p, i, t, kw):
        lat_seen[i] = kw["latents"]
        return {}
    out = pipe(video=video[None].cuda(), prompt_embeds=[c.cuda().bfloat16() for c in ctx], height=32, width=48,
               source_frames=9, reasoning_frames=4, num_inference_steps=4, guidance_scale=1.0, shift=3, repeat_rope=True,
               cot=True, generator=torch.Generator().manual_seed(11), callback_on_step_end=cb)
    got_pix = out.videos[0].float()
    got_lat = lat_seen[3].float().cpu()
    rel_lat = float((got_lat - want_lat).norm() / want_lat.norm())
    mse = float(((got_pix - want_pix) ** 2).mean())
    psnr = 10 * np.log10(1.0 / max(mse, 1e-20))          # frames in [0, 1]
    assert tuple(got_pix.shape) == tuple(want_pix.shape) == (3, 10, 32, 48)
    assert rel_lat < 3e-2, rel_lat
    assert psnr >= 40.0, psnr
