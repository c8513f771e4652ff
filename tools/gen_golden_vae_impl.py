"""VAE goldens: the reference's chunked AutoencoderKLWan.encode/.decode (wan_vae.py:620-682) on
deterministic parameters (oracle.vae_oracle.make_vae_params) and seeded inputs."""
import os

import numpy as np
import torch

import ref_loader
from gen_golden import GOLD, checksum

# name: (T frames, H, W)
VAE_CASES = {"vae_t9": (9, 32, 48), "vae_t1": (1, 32, 32), "vae_t13": (13, 16, 32)}


def vae_inputs(T, H, W, seed=7):
    g = torch.Generator().manual_seed(seed)
    video = torch.rand(3, T, H, W, generator=g) * 2 - 1
    f = (T - 1) // 4 + 1
    z = torch.randn(16, f, H // 8, W // 8, generator=g)
    return video, z


def gen_vae_impl():
    from oracle.vae_oracle import VAEConfig, make_vae_params
    ns = ref_loader.load_reference()
    cfg = VAEConfig()
    params = make_vae_params(cfg, seed=17)
    model = ns.vae.AutoencoderKLWan().eval()
    missing, unexpected = model.load_state_dict(params, strict=True)
    for name, (T, H, W) in VAE_CASES.items():
        video, z = vae_inputs(T, H, W)
        with torch.no_grad():
            post = model.encode(video[None])[0]
            mu = post.mode()[0]
            logvar = post.h.chunk(2, dim=1)[1][0]
            dec = model.decode(z[None]).sample[0]
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), param_checksum=np.float64(checksum(params)),
                            mu=mu.numpy(), logvar=logvar.numpy(), dec=dec.numpy())
        print("wrote", name, tuple(mu.shape), tuple(dec.shape))
