"""Host-side logic of videocof_b200/vae.py (plan walking, weight packing, tap tables, parity views, first-frame
rules, frame interleave, nearest-2x folding) checked on CPU: the libvcof entry points are replaced by the
executable C-ABI statements of tests/vcof_emulator.py, and the result must match the goldens of the executed
reference within bf16-activation tolerance."""
import os

import numpy as np
import pytest
import torch

import vcof_emulator
from gen_golden_vae_impl import VAE_CASES, vae_inputs
from oracle.vae_oracle import VAEConfig, make_vae_params


@pytest.fixture(scope="module")
def model():
    from videocof_b200.vae import AutoencoderKLWan
    m = AutoencoderKLWan()
    m.load_state_dict(make_vae_params(VAEConfig(), seed=17), strict=True)
    return m.to(torch.bfloat16).eval()


def psnr(a, b):
    mse = float(((a.float() - b.float()) ** 2).mean())
    return 10 * np.log10(4.0 / max(mse, 1e-20))     # signal range [-1, 1]


@pytest.mark.parametrize("name", ["vae_t9", "vae_t1"])
def test_vae_host_logic_matches_reference(name, golden_dir, model, monkeypatch):
    vcof_emulator.install(monkeypatch)
    T, H, W = VAE_CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    video, z = vae_inputs(T, H, W)
    with torch.no_grad():
        post = model.encode(video[None].bfloat16())[0]
        dec = model.decode(z[None].bfloat16()).sample[0]
    mu = post.mode()[0].float()
    ref_mu, ref_dec = torch.from_numpy(gold["mu"]), torch.from_numpy(gold["dec"])
    assert mu.shape == ref_mu.shape and dec.shape == ref_dec.shape
    rel_mu = float((mu - ref_mu).norm() / ref_mu.norm())
    assert rel_mu < 2e-2, rel_mu
    assert psnr(dec, ref_dec) > 40.0, psnr(dec, ref_dec)


def test_state_dict_keys_match_reference_vae():
    from videocof_b200.vae import AutoencoderKLWan
    ours = {k: tuple(v.shape) for k, v in AutoencoderKLWan().state_dict().items()}
    ref = {k: tuple(v.shape) for k, v in make_vae_params(VAEConfig(), seed=0).items()}
    assert ours == ref      # make_vae_params loads strict=True into the reference class (tools/gen_golden.py)
    assert len(ours) == 194
