"""The CPU oracle against golden outputs of the UNMODIFIED reference (tools/gen_golden.py).

These fixtures are what pins oracle/: the reference ships no tests or golden vectors of its own
(SURVEY.md §4), so the goldens were produced by executing its files in the build container.
"""
import os

import numpy as np
import pytest
import torch

from gen_golden import DIT_CASES, ROPE_MODES, checksum, dit_inputs
from oracle.dit_oracle import DiTConfig, dit_forward, make_dit_params, temporal_positions


@pytest.mark.parametrize("name", list(DIT_CASES))
def test_dit_oracle_matches_reference_golden(name, golden_dir):
    ckw, shape, n_ctx, B = DIT_CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = DiTConfig(**ckw)
    params = make_dit_params(cfg, seed=11)
    assert checksum(params) == pytest.approx(float(gold["param_checksum"]), rel=1e-12), \
        "parameter generator drifted from the one the golden was made with"
    x, ctx, t = dit_inputs(shape, n_ctx, cfg.text_dim, B, seed=23)
    f = shape[1]
    seq_len = f * (shape[2] // 2) * (shape[3] // 2)
    for mode, mk in ROPE_MODES.items():
        key = "out_" + mode
        if key not in gold.files:
            continue
        y = dit_forward(params, cfg, x, t, ctx, seq_len, **mk(f, B))
        ref = torch.from_numpy(gold[key])
        err = float((y - ref).abs().max())
        assert err < 2e-5 * max(1.0, float(ref.abs().max())), (name, mode, err)


def test_temporal_positions_modes():
    # wan_transformer3d.py:153-191: plain / paired / chain-of-frames
    assert temporal_positions(5) == [0, 1, 2, 3, 4]
    assert temporal_positions(5, 2) == [0, 1, 0, 1, 2]
    assert temporal_positions(5, 2, (2, 3)) == [1, 2, 0, 1, 2]
    assert temporal_positions(21, 10, (10, 11)) == list(range(1, 11)) + [0] + list(range(1, 11))


def test_bf16_emulation_close_to_fp32_gold():
    ckw, shape, n_ctx, B = DIT_CASES["dit_tiny"]
    cfg = DiTConfig(**ckw)
    params = make_dit_params(cfg, seed=11)
    x, ctx, t = dit_inputs(shape, n_ctx, cfg.text_dim, B, seed=23)
    seq_len = shape[1] * (shape[2] // 2) * (shape[3] // 2)
    a = dit_forward(params, cfg, x, t, ctx, seq_len)
    b = dit_forward(params, cfg, x.bfloat16().float(), t, [c.bfloat16().float() for c in ctx], seq_len,
                    emulate_bf16=True)
    rel = float((a - b).norm() / a.norm())
    assert rel < 3e-2, rel


# ---------------------------------------------------------------------------------------------
# VAE: the un-chunked closed form of the oracle vs the reference's chunked streaming loop
# ---------------------------------------------------------------------------------------------
from gen_golden_vae_impl import VAE_CASES, vae_inputs  # noqa: E402
from oracle.vae_oracle import VAEConfig, make_vae_params, vae_decode, vae_encode  # noqa: E402


@pytest.fixture(scope="module")
def vae_params():
    return make_vae_params(VAEConfig(), seed=17)


@pytest.mark.parametrize("name", list(VAE_CASES))
def test_vae_oracle_matches_reference_golden(name, golden_dir, vae_params):
    T, H, W = VAE_CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    assert checksum(vae_params) == pytest.approx(float(gold["param_checksum"]), rel=1e-12)
    video, z = vae_inputs(T, H, W)
    cfg = VAEConfig()
    mu, logvar = vae_encode(vae_params, cfg, video)
    dec = vae_decode(vae_params, cfg, z)
    for got, key in ((mu, "mu"), (logvar, "logvar"), (dec, "dec")):
        ref = torch.from_numpy(gold[key])
        assert got.shape == ref.shape, (key, got.shape, ref.shape)
        err = float((got - ref).abs().max())
        assert err < 5e-5 * max(1.0, float(ref.abs().max())), (name, key, err)


def test_vae_decode_is_strictly_causal(vae_params):
    """decode(z[:, :k]) equals the first 4(k-1)+1 frames of decode(z)."""
    cfg = VAEConfig()
    _, z = vae_inputs(13, 16, 16)
    full = vae_decode(vae_params, cfg, z)
    part = vae_decode(vae_params, cfg, z[:, :2])
    assert torch.allclose(full[:, :5], part, atol=1e-5)


def test_c1_full_first_forward(golden_dir):
    """The oracle at FULL depth (1.3B widths, 30 layers, chain of frames) against the first DiT forward of the executed
    reference pipeline run of tools/gen_golden_c1.py (tests/golden/c1_full.npz)."""
    from gen_golden_c1 import DIT_SEED, c1_inputs
    gold = np.load(os.path.join(golden_dir, "c1_full.npz"))
    cfg = DiTConfig.wan_1_3b()
    params = make_dit_params(cfg, seed=DIT_SEED)
    _, embeds = c1_inputs()
    x = torch.from_numpy(gold["init_latents"])
    with torch.no_grad():
        y = dit_forward(params, cfg, x, torch.tensor([999.0]), [embeds[0]], 1280, frame_split_indices=[2],
                        ground_frame_indices=[(2, 3)])
    want = torch.from_numpy(gold["velocity"][0])
    assert float((y - want).norm() / want.norm()) < 1e-4
