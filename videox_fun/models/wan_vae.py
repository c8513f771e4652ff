"""Overrides the reference module of the same name with the libvcof-backed VAE."""
from videocof_b200.vae import *  # noqa: F401,F403
from videocof_b200.vae import AutoencoderKLWan, AutoencoderKLWan_, CausalConv3d  # noqa: F401
