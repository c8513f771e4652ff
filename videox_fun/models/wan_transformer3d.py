"""Overrides the reference module of the same name with the libvcof-backed DiT."""
from videocof_b200.dit import *  # noqa: F401,F403
from videocof_b200.dit import WanTransformer3DModel, rope_params, sinusoidal_embedding_1d  # noqa: F401
