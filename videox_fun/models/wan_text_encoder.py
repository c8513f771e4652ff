"""Overrides the reference module of the same name with the libvcof-backed umT5 encoder."""
from videocof_b200.text_encoder import *  # noqa: F401,F403
from videocof_b200.text_encoder import WanT5EncoderModel  # noqa: F401
