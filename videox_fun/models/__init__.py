"""videox_fun.models — same export names as the reference (videox_fun/models/__init__.py:1-30) for the classes
the VideoCoF CLIs import (fast_infer.py:24-29); the DiT and the VAE are the libvcof-backed implementations."""
from videocof_b200.dit import (Head, WanAttentionBlock, WanLayerNorm, WanRMSNorm, WanSelfAttention,
                               WanT2VCrossAttention, WanTransformer3DModel)
from videocof_b200.vae import AutoencoderKLWan, AutoencoderKLWan_

from .. import extend_with_reference

extend_with_reference(__path__, "models")

try:  # tokenizer class comes from transformers exactly as in the reference
    from transformers import AutoTokenizer
except Exception:  # pragma: no cover - transformers missing
    AutoTokenizer = None


def __getattr__(name):
    """The umT5 text encoder is outside the hot path: import the reference's own file lazily."""
    if name == "WanT5EncoderModel":
        try:
            from .wan_text_encoder import WanT5EncoderModel
        except Exception as e:  # reference checkout (or its deps) not importable
            raise ImportError("WanT5EncoderModel is not re-implemented here: put the reference checkout on "
                              "PYTHONPATH after this repo, or set VIDEOCOF_REFERENCE_ROOT") from e
        return WanT5EncoderModel
    raise AttributeError(name)
