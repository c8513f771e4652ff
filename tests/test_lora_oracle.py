"""oracle/lora_oracle.py against the golden produced by the executed reference merge_lora / unmerge_lora
(tools/gen_golden_lora.py; lora_utils.py:371-618): bit-exact bf16 weights after merge and after unmerge."""
import os
import zlib

import numpy as np
import torch

from oracle.dit_oracle import DiTConfig, make_dit_params
from oracle.lora_oracle import make_lora_state, merge, normalise_keys, resolve

GOLD = os.path.join(os.path.dirname(__file__), "golden", "lora_tiny.npz")
CFG = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=32)
MULT, RANK = 0.8, 8


def _bits(t):
    return t.detach().contiguous().view(torch.int16).numpy()


def _setup():
    params = make_dit_params(DiTConfig(**CFG), seed=11)
    shapes = {k[:-7]: tuple(v.shape) for k, v in params.items()
              if k.endswith(".weight") and v.dim() == 2 and ".norm" not in k and k.startswith("blocks.")}
    weights = {k[:-7]: v.to(torch.bfloat16) for k, v in params.items() if k.endswith(".weight")}
    return weights, make_lora_state(shapes, rank=RANK, seed=5)


def test_key_normalisation_and_resolution():
    weights, sd = _setup()
    upd = normalise_keys(sd)
    assert "lora_unet__blocks_0_self_attn_q" in upd and "alpha" in upd["lora_unet__blocks_0_self_attn_q"]
    assert "lora_unet__blocks_1_ffn_2" in upd and "alpha" not in upd["lora_unet__blocks_1_ffn_2"]
    assert resolve("lora_unet__blocks_1_cross_attn_k", list(weights)) == "blocks.1.cross_attn.k"
    assert resolve("lora_unet__blocks_99_self_attn_q", list(weights)) is None
    assert resolve("lora_unet__blocks_0", list(weights)) is None            # the norm-only entry lands on the block


def test_merge_and_unmerge_match_reference_bits():
    g = np.load(GOLD)
    weights, sd = _setup()
    touched = merge(weights, sd, MULT)
    changed = [str(k) for k in g["changed"]]
    assert sorted(t + ".weight" for t in touched) == changed
    for k, crc in zip(changed, g["crc_merged"]):
        assert zlib.crc32(_bits(weights[k[:-7]]).tobytes()) == int(crc), k
    for key in g.files:
        if key.startswith("merged/"):
            assert np.array_equal(_bits(weights[key[len("merged/"):-7]]), g[key]), key
    merge(weights, {k: v for k, v in sd.items() if not k.startswith("lora_te")}, MULT, sign=-1.0)
    for k, crc in zip(changed, g["crc_unmerged"]):
        assert zlib.crc32(_bits(weights[k[:-7]]).tobytes()) == int(crc), k


def test_text_encoder_entries_match_reference_bits():
    """`lora_te_…` tokens resolve under pipeline.text_encoder (lora_utils.py:406-411); golden from the executed
    reference on its own umT5 encoder (tools/gen_golden_lora_te.py)."""
    from gen_golden_lora_te import MULT as M_TE, RANK as R_TE, te_linear_shapes
    from gen_golden_pipeline import T5_KW
    from oracle.lora_oracle import make_te_lora_state
    from oracle.t5_oracle import T5Config, make_t5_params
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lora_te_tiny.npz"))
    params = make_t5_params(T5Config(**T5_KW), seed=19)
    sd = make_te_lora_state(te_linear_shapes(params), rank=R_TE, seed=9)
    weights = {k[:-7]: v.to(torch.bfloat16) for k, v in params.items() if k.endswith(".weight")}
    touched = merge(weights, sd, M_TE, text_encoder=True)
    changed = [str(k) for k in g["changed"]]
    assert sorted(t + ".weight" for t in touched) == changed
    for k, crc in zip(changed, g["crc_merged"]):
        assert zlib.crc32(_bits(weights[k[:-7]]).tobytes()) == int(crc), k
    merge(weights, sd, M_TE, sign=-1.0, text_encoder=True)
    for k, crc in zip(changed, g["crc_unmerged"]):
        assert zlib.crc32(_bits(weights[k[:-7]]).tobytes()) == int(crc), k
    assert merge(dict(weights), sd, M_TE) == []                       # the DiT search ignores lora_te entries
