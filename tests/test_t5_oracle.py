"""oracle/t5_oracle.py against golden outputs of the UNMODIFIED reference text encoder
(videox_fun/models/wan_text_encoder.py executed by tools/gen_golden_t5.py); SURVEY.md §8f rank 3."""
import os

import numpy as np
import pytest
import torch

from gen_golden_t5 import T5_CASES, checksum, t5_inputs
from oracle.t5_oracle import T5Config, make_t5_params, relative_position_bucket, t5_forward


@pytest.mark.parametrize("name", list(T5_CASES))
def test_t5_oracle_matches_reference_golden(name, golden_dir):
    ckw, B, L, lens = T5_CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = T5Config(**ckw)
    params = make_t5_params(cfg, seed=19)
    assert checksum(params) == pytest.approx(float(gold["param_checksum"]), rel=1e-12)
    ids, mask = t5_inputs(cfg.vocab, B, L, lens)
    out = t5_forward(params, cfg, ids, mask)
    ref = torch.from_numpy(gold["out"])
    assert float((out - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


def test_bucket_table_is_bit_exact(golden_dir):
    """Integer work: the log-spaced buckets (wan_text_encoder.py:224-247) for every offset in [-300, 300]."""
    gold = np.load(os.path.join(golden_dir, "t5_tiny.npz"))
    b = relative_position_bucket(torch.arange(-300, 301))
    assert np.array_equal(b.numpy(), gold["buckets"])
    assert int(b.min()) == 0 and int(b.max()) == 31


def test_masked_keys_do_not_leak(golden_dir):
    """A prefix mask makes the valid rows independent of what sits in the padding (wan_text_encoder.py:94-98)."""
    ckw, B, L, lens = T5_CASES["t5_tiny"]
    cfg = T5Config(**ckw)
    params = make_t5_params(cfg, seed=19)
    ids, mask = t5_inputs(cfg.vocab, B, L, lens)
    a = t5_forward(params, cfg, ids, mask)
    ids2 = ids.clone()
    ids2[1, lens[1]:] = 5
    b = t5_forward(params, cfg, ids2, mask)
    assert torch.equal(a[1, :lens[1]], b[1, :lens[1]])
    short = t5_forward(params, cfg, ids[1:2, :lens[1]], None)        # no padding at all: same rows
    assert float((short[0] - a[1, :lens[1]]).abs().max()) < 1e-5


def test_bf16_emulation_close_to_fp32_gold():
    ckw, B, L, lens = T5_CASES["t5_d64"]
    cfg = T5Config(**ckw)
    params = make_t5_params(cfg, seed=19)
    ids, mask = t5_inputs(cfg.vocab, B, L, lens)
    a = t5_forward(params, cfg, ids, mask)
    b = t5_forward(params, cfg, ids, mask, emulate_bf16=True)
    rel = float((a - b).norm() / a.norm())
    assert rel < 2e-2, rel
