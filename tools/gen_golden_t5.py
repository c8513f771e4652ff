"""tests/golden/t5_*.npz: the UNMODIFIED reference text encoder (videox_fun/models/wan_text_encoder.py, loaded by
tools/ref_loader.py) on the deterministic parameters of oracle.t5_oracle.make_t5_params and seeded token ids.

    python tools/gen_golden_t5.py

Pins oracle/t5_oracle.py (tests/test_t5_oracle.py).  Parameters are regenerated from the seed, not stored; each
fixture carries their checksum."""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
GOLD = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore")

# name: (config kwargs, batch, tokens, valid lengths per sample or None for no mask)
T5_CASES = {
    "t5_tiny": (dict(vocab=97, dim=64, dim_attn=64, dim_ffn=128, num_heads=4, num_layers=2, shared_pos=False),
                2, 160, [160, 13]),
    "t5_tiny_shared": (dict(vocab=97, dim=64, dim_attn=64, dim_ffn=128, num_heads=4, num_layers=2, shared_pos=True),
                       1, 24, None),
    "t5_d64": (dict(vocab=211, dim=256, dim_attn=256, dim_ffn=512, num_heads=4, num_layers=2, shared_pos=False),
               2, 96, [96, 1]),
}


def t5_inputs(vocab, B, L, lens, seed=31):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, vocab, (B, L), generator=g)
    mask = None
    if lens is not None:
        mask = torch.zeros(B, L, dtype=torch.long)
        for b, n in enumerate(lens):
            mask[b, :n] = 1
            ids[b, n:] = 0                                     # the tokenizer's pad id
    return ids, mask


def checksum(params):
    return float(sum(float(v.double().abs().sum()) * (1 + (i % 7)) for i, (k, v) in enumerate(sorted(params.items()))))


def main():
    import ref_loader
    from oracle.t5_oracle import T5Config, make_t5_params
    ns = ref_loader.load_reference_pipeline()
    for name, (ckw, B, L, lens) in T5_CASES.items():
        cfg = T5Config(**ckw)
        params = make_t5_params(cfg, seed=19)
        model = ns.t5.WanT5EncoderModel(**cfg.to_kwargs()).eval()
        model.load_state_dict(params, strict=True)
        ids, mask = t5_inputs(cfg.vocab, B, L, lens)
        with torch.no_grad():
            out = model(ids, attention_mask=mask)[0]
            rel = torch.arange(-300, 301)
            buckets = model.blocks[0].pos_embedding._relative_position_bucket(rel) if not cfg.shared_pos \
                else model.pos_embedding._relative_position_bucket(rel)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), param_checksum=np.float64(checksum(params)),
                            out=out.float().numpy(), buckets=buckets.numpy())
        print("wrote", name, tuple(out.shape), float(out.abs().mean()))


if __name__ == "__main__":
    main()
