"""Golden RoPE tables of the reference's RIFLEx switch: executes the UNMODIFIED reference
`WanTransformer3DModel.enable_riflex` / `disable_riflex` (videox_fun/models/wan_transformer3d.py:776-800, built on
get_1d_rotary_pos_embed_riflex :57-114) on a tiny model and stores a few rows of the resulting complex128 `freqs`.
Runs only in the build container (/root/reference present):  python tools/gen_golden_riflex.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_loader  # noqa: E402

ROWS = [0, 1, 2, 17, 65, 66, 500, 1023]
CASES = {"default": dict(), "k4": dict(k=4, L_test=30, L_test_scale=None), "k1_scaled": dict(k=1, L_test=120, L_test_scale=2.0)}


def main():
    ref = ref_loader.load_reference()
    m = ref.dit.WanTransformer3DModel(dim=256, ffn_dim=512, num_heads=2, num_layers=1, text_dim=64, text_len=32)
    out = {"rows": np.array(ROWS)}
    base = m.freqs.clone()
    for name, kw in CASES.items():
        m.enable_riflex(**kw)
        f = m.freqs[ROWS]
        out[name + "_re"], out[name + "_im"] = f.real.numpy(), f.imag.numpy()
        assert f.dtype == torch.complex128 and tuple(m.freqs.shape) == (1024, 64)
    m.disable_riflex()
    assert torch.equal(m.freqs, base)
    out["plain_re"], out["plain_im"] = base[ROWS].real.numpy(), base[ROWS].imag.numpy()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "riflex.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
