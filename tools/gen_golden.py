"""Generate tests/golden/*.npz by running the UNMODIFIED reference (from /root/reference, via
tools/ref_loader.py) on deterministic parameters and inputs.  Run in the build container:

    python tools/gen_golden.py [dit] [vae] [unipc]

The fixtures pin oracle/ (tests/test_oracle_golden.py); parameters are NOT stored — they are
regenerated from `oracle.*.make_*_params(cfg, seed)` (bf16-exact values from a seeded CPU
torch.Generator), and each fixture carries a parameter checksum so drift is detected.
"""
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

import ref_loader  # noqa: E402


def checksum(params):
    """Order-independent fp64 checksum of a parameter dict."""
    return float(sum(float(v.double().abs().sum()) * (1 + (i % 7)) for i, (k, v) in
                     enumerate(sorted(params.items()))))


DIT_CASES = {
    # name: (cfg kwargs, latent shape, n_ctx, batch, seq_pad)
    "dit_tiny": (dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=32),
                 (16, 5, 8, 12), 7, 1),
    "dit_tiny_b2": (dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=32),
                    (16, 3, 6, 10), 9, 2),
    "dit_c1_2layer": (dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=2, text_dim=4096, text_len=512),
                      (16, 5, 32, 32), 77, 1),
}
ROPE_MODES = {
    "plain": lambda f, B: {},
    "paired": lambda f, B: dict(frame_split_indices=[f // 2] * B),
    "cot": lambda f, B: dict(frame_split_indices=[f // 2] * B,
                             ground_frame_indices=[(f // 2, f // 2 + 1)] * B),
}


def dit_inputs(shape, n_ctx, text_dim, B, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, *shape, generator=g)
    ctx = [torch.randn(n_ctx + 3 * i, text_dim, generator=g) for i in range(B)]
    t = torch.tensor([749.0, 499.0][:B])
    return x, ctx, t


def gen_dit():
    from oracle.dit_oracle import DiTConfig, make_dit_params
    ns = ref_loader.load_reference()
    for name, (ckw, shape, n_ctx, B) in DIT_CASES.items():
        cfg = DiTConfig(**ckw)
        params = make_dit_params(cfg, seed=11)
        model = ns.dit.WanTransformer3DModel(**cfg.to_kwargs()).eval()
        model.load_state_dict(params, strict=True)
        x, ctx, t = dit_inputs(shape, n_ctx, cfg.text_dim, B, seed=23)
        f = shape[1]
        seq_len = f * (shape[2] // 2) * (shape[3] // 2)
        out = {}
        for mode, mk in ROPE_MODES.items():
            if name == "dit_c1_2layer" and mode != "cot":
                continue
            with torch.no_grad():
                y = model(x=x, t=t, context=ctx, seq_len=seq_len, **mk(f, B))
            out["out_" + mode] = y.float().numpy()
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), param_checksum=np.float64(checksum(params)),
                            x=x.numpy() if name != "dit_c1_2layer" else np.zeros(0, np.float32), **out)
        print("wrote", name, {k: v.shape for k, v in out.items()})


TEACACHE = dict(coefficients=[-5.21862437e+04, 9.23041404e+03, -5.28275948e+02, 1.36987616e+01, -4.99875664e-02],
                num_steps=4, rel_l1_thresh=0.6, num_skip_start_steps=1, offload=False)
TEACACHE_T = [999.0, 899.0, 749.0, 499.0]


def gen_teacache():
    """Four consecutive forwards (the 4-step schedule) with TeaCache on: records the outputs and which steps the
    reference skipped (wan_transformer3d.py:956-1031, cache_utils.py:21-76)."""
    from oracle.dit_oracle import DiTConfig, make_dit_params
    ns = ref_loader.load_reference()
    ckw, shape, n_ctx, B = DIT_CASES["dit_tiny"]
    cfg = DiTConfig(**ckw)
    params = make_dit_params(cfg, seed=11)
    model = ns.dit.WanTransformer3DModel(**cfg.to_kwargs()).eval()
    model.load_state_dict(params, strict=True)
    model.enable_teacache(**TEACACHE)
    x, ctx, _ = dit_inputs(shape, n_ctx, cfg.text_dim, B, seed=23)
    f = shape[1]
    seq_len = f * (shape[2] // 2) * (shape[3] // 2)
    outs, calc = [], []
    for tv in TEACACHE_T:
        with torch.no_grad():
            y = model(x=x, t=torch.tensor([tv]), context=ctx, seq_len=seq_len, **ROPE_MODES["cot"](f, B))
        outs.append(y.float().numpy())
        calc.append(bool(model.should_calc))
    np.savez_compressed(os.path.join(GOLD, "dit_tiny_teacache.npz"), outs=np.stack(outs), should_calc=np.array(calc))
    print("wrote dit_tiny_teacache should_calc =", calc)


def main():
    os.makedirs(GOLD, exist_ok=True)
    what = set(sys.argv[1:]) or {"dit", "vae", "unipc"}
    if "dit" in what:
        gen_dit()
    if "teacache" in what or "dit" in what:
        gen_teacache()
    if "vae" in what:
        from gen_golden_vae import gen_vae
        gen_vae()
    if "unipc" in what:
        from gen_golden_vae import gen_unipc
        gen_unipc()


if __name__ == "__main__":
    main()
