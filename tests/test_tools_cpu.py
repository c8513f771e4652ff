"""The measurement tools stay runnable: tools/torch_block_bench.py (plain-PyTorch library baseline of one DiT block) on
tiny shapes through its CPU / SDPA branch, checked against the oracle block it restates."""
import json

import torch

import torch_block_bench as tb
from oracle import dit_oracle


def test_torch_block_bench_runs_and_prints_one_json_line(capsys):
    assert tb.main(["--tokens", "192", "--dim", "256", "--ffn", "512", "--heads", "2", "--layers", "3", "--warmup", "1",
                    "--iters", "1", "--device", "cpu"]) == 0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["finite"] and line["ms_per_block"] > 0 and line["tokens"] == 192
    assert abs(line["est_ms_per_step"] - 3 * line["ms_per_block"]) < 1e-6


def test_flop_count_matches_the_survey_figure():
    # 1.631e14 FLOP per c2 block, 1.256e11 per c1 block
    assert abs(tb.block_flops(75600, 5120, 13824) / 1.631e14 - 1) < 2e-3
    assert abs(tb.block_flops(1280, 1536, 8960) / 1.256e11 - 1) < 2e-3


def test_block_restates_the_oracle_block():
    """Same weights into the oracle's block_forward (fp32, no rotation: angles zero) and into the tool's block: the
    bf16 library path must agree with the fp32 oracle to bf16 accuracy, so the baseline times the right arithmetic."""
    cfg = dit_oracle.DiTConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=1, text_dim=64, text_len=32)
    p = dit_oracle.make_block_params(cfg, 0, seed=3)
    g = torch.Generator().manual_seed(1)
    L, C = 96, 256
    blk = tb.Block(C, 512, 2, torch.device("cpu"), g)
    pre = "blocks.0."
    names = {"sq": "self_attn.q", "sk": "self_attn.k", "sv": "self_attn.v", "so": "self_attn.o",
             "cq": "cross_attn.q", "ck": "cross_attn.k", "cv": "cross_attn.v", "co": "cross_attn.o"}
    for k, n in names.items():
        blk.w[k], blk.b[k] = p[pre + n + ".weight"].bfloat16(), p[pre + n + ".bias"].bfloat16()
    blk.w1, blk.b1 = p[pre + "ffn.0.weight"].bfloat16(), p[pre + "ffn.0.bias"].bfloat16()
    blk.w2, blk.b2 = p[pre + "ffn.2.weight"].bfloat16(), p[pre + "ffn.2.bias"].bfloat16()
    blk.nq, blk.nk = p[pre + "self_attn.norm_q.weight"].bfloat16(), p[pre + "self_attn.norm_k.weight"].bfloat16()
    blk.cnq, blk.cnk = p[pre + "cross_attn.norm_q.weight"].bfloat16(), p[pre + "cross_attn.norm_k.weight"].bfloat16()
    blk.n3w, blk.n3b = p[pre + "norm3.weight"].float(), p[pre + "norm3.bias"].float()
    blk.modulation = p[pre + "modulation"].float().view(6, C)
    x = torch.randn(L, C, generator=g)
    e0 = torch.randn(6, C, generator=g) * 0.1
    ctx = torch.randn(32, C, generator=g).bfloat16()
    freqs = torch.polar(torch.ones(L, 64, dtype=torch.float64), torch.zeros(L, 64, dtype=torch.float64))
    with torch.no_grad():
        got = blk.forward(x, e0, ctx, freqs)
        angles = torch.zeros(1024, 64, dtype=torch.float64)
        ref = dit_oracle.block_forward(p, 0, x, e0, ctx.float(), cfg, (1, 1, L), angles, [0], L)
    rel = float((got - ref).norm() / ref.norm())
    assert rel < 2e-2, rel


# ---- tools/gpu_reference.py: the reference's own CUDA path (the GPU reference arm) — plumbing only on a CPU box ----------
def _ref_present():
    import gpu_reference as gr
    return gr.available()


def test_stage_reference_is_idempotent_and_byte_identical(tmp_path):
    """baseline/_ref holds the unmodified package: every staged file's sha256 equals the manifest's, and the
    manifest's equals the mounted reference's where that exists (the build container)."""
    import hashlib
    import os

    import pytest

    import stage_reference as st
    if not st.stage(verbose=False):
        pytest.skip("neither /root/reference nor baseline/_ref present")
    man = json.load(open(os.path.join(st.DST, "MANIFEST.json")))
    assert "videox_fun/models/wan_transformer3d.py" in man["files"] and len(man["files"]) >= 40
    for rel_path, digest in man["files"].items():
        assert hashlib.sha256(open(os.path.join(st.DST, rel_path), "rb").read()).hexdigest() == digest, rel_path
        src = os.path.join(st.SRC, rel_path)
        if os.path.exists(src):
            assert hashlib.sha256(open(src, "rb").read()).hexdigest() == digest, rel_path


def test_gpu_reference_block_runs_the_reference_code_and_matches_the_oracle():
    """The reference's WanAttentionBlock loaded under a private package name (the repo's own `videox_fun` overlay stays
    importable) evaluates to the oracle's block on the same weights: the GPU arm times the right arithmetic."""
    import pytest
    if not _ref_present():
        pytest.skip("reference not reachable")
    import gpu_reference as gr
    cfg = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=1)
    blk = gr.make_block(cfg, "cpu", seed=4)
    f, h, w, fs = 3, 4, 6, 1
    inp = gr.block_inputs(cfg, f, h, w, "cpu", seed=6, text_len=32)
    freqs = gr.rope_freqs(128, "cpu")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y = gr.run_block(blk, inp, freqs, fs, (fs, fs + 1), attention_type="SDPA")
    ocfg = dit_oracle.DiTConfig(text_dim=64, text_len=32, **cfg)
    p = {"blocks.0." + k: v.float() for k, v in blk.state_dict().items()}
    with torch.no_grad():
        ref = dit_oracle.block_forward(p, 0, inp["x"][0], inp["e"][0], inp["context"][0].float(), ocfg, (f, h, w),
                                       dit_oracle.rope_table(128), dit_oracle.temporal_positions(f, fs, (fs, fs + 1)),
                                       f * h * w)
    upd = (ref - inp["x"][0]).norm()
    assert float((y[0].float() - ref).norm() / upd) < 2e-2       # CPU autocast(bf16) on the Linears vs fp32
    import videox_fun
    assert "baseline" not in (videox_fun.__file__ or "") and "reference" not in (videox_fun.__file__ or "")
    assert abs(gr.block_flops(75600, 5120, 13824) / 1.631e14 - 1) < 2e-3
