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
