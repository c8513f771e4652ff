"""libvcof against the reference's OWN CUDA path on the same GPU.

The unmodified reference files (staged under the git-ignored baseline/_ref by tools/stage_reference.py, loaded by
tools/gpu_reference.py) run as the CLIs run them — bf16 weights, autocast(bf16), flash-attn 2, cuBLAS — and are compared
with the libvcof block / model holding the SAME weights on the SAME inputs.  Three-way, with an fp32 "truth" computed
on the GPU by the reference code itself (fp32 weights, no autocast, SDPA):

    err(x)    = || x - truth ||_F / || truth - input ||_F         (error relative to the block's / model's update)
    assertion: err(libvcof) <= 1.25 * err(reference CUDA path) + 1e-3    "as accurate as the reference's own CUDA path"
               || libvcof - reference CUDA || / || update || < 1.5e-2   two independent bf16 evaluations agree
    and the oracle's `emulate_bf16=True` rounding points (asserted against at 1.5e-2 by tests/test_dit_gpu.py) are
    pinned to the executed CUDA reference: || oracle_emu - reference CUDA || / || update || < 1e-2.

Skipped (not failed) when baseline/_ref is absent — it is git-ignored and exists wherever __graft_entry__.build() ran
with /root/reference mounted; it travels to the GPU box with the snapshot.  Measured values are appended to
gpurun_out/gpu_reference_parity.jsonl so a run leaves its numbers behind.
"""
import json
import os

import pytest
import torch

import gpu_reference as gr
from oracle.dit_oracle import DiTConfig, block_forward, rope_table, temporal_positions

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not gr.available(), reason="baseline/_ref not staged (tools/stage_reference.py)")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C2 = dict(dim=5120, ffn_dim=13824, num_heads=40, num_layers=40)
C1 = dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30)


def _log(**kw):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_reference_parity.jsonl"), "a") as fh:
        fh.write(json.dumps(kw) + "\n")


def _our_block(ref_blk, cfg):
    from videocof_b200.dit import WanAttentionBlock
    ours = WanAttentionBlock("t2v_cross_attn", cfg["dim"], cfg["ffn_dim"], cfg["num_heads"], (-1, -1), True, True, 1e-6)
    ours.load_state_dict(ref_blk.state_dict(), strict=True)
    return ours.to("cuda", torch.bfloat16).eval()


def _truth_block(ref_blk, inp, freqs, fs, ground):
    """The reference block evaluated in fp32 by the reference code itself: fp32 copies of the (bf16-valued) weights,
    no autocast, SDPA in fp32."""
    import copy
    blk32 = copy.deepcopy(ref_blk).float()
    was = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad(), gr.backend("SDPA"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                y = blk32(inp["x"].float(), e=inp["e"], seq_lens=inp["seq_lens"], grid_sizes=inp["grid_sizes"],
                          freqs=freqs, context=inp["context"].float(), context_lens=None, dtype=torch.float32, t=0,
                          frame_split_indices=[fs], ground_frame_indices=[ground])
    finally:
        torch.backends.cuda.matmul.allow_tf32 = was
    del blk32
    return y


def _three_way(cfg, f, h, w, fs, with_oracle):
    dev = torch.device("cuda")
    ref_blk = gr.make_block(cfg, dev, seed=3)
    inp = gr.block_inputs(cfg, f, h, w, dev, seed=5)
    freqs = gr.rope_freqs(cfg["dim"] // cfg["num_heads"], dev)
    ground = (fs, fs + 1)
    y_ref = gr.run_block(ref_blk, inp, freqs, fs, ground).float()
    y_true = _truth_block(ref_blk, inp, freqs, fs, ground)
    ours = _our_block(ref_blk, cfg)
    with torch.no_grad():
        y = ours(inp["x"], inp["e"], inp["seq_lens"], inp["grid_sizes"], freqs, inp["context"],
                 frame_split_indices=[fs], ground_frame_indices=[ground]).float()
    upd = (y_true - inp["x"]).norm()
    res = dict(tokens=f * h * w, dim=cfg["dim"],
               err_ours=float((y - y_true).norm() / upd), err_ref=float((y_ref - y_true).norm() / upd),
               ours_vs_ref=float((y - y_ref).norm() / upd))
    if with_oracle:
        ocfg = DiTConfig(**cfg)
        p = {"blocks.0." + k: v.float().cpu() for k, v in ref_blk.state_dict().items()}
        x0 = inp["x"][0].float().cpu()
        args = (p, 0, x0, inp["e"][0].float().cpu(), inp["context"][0].float().cpu(), ocfg, (f, h, w),
                rope_table(ocfg.head_dim), temporal_positions(f, fs, ground), f * h * w)
        with torch.no_grad():
            o_emu = block_forward(*args, emu=True)
        res["oracle_emu_vs_ref"] = float((o_emu - y_ref[0].cpu()).norm() / upd.cpu())
        res["oracle_emu_vs_ours"] = float((o_emu - y[0].cpu()).norm() / upd.cpu())
    _log(test="block_three_way", **res)
    return res


@pytest.mark.parametrize("cfg,f,h,w,fs", [(C1, 5, 16, 16, 2), (C2, 3, 8, 10, 1)], ids=["c1_width_L1280", "c2_width_L240"])
def test_block_as_accurate_as_reference_cuda_path(cfg, f, h, w, fs):
    r = _three_way(cfg, f, h, w, fs, with_oracle=True)
    assert r["err_ours"] <= 1.25 * r["err_ref"] + 1e-3, r
    assert r["ours_vs_ref"] < 1.5e-2, r
    assert r["oracle_emu_vs_ref"] < 1e-2, r          # pins the oracle's bf16 rounding points to executed CUDA code


def test_c2_block_full_size_vs_reference_cuda_path():
    """One 14B block at the full C2 token count (75 600 tokens, chain of frames 10|1|10): the reference's FA2 / cuBLAS
    evaluation, its fp32 evaluation, and libvcof."""
    r = _three_way(C2, 21, 45, 80, 10, with_oracle=False)
    assert r["err_ours"] <= 1.25 * r["err_ref"] + 1e-3, r
    assert r["ours_vs_ref"] < 1.5e-2, r
