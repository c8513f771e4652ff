"""Sequence-parallel DiT forward vs the single-GPU forward (run under torchrun on N GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/sp_check.py [push] [c2 [layers=N]]

Every rank builds the same random-init model (same seed), runs the un-sharded forward locally and the
token-sharded forward across ranks — once with the head exchange (all-to-all) and once with the K/V all-gather
(videocof_b200/dist.py) — and both outputs must agree with the single-GPU one to bf16 round-off (identical
per-token math; only the attention tile/accumulation order can differ)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from videocof_b200.dit import WanTransformer3DModel  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(9)
    if "c2" in sys.argv[1:]:
        # BASELINE config C2 widths and token count (14B: dim 5120, 40 heads; 21 x 45 x 80 = 75 600 tokens, chain of frames
        # 10|1|10) with `layers` blocks (default 1: the exchange logic is per layer; `c2 layers=4` for more depth)
        layers = next((int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("layers=")), 1)
        cfg = dict(dim=5120, ffn_dim=13824, num_heads=40, num_layers=layers)
        x = torch.randn(1, 16, 21, 90, 160, generator=g).bfloat16().to(dev)
        ctx = [torch.randn(77, 4096, generator=g).bfloat16().to(dev)]
        kw = dict(seq_len=75600, frame_split_indices=[10], ground_frame_indices=[(10, 11)])
    else:
        cfg = dict(dim=1024, ffn_dim=2048, num_heads=8, num_layers=3, text_dim=128, text_len=64)   # 8 heads: P | 8
        # 5 x 7 x 9 = 315 tokens: not divisible by 2/4/8 -> exercises the padding rule
        x = torch.randn(1, 16, 5, 14, 18, generator=g).bfloat16().to(dev)
        ctx = [torch.randn(11, 128, generator=g).bfloat16().to(dev)]
        kw = dict(seq_len=315, frame_split_indices=[2], ground_frame_indices=[(2, 3)])
    model = WanTransformer3DModel.random_init(device=dev, seed=3, **cfg)
    t = torch.tensor([749.0], device=dev)
    with torch.no_grad():
        ref = model(x=x, t=t, context=ctx, **kw)
        model.enable_multi_gpus_inference()
        worst = 0.0
        report = {}
        # `sp_check.py push` adds the push exchange (symmetric memory, written without hardware at the end of round 1)
        modes = ("heads", "gather") + (("push",) if "push" in sys.argv[1:] else ())
        for mode in modes:
            os.environ["VCOF_SP_MODE"] = mode
            out = model(x=x, t=t, context=ctx, **kw)
            if mode == "push":
                out = model(x=x, t=t, context=ctx, **kw)      # second pass: the receive buffers are reused
            torch.cuda.synchronize()
            rel = float((out.float() - ref.float()).norm() / ref.float().norm())
            res = torch.tensor([rel], device=dev)
            dist.all_reduce(res, op=dist.ReduceOp.MAX)
            report[mode] = {"rel_fro_max_over_ranks": float(res),
                            "max_abs_rank0": float((out.float() - ref.float()).abs().max()),
                            "bit_identical_rank0": bool(torch.equal(out, ref))}
            worst = max(worst, float(res))
        os.environ.pop("VCOF_SP_MODE", None)
    if rank == 0:
        print(json.dumps({"sp_check": "ok" if worst < 5e-3 else "FAIL", "world": world, "dim": cfg["dim"],
                          "tokens": kw["seq_len"], "layers": cfg["num_layers"], **report}))
    dist.destroy_process_group()
    return 0 if worst < 5e-3 else 1


if __name__ == "__main__":
    sys.exit(main())
