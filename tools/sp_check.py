"""Sequence-parallel DiT forward vs the single-GPU forward (run under torchrun on N GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/sp_check.py

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
    cfg = dict(dim=1024, ffn_dim=2048, num_heads=8, num_layers=3, text_dim=128, text_len=64)   # 8 heads: P | 8
    model = WanTransformer3DModel.random_init(device=dev, seed=3, **cfg)
    g = torch.Generator().manual_seed(9)
    # 5 x 7 x 9 = 315 tokens: not divisible by 2/4/8 -> exercises the padding rule
    x = torch.randn(1, 16, 5, 14, 18, generator=g).bfloat16().to(dev)
    ctx = [torch.randn(11, 128, generator=g).bfloat16().to(dev)]
    t = torch.tensor([749.0], device=dev)
    kw = dict(seq_len=315, frame_split_indices=[2], ground_frame_indices=[(2, 3)])
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
                            "max_abs_rank0": float((out.float() - ref.float()).abs().max())}
            worst = max(worst, float(res))
        os.environ.pop("VCOF_SP_MODE", None)
    if rank == 0:
        print(json.dumps({"sp_check": "ok" if worst < 5e-3 else "FAIL", "world": world, **report}))
    dist.destroy_process_group()
    return 0 if worst < 5e-3 else 1


if __name__ == "__main__":
    sys.exit(main())
