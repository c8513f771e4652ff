"""Temporally-sharded VAE decode vs the single-GPU decode (run under torchrun on N GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
        tools/vae_shard_check.py
Each rank decodes the full clip locally (reference result) and then its frame range with 2-frame halo exchange;
the gathered video must match bit for bit (same kernels, same per-position math)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from videocof_b200.vae import AutoencoderKLWan  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    vae = AutoencoderKLWan()
    for p in vae.parameters():
        if p.dim() > 1:
            torch.nn.init.normal_(p, std=1.0 / (p[0].numel() ** 0.5))
    vae = vae.to(dev, torch.bfloat16).eval()
    res = {}
    cases = [("small", (1, 16, max(5, 2 * world + 1), 6, 10)), ("720p", (1, 16, max(6, 21 if world >= 4 else 6), 90, 160))]
    if world >= 3:
        # fewer than two latent frames per rank: only min(P, f // 2) ranks hold frames, the others idle and join the gather
        cases.insert(1, ("idle_ranks", (1, 16, 5, 6, 10)))
        cases.append(("720p_idle_ranks", (1, 16, world + 1, 90, 160)))
    for name, shape in cases:
        g = torch.Generator().manual_seed(1)
        z = torch.randn(*shape, generator=g).bfloat16().to(dev)
        with torch.no_grad():
            ref = vae.decode(z).sample
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ref = vae.decode(z).sample
            torch.cuda.synchronize()
            t_single = time.perf_counter() - t0
            vae.enable_temporal_sharding()
            out = vae.decode(z).sample
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            out = vae.decode(z).sample
            torch.cuda.synchronize()
            t_shard = time.perf_counter() - t0
            vae._shard = None
            # encoder: video of the decoded length (1 + 4k frames)
            vid = ref.clamp(-1, 1)
            vae._shard = None
            mu_ref = vae.encode(vid)[0].mode()
            vae.enable_temporal_sharding()
            mu = vae.encode(vid)[0].mode()
            torch.cuda.synchronize()
            vae._shard = None
        enc_diff = float((mu.float() - mu_ref.float()).abs().max())
        diff = max(float((out.float() - ref.float()).abs().max()), enc_diff)
        d = torch.tensor([diff], device=dev)
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
        res[name] = dict(shape=list(out.shape), latent=list(mu.shape), max_abs_diff=float(d),
                         decode_single_ms=t_single * 1e3, decode_sharded_ms=t_shard * 1e3)
    if rank == 0:
        ok = all(v["max_abs_diff"] == 0.0 for v in res.values())
        print(json.dumps({"vae_shard_check": "ok" if ok else "FAIL", "world": world, **res}))
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
