"""Per-layer communication cost of the two sequence-parallel schemes at the C2 shape (run under torchrun):
all_to_all_single of one [L/P, C] projection (head exchange) vs all_gather_into_tensor of one [L/P, C] shard
(K/V all-gather), plus the pack / unpack copies, timed with CUDA events on an otherwise idle GPU."""
import json
import os
import sys

import torch
import torch.distributed as dist


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    rank, P = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L, C = 75600, 5120
    rows = (L + P - 1) // P
    cp = C // P
    x = torch.randn(rows, C, device=dev).bfloat16()
    send = torch.empty(P, rows, cp, dtype=torch.bfloat16, device=dev)
    recv = torch.empty(P * rows, cp, dtype=torch.bfloat16, device=dev)
    full = torch.empty(P * rows, C, dtype=torch.bfloat16, device=dev)
    out = torch.empty(rows, C, dtype=torch.bfloat16, device=dev)
    res = {"world": P, "rows": rows, "shard_MB": rows * C * 2 / 1e6}
    res["pack_ms"] = timed(lambda: send.copy_(x.view(rows, P, cp).transpose(0, 1)))
    res["a2a_ms"] = timed(lambda: dist.all_to_all_single(recv.view(-1), send.view(-1)))
    res["unpack_ms"] = timed(lambda: out.view(rows, P, cp).copy_(send.transpose(0, 1)))
    res["all_gather_ms"] = timed(lambda: dist.all_gather_into_tensor(full, x))
    res["a2a_GBps_per_rank"] = rows * C * 2 * (P - 1) / P / res["a2a_ms"] / 1e6
    res["all_gather_GBps_per_rank"] = rows * C * 2 * (P - 1) / res["all_gather_ms"] / 1e6
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
