"""videox_fun.dist — the reference routes sequence parallelism through xfuser (dist/fuser.py, wan_xfuser.py),
which cannot run VideoCoF's chain-of-frames kwargs.  Here the same entry points sit on
videocof_b200.dist (token-sharded DiT with a per-layer K/V all-gather over NCCL)."""
import torch.distributed as dist

from videocof_b200.dist import SequenceParallel  # noqa: F401


def get_sequence_parallel_world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def get_sequence_parallel_rank():
    return dist.get_rank() if dist.is_initialized() else 0


def set_multi_gpus_devices(ulysses_degree=1, ring_degree=1, classifier_free_guidance_degree=1):
    """reference dist/fuser.py:35-55 — returns the CUDA device of this rank (process group must be initialised)."""
    import torch
    if not dist.is_initialized():
        dist.init_process_group("nccl")
    local = dist.get_rank() % max(torch.cuda.device_count(), 1)
    torch.cuda.set_device(local)
    return torch.device("cuda", local)
