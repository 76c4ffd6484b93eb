"""Multi-GPU plumbing: one process per GPU (torchrun), envs sharded in contiguous blocks, scene replicated,
NO collective in the render path.  The only optional exchange is a gather of the observation
tensors to rank 0 for a centralised policy.
"""
from __future__ import annotations

from typing import List, Optional, Tuple


def shard_envs(n_envs: int, world_size: int, rank: int) -> range:
    """Contiguous block of envs for `rank`: [r*E/G, (r+1)*E/G) with the remainder spread over the first ranks."""
    base, rem = divmod(n_envs, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_sizes(n_envs: int, world_size: int) -> List[int]:
    return [len(shard_envs(n_envs, world_size, r)) for r in range(world_size)]


def gather_observations(obs, dst: int = 0, group=None):
    """torch.distributed gather of one observation tensor [E_local, ...] to rank `dst` (NCCL on GPUs, gloo on
    CPU).  Shards may differ by one env; they are padded to the largest shard for the collective and trimmed
    on the destination.  Returns the concatenated [E, ...] tensor on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local = torch.tensor([obs.shape[0]], device=obs.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s) for s in sizes]
    m = max(sizes)
    padded = obs if obs.shape[0] == m else torch.cat([obs, obs.new_zeros((m - obs.shape[0],) + tuple(obs.shape[1:]))])
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded.contiguous(), bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)])


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; initialises the process group if world > 1."""
    import os

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"))
    return rank, world, local


def bind_to_gpu_numa(device: int) -> Optional[List[int]]:
    """Pin this process to the CPUs NVML reports as local to `device` (its NUMA node), before any pinned host memory
    is allocated: with one process per GPU the observation copies (16.5 MB per step and GPU in BASELINE configs[1])
    then stay on the socket the GPU hangs off instead of crossing the inter-socket link.  Returns the CPU list, or
    None when NVML / the affinity call is unavailable (nothing is changed then)."""
    import os

    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        uuid = "GPU-" + str(torch.cuda.get_device_properties(device).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
