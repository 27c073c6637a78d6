"""Multi-GPU plumbing: environments are independent, so they shard by contiguous env ranges with
no traffic on the step path; the only collective is one all-gather of the fixed-size rollout
statistics vector at the end of a run."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_shard(total_envs: int, rank: int, world: int):
    """Contiguous [begin, end) env range of ``rank`` (first ``total % world`` ranks get one extra)."""
    base, rem = divmod(total_envs, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env; returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def gather_stats(stats: torch.Tensor):
    """all_gather of a per-rank stats vector -> [world, len]; identity for world == 1."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return stats.unsqueeze(0)
    out = [torch.empty_like(stats) for _ in range(dist.get_world_size())]
    dist.all_gather(out, stats.contiguous())
    return torch.stack(out, dim=0)


def reduce_stats(all_stats: torch.Tensor):
    """Combine gathered rollout stats rows (layout include/mds_b200.h MDS_STAT_*)."""
    s = all_stats.double()
    out = s.sum(0)
    out[2] = s[:, 2].max()   # max position error
    out[3] = s[:, 3].min()   # min barrier value
    return out
