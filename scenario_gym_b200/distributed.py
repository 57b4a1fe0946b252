"""
Multi-GPU plumbing.  Scenarios are independent (no cross-scenario term anywhere on the
path), so the batch is sharded by contiguous scenario blocks, one process per GPU, with
NO per-tick communication.  The only collective is the final gather of fixed-size
per-scenario metric records (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist

RECORD_FIELDS = (
    "ego_avg_speed", "ego_max_speed", "ego_dist", "first_coll_tick", "first_coll_a",
    "first_coll_b", "n_pair_ticks", "rss_flags", "tick", "t",
)


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of scenarios owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def pack_records(fields: dict) -> torch.Tensor:
    """[n, len(RECORD_FIELDS)] float64 record per scenario (all integers are exact in fp64)."""
    cols = []
    pair = fields["first_coll_pair"].reshape(-1, 2)
    src = dict(fields)
    src["first_coll_a"], src["first_coll_b"] = pair[:, 0], pair[:, 1]
    for k in RECORD_FIELDS:
        cols.append(src[k].to(torch.float64).reshape(-1))
    return torch.stack(cols, dim=1).contiguous()


def gather_records(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """
    All-gather the per-scenario records of every rank into scenario order.  Shards may
    differ by one row, so rows are padded to the largest shard for the collective.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    if all(hi - lo == width for lo, hi in sizes):  # equal shards: the gathered buffer is the answer
        out = torch.empty((world * width, local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    padded = torch.zeros((width, local.shape[1]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * width, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    out = out.view(world, width, -1)
    return torch.cat([out[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
