"""Data-parallel sharding of clips over ranks and the single collective of the hot path.

Clips are independent: each rank takes a contiguous slice of the batch, weights are replicated, and the only
exchange is one all-gather of the generated code sequences (B/G, T-1) int64.  The one cross-sample effect in the
reference is the batch-index positional-encoding quirk of the VQ-VAE decoder (models/lib/base_models.py:271-273, SURVEY
F4): sample b receives pe[b], so every shard carries its GLOBAL batch positions.

Backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.  One process per GPU; rendezvous from the torchrun env.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK/WORLD_SIZE/MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_range(total: int, rank: int, world: int):
    """Contiguous [start, end) of rank's slice; the first total % world ranks get one extra item."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(batch: dict, rank: int, world: int):
    """Slice every (B, ...) tensor of `batch`; adds 'batch_index' (int32 global positions) for the F4 quirk."""
    B = next(v for v in batch.values() if torch.is_tensor(v)).shape[0]
    s, e = shard_range(B, rank, world)
    out = {k: (v[s:e] if torch.is_tensor(v) and v.shape[:1] == (B,) else v) for k, v in batch.items()}
    out["batch_index"] = torch.arange(s, e, dtype=torch.int32)
    return out


def all_gather_codes(codes_local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather per-rank code sequences (b_r, S) into (total, S) on every rank, in global batch order.
    Uneven shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        assert codes_local.shape[0] == total
        return codes_local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    S = codes_local.shape[1]
    most = -(-total // world)
    s, e = shard_range(total, rank, world)
    assert codes_local.shape[0] == e - s, "shard does not match shard_range"
    send = codes_local
    if e - s < most:
        send = torch.cat([codes_local, codes_local.new_full((most - (e - s), S), -1)], 0)
    recv = torch.empty(world * most, S, dtype=codes_local.dtype, device=codes_local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    parts = []
    for r in range(world):
        rs, re = shard_range(total, r, world)
        parts.append(recv[r * most: r * most + (re - rs)])
    return torch.cat(parts, 0)
