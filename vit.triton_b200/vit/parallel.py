"""Batch-sharded data parallelism: one process per GPU, replicated weights, ONE collective per step.

The reference is single-GPU (no distributed code at all).  Every image is independent
through the whole forward (vit/vit.py:240-247), so the batch shards with no data-path exchange; the
only collective is the all-gather of the pooled CLS embeddings ((B/G, D) per rank, a few hundred KB:
latency-bound over NVLink/NVSwitch).  Works with any torch.distributed backend (NCCL on GPUs; the
CPU tests drive the same code over gloo).
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of ``total`` items owned by ``rank``; the first ``total % world``
    ranks hold one extra item."""
    assert world_size > 0 and 0 <= rank < world_size, f"Invalid rank/world size: {rank}, {world_size}"
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_gather_rows(local: torch.Tensor, total_rows: int, group=None) -> torch.Tensor:
    """All-gather row shards laid out by ``shard_bounds`` into the full (total_rows, ...) tensor.
    Equal shards take the single-buffer path; ragged shards are padded to the largest one."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(total_rows, world, rank)
    assert local.shape[0] == hi - lo, f"Rank {rank} should hold {hi - lo} rows, provided: {local.shape[0]}"
    local = local.contiguous()
    tail = tuple(local.shape[1:])
    if total_rows % world == 0:
        out = torch.empty((total_rows,) + tail, device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    widest = (total_rows + world - 1) // world
    padded = torch.zeros((widest,) + tail, device=local.device, dtype=local.dtype)
    padded[:local.shape[0]] = local
    pieces = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(pieces, padded, group=group)
    rows = []
    for r, piece in enumerate(pieces):
        a, b = shard_bounds(total_rows, world, r)
        rows.append(piece[:b - a])
    return torch.cat(rows, dim=0)


class DataParallelVIT(torch.nn.Module):
    """Wraps a replicated ``VIT``: each rank runs its slice of the global batch and all ranks receive
    the gathered (B, D) pooled embeddings."""

    def __init__(self, model: torch.nn.Module, group=None):
        super().__init__()
        self.model = model
        self.group = group

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    @property
    def rank(self) -> int:
        return dist.get_rank(self.group) if dist.is_initialized() else 0

    def local_slice(self, global_batch: int) -> Tuple[int, int]:
        return shard_bounds(global_batch, self.world_size, self.rank)

    def forward_local(self, x_local: torch.Tensor) -> torch.Tensor:
        """Pooled embeddings of this rank's images, (B_local, D); no communication."""
        return self.model.pooled(x_local)

    def forward(self, x_local: torch.Tensor, global_batch: Optional[int] = None) -> torch.Tensor:
        """x_local: this rank's shard of the global batch (as laid out by ``local_slice``).
        Returns the pooled embeddings of the WHOLE batch, (B, D), identical on every rank."""
        pooled = self.forward_local(x_local)
        if self.world_size == 1:
            return pooled
        if global_batch is None:
            global_batch = x_local.shape[0] * self.world_size
        return all_gather_rows(pooled, global_batch, self.group)
