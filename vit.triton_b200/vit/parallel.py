"""Batch-sharded data parallelism: one process per GPU, replicated weights, ONE collective per step.

The reference is single-GPU (no distributed code at all).  Every image is independent
through the whole forward (vit/vit.py:240-247), so the batch shards with no data-path exchange; the
only collective is the all-gather of the pooled CLS embeddings ((B/G, D) per rank, a few hundred KB:
latency-bound over NVLink/NVSwitch).  On GPUs with peer access that step is ONE kernel of ours
(``PeerGather``: pool + P2P stores into every peer's gather buffer + flag exchange, no NCCL call);
``all_gather_rows`` over torch.distributed is the portable form (NCCL for ragged shards or when
symmetric memory is unavailable; the CPU tests drive the same wrapper over gloo).
"""
import ctypes
import os
import warnings
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


class PeerGather:
    """Pool + all-gather as one kernel over NVLink peer memory (``vt_pool_cls_allgather``).

    Owns two symmetric gather buffers of (world * B_local, D) (alternated by the parity of the step,
    see the protocol note in csrc/rowwise.cu) and one symmetric array of ``world`` uint32 flag
    counters, all allocated and exchanged through ``torch.distributed._symmetric_memory``.
    ``__call__(hidden)`` returns the (world * B_local, D) gathered CLS rows: a VIEW of the current
    buffer, valid until the next-but-one call (copy it to keep it longer).  Not CUDA-graph capturable
    (the epoch is a launch argument)."""

    def __init__(self, batch_local: int, dim: int, dtype: torch.dtype, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from .kernels import _lib
        self._lib = _lib
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        assert self.world <= 16, f"At most 16 peers, provided: {self.world}"
        self.batch_local, self.dim, self.dtype = batch_local, dim, dtype
        self.bufs = symm_mem.empty((2, self.world * batch_local, dim), dtype=dtype, device=device)
        self.flags = symm_mem.empty((self.world,), dtype=torch.int32, device=device)
        self.flags.zero_()
        self._buf_handle = symm_mem.rendezvous(self.bufs, self.group)
        self._flag_handle = symm_mem.rendezvous(self.flags, self.group)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)                       # every rank's flags are zero before anyone writes
        buf_bytes = self.bufs[0].numel() * self.bufs.element_size()
        base = [int(a) for a in self._buf_handle.buffer_ptrs]
        self._out_tables = [(ctypes.c_void_p * self.world)(*[a + parity * buf_bytes for a in base])
                            for parity in (0, 1)]
        self._flag_table = (ctypes.c_void_p * self.world)(*[int(a) for a in self._flag_handle.buffer_ptrs])
        self.epoch = 0

    def __call__(self, hidden: torch.Tensor) -> torch.Tensor:
        assert hidden.is_cuda and hidden.dim() == 3, f"Expected CUDA (B, N, D) hidden states, provided: {tuple(hidden.shape)}"
        assert hidden.shape[0] == self.batch_local and hidden.shape[2] == self.dim and hidden.dtype == self.dtype, \
            f"PeerGather was built for ({self.batch_local}, *, {self.dim}) {self.dtype}, provided: {tuple(hidden.shape)} {hidden.dtype}"
        assert hidden.stride(2) == 1 and hidden.stride(0) % 8 == 0
        self.epoch += 1
        parity = self.epoch & 1
        _lib = self._lib
        _lib.call("vt_pool_cls_allgather", hidden.data_ptr(), self.batch_local, self.dim, hidden.stride(0),
                  _lib.dtype_code(hidden), ctypes.cast(self._out_tables[parity], ctypes.c_void_p),
                  ctypes.cast(self._flag_table, ctypes.c_void_p), self.rank, self.world, self.epoch,
                  _lib.stream_ptr(hidden))
        return self.bufs[parity]


class DataParallelVIT(torch.nn.Module):
    """Wraps a replicated ``VIT``: each rank runs its slice of the global batch and all ranks receive
    the gathered (B, D) pooled embeddings."""

    def __init__(self, model: torch.nn.Module, group=None, peer_gather: Optional[bool] = None,
                 output: str = "pooled"):
        """``output``: "pooled" gathers the CLS embeddings (B, D); "logits" gathers the class logits
        (B, num_labels) of a model built with a classifier head (``VIT(num_labels=...)``).
        ``peer_gather``: None = use the fused pool + peer-store kernel whenever it applies (CUDA,
        NCCL group, equal shards, symmetric memory available; VT_PEER_GATHER=0 disables), True =
        require it, False = always all-gather through torch.distributed."""
        super().__init__()
        self.model = model
        self.group = group
        self.peer_gather = peer_gather
        assert output in ("pooled", "logits"), f"output must be 'pooled' or 'logits', provided: {output}"
        self.output = output
        self._peer = None          # PeerGather, built on first use for one (B_local, D, dtype)
        self._peer_failed = False
        self.gather_impl = "none"  # what the last forward used: "peer-store kernel" | "torch.distributed" | "none"

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    @property
    def rank(self) -> int:
        return dist.get_rank(self.group) if dist.is_initialized() else 0

    def local_slice(self, global_batch: int) -> Tuple[int, int]:
        return shard_bounds(global_batch, self.world_size, self.rank)

    def forward_local(self, x_local: torch.Tensor) -> torch.Tensor:
        """Pooled embeddings (or logits) of this rank's images, (B_local, D | num_labels); no communication."""
        return self.model.logits(x_local) if self.output == "logits" else self.model.pooled(x_local)

    def forward(self, x_local: torch.Tensor, global_batch: Optional[int] = None) -> torch.Tensor:
        """x_local: this rank's shard of the global batch (as laid out by ``local_slice``).
        Returns the pooled embeddings of the WHOLE batch, (B, D), identical on every rank."""
        world = self.world_size
        if world == 1:
            return self.forward_local(x_local)
        if global_batch is None:
            global_batch = x_local.shape[0] * world
        if self._wants_peer(x_local, global_batch):
            if self.output == "logits":
                hidden = self.model.logits(x_local).unsqueeze(1)      # (B_local, 1, num_labels): "CLS row" = the logits
            else:
                hidden = self.model.forward_uint8(x_local) if x_local.dtype == torch.uint8 else self.model(x_local)
            peer = self._peer_for(hidden)
            if peer is not None:
                self.gather_impl = "peer-store kernel"
                return peer(hidden.contiguous())
            out = torch.empty((hidden.shape[0], hidden.shape[2]), device=hidden.device, dtype=hidden.dtype)
            from .kernels import _lib
            _lib.call("vt_pool_cls", hidden.data_ptr(), out.data_ptr(), hidden.shape[0], hidden.shape[2],
                      hidden.stride(0), _lib.dtype_code(hidden), _lib.stream_ptr(hidden))
            self.gather_impl = "torch.distributed"
            return all_gather_rows(out, global_batch, self.group)
        self.gather_impl = "torch.distributed"
        return all_gather_rows(self.forward_local(x_local), global_batch, self.group)

    def _wants_peer(self, x_local: torch.Tensor, global_batch: int) -> bool:
        if self.peer_gather is False or self._peer_failed:
            return False
        if self.peer_gather is None and os.environ.get("VT_PEER_GATHER", "1") == "0":
            return False
        ok = x_local.is_cuda and global_batch == x_local.shape[0] * self.world_size and x_local.shape[0] > 0 \
            and dist.get_backend(self.group) == "nccl"
        if ok and self.output == "logits":
            ok = self.model.classifier.weight.shape[1] % 8 == 0      # 16-byte rows for the vector stores
        if self.peer_gather is True:
            assert ok, "peer_gather=True needs CUDA inputs, an NCCL group and equal shards"
        return ok

    def _peer_for(self, hidden: torch.Tensor) -> Optional[PeerGather]:
        key = (hidden.shape[0], hidden.shape[2], hidden.dtype)
        if self._peer is not None and (self._peer.batch_local, self._peer.dim, self._peer.dtype) == key:
            return self._peer
        # building the buffers is a collective: every rank gets here with the same key or none does
        try:
            self._peer = PeerGather(hidden.shape[0], hidden.shape[2], hidden.dtype, hidden.device, self.group)
        except Exception as exc:   # symmetric memory unavailable (no P2P, old driver): use NCCL from now on
            if self.peer_gather is True:
                raise
            warnings.warn(f"peer-memory gather unavailable ({exc!r}); falling back to NCCL all-gather")
            self._peer, self._peer_failed = None, True
        return self._peer
