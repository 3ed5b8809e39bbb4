"""The one exchange of a multi-GPU sweep: an all-gather of every rank's per-set best
(value f64, index i64, nan count i32: 24 bytes per set) over NCCL / NVLink, followed on every rank by the same
deterministic reduction (csrc/sweep.cu combine_kernel: larger value wins, ties to the lower grid index, then across
sets the first set attaining the maximum, NaN = -inf).  No other data crosses GPUs: candidates are independent given
the per-set state, which each rank rebuilds for the sets it touches."""
from __future__ import annotations

import torch


def gather_set_bests(local_table: torch.Tensor, gathered: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """local_table: uint8 tensor of S*24 bytes (this rank's global-indexed table, -inf/-1 where it holds nothing);
    gathered: uint8 tensor of world*S*24 bytes, filled rank-major.  Works with the nccl backend on device tensors and
    with gloo on CPU tensors (the world_size-2 CPU tests)."""
    import torch.distributed as dist
    if world == 1:
        gathered[:local_table.numel()].copy_(local_table)
        return gathered
    chunks = list(gathered.view(world, -1).unbind(0))
    dist.all_gather(chunks, local_table, group=group)
    return gathered
