"""Batch sharding of the DyCON losses over ranks (one process per GPU).

The path partitions along the batch dimension: UnCL is per voxel, the FeCL
student term is per sample, and the only batch-global quantities are the two mean denominators
(B*V voxels, B*N rows) and the hard-negative count ``cnt`` of the teacher term
(code/utils/dycon_losses.py:116,193,229).  Each rank therefore runs the kernels on its own samples
with the GLOBAL denominators baked in (``inv_count``, ``inv_rows``), all-reduces its partial sums
-- 1 double for UnCL, 3 doubles {student_sum, cross_sum, cnt} for FeCL -- and evaluates the loss
from the reduced sums.  The backward needs no communication: the FeCL kernel reads the reduced
``cnt`` from device memory and the local gradients are already scaled for a SUM over ranks (the
loss is the global mean), which is what DDP's gradient all-reduce of the network parameters needs
once its default averaging is undone (multiply by world size) or ``global_batch`` is left unset.

This module holds only host logic (no kernels) so the N>1 protocol can be tested on CPU with gloo.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

FECL_TINY = 1e-18      # dycon_losses.py:229


def all_reduce_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of a small vector of partial sums (no-op without a process group)."""
    if group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def uncl_loss_from_sum(total: torch.Tensor, inv_count: float) -> torch.Tensor:
    """total = sum over ALL ranks of sum_v L_v; inv_count = 1/(B_global * V)."""
    return (total.reshape(-1)[0] * inv_count).to(torch.float32)


def fecl_loss_from_sums(sums: torch.Tensor, inv_rows: float, lambda_cross: float, has_teacher: bool) -> torch.Tensor:
    """sums = reduced {student_sum, cross_sum, cross_cnt}; inv_rows = 1/(B_global * N)."""
    loss = sums[0] * inv_rows
    if has_teacher:
        loss = loss + lambda_cross * (sums[1] / (sums[2] + FECL_TINY))
    return loss.to(torch.float32)


def shard_bounds(global_batch: int, rank: int, world_size: int):
    """Contiguous sample range of `rank`; requires world_size <= global_batch (else: replicas only)."""
    if world_size > global_batch:
        raise ValueError(f"cannot shard a batch of {global_batch} samples over {world_size} ranks: run replicas instead")
    base, extra = divmod(global_batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
