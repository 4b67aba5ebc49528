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
once its default averaging is undone (multiply by world size).  With a process group the denominators
are ALWAYS global: ``global_batch`` defaults to ``B_local * world_size`` (equal shards); pass it
explicitly for ragged shards.

The exchange itself runs over NVLink peer memory in the TAIL of the forward kernels (``dycon_uncl_fwd_sharded`` /
``dycon_fecl_fwd_sharded``, csrc/exchange.cuh): no extra launch, no NCCL call on the step's critical path.  Ranks may
lag behind each other (rank-0-only validation, a data-loader stall): the others wait inside their kernel, like in a
blocking collective, for at most ``DYCON_EXCHANGE_TIMEOUT_S`` seconds (default 600, 0 = for ever); a rank that gives
up gets NaN losses and a raised error word (``PeerExchange.timed_out()``) instead of a destroyed CUDA context.  Set
``DYCON_PEER_EXCHANGE=0`` to use the backend's all-reduce instead.

This module holds only host logic (no kernels) so the N>1 protocol can be tested on CPU with gloo.
"""
from __future__ import annotations

import os
import sys

import torch
import torch.distributed as dist

FECL_TINY = 1e-18      # dycon_losses.py:229


class PeerExchange:
    """All-reduce of a few doubles over NVLink peer memory (``dycon_exchange_sums``, csrc/exchange.cu).

    Every rank allocates an inbox, shares it with the ranks of the group through CUDA IPC and hands the
    kernel the table of peer pointers.  Construction is collective (all ranks of the group, same program
    point); afterwards a call is one tiny kernel on the caller's stream -- no NCCL, no stream hand-off,
    CUDA-graph replayable -- instead of the ~25 us a 32-byte NCCL all-reduce costs per call.
    """

    def __init__(self, group, device):
        import ctypes
        from torch.multiprocessing.reductions import reduce_tensor
        from . import _lib
        self.lib = _lib
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        nbytes = _lib.lib().dycon_exchange_inbox_bytes()
        self.inbox = torch.zeros(nbytes // 8, dtype=torch.float64, device=device)
        self.seq = torch.zeros(_lib.EXCHANGE_CHANNELS, dtype=torch.int64, device=device)      # one counter per channel
        self.error_offset = _lib.lib().dycon_exchange_error_offset() // 8
        self.timeout_s = -1.0          # negative: DYCON_EXCHANGE_TIMEOUT_S from the environment, default 600 s
        torch.cuda.synchronize(device)                 # zero-filled before any peer can learn the handle
        handles = [None] * self.world
        dist.all_gather_object(handles, reduce_tensor(self.inbox), group=group)
        self.peers = []
        devices = [None] * self.world
        dist.all_gather_object(devices, device.index, group=group)
        with torch.cuda.device(device):
            for r, (rebuild, args) in enumerate(handles):
                if r == self.rank:
                    self.peers.append(self.inbox)
                    continue
                _lib.check(_lib.lib().dycon_exchange_enable_peer(devices[r]), "dycon_exchange_enable_peer")
                # Open the IPC handle with THIS rank's device current (argument 6 of rebuild_cuda_tensor is the
                # device the handle is opened on): cudaIpcOpenMemHandle then maps the peer's memory for access
                # from here.  The tensor object claims the local device; only its data_ptr() is used.
                args = list(args)
                args[6] = device.index
                self.peers.append(rebuild(*args))
        self.table = (ctypes.c_void_p * self.world)(*[t.data_ptr() for t in self.peers])

    def all_reduce_(self, sums: torch.Tensor, kind=0, scale=0.0, lambda_cross=0.0, loss_out=None) -> torch.Tensor:
        """In-place all-reduce of `sums`; with `kind` != 0 the same launch also writes the loss to `loss_out`."""
        import ctypes
        stream = torch.cuda.current_stream(sums.device).cuda_stream
        ptr = ctypes.c_void_p(sums.data_ptr())
        lptr = ctypes.c_void_p(loss_out.data_ptr()) if loss_out is not None else None
        self.lib.check(self.lib.lib().dycon_exchange_sums(ptr, sums.numel(), ptr, self.table, self.rank, self.world,
                                                          ctypes.c_void_p(self.seq.data_ptr()), int(kind), float(scale),
                                                          float(lambda_cross), lptr, self.timeout_s, stream),
                       "dycon_exchange_sums")
        return sums

    def abi_args(self):
        """(peer_inboxes, rank, world, seq_counters, timeout_s): the trailing arguments of the *_sharded entry points."""
        import ctypes
        return (self.table, self.rank, self.world, ctypes.c_void_p(self.seq.data_ptr()), self.timeout_s)

    def timed_out(self) -> bool:
        """True if a wait of any channel has given up so far (host sync: diagnostics only)."""
        words = self.inbox[self.error_offset:self.error_offset + self.lib.EXCHANGE_CHANNELS].view(torch.int64)
        return bool((words != 0).any().item())


_exchanges = {}      # (id(group), device index) -> PeerExchange, or None when peer memory is unavailable


def _peer_exchange(group, device):
    key = (id(group), device.index)
    if key in _exchanges:
        return _exchanges[key]
    ex = None
    ok = torch.ones(1, dtype=torch.int32, device=device)
    if os.environ.get("DYCON_PEER_EXCHANGE", "1") == "0" or dist.get_world_size(group) > 16:
        ok.zero_()
    else:
        try:
            ex = PeerExchange(group, device)
        except Exception as err:      # e.g. no P2P between the devices: fall back to the backend's all-reduce
            print(f"dycon: peer-memory exchange unavailable ({type(err).__name__}: {err}); using all_reduce",
                  file=sys.stderr)
            ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)      # all ranks or none (this also orders the set-up)
    if int(ok.item()) == 0:
        ex = None
    _exchanges[key] = ex
    return ex


def all_reduce_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of a small vector of partial sums (no-op without a process group).
    CUDA float64 vectors of at most 7 entries go over NVLink peer memory (``PeerExchange``); anything else
    (CPU / gloo in the tests, or no P2P) uses the backend's all-reduce."""
    if group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if sums.is_cuda and sums.dtype == torch.float64 and sums.is_contiguous() and sums.numel() <= 7:
            ex = _peer_exchange(group, sums.device)
            if ex is not None:
                return ex.all_reduce_(sums)
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


def reduce_uncl(total: torch.Tensor, inv_count: float, group) -> torch.Tensor:
    """All-reduce the UnCL partial sum in place and return the global loss (0-dim fp32)."""
    ex = _exchange_for(total, group)
    if ex is not None:
        loss = torch.empty((), dtype=torch.float32, device=total.device)
        ex.all_reduce_(total, _lib_const("EXCHANGE_UNCL"), inv_count, 0.0, loss)
        return loss
    return uncl_loss_from_sum(all_reduce_sums(total, group), inv_count)


def reduce_fecl(sums: torch.Tensor, inv_rows: float, lambda_cross: float, has_teacher: bool, group) -> torch.Tensor:
    """All-reduce the FeCL partial sums {student, cross_sum, cross_cnt} in place and return the global loss."""
    ex = _exchange_for(sums, group)
    if ex is not None:
        loss = torch.empty((), dtype=torch.float32, device=sums.device)
        kind = _lib_const("EXCHANGE_FECL_TEACHER" if has_teacher else "EXCHANGE_FECL")
        ex.all_reduce_(sums, kind, inv_rows, lambda_cross, loss)
        return loss
    return fecl_loss_from_sums(all_reduce_sums(sums, group), inv_rows, lambda_cross, has_teacher)


def _lib_const(name):
    from . import _lib
    return getattr(_lib, name)


def _exchange_for(sums, group):
    if (group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
            and sums.is_cuda and sums.dtype == torch.float64 and sums.is_contiguous() and sums.numel() <= 7):
        return _peer_exchange(group, sums.device)
    return None


def fused_exchange(device, group):
    """The PeerExchange whose inboxes the *_sharded forward entry points use (the exchange then runs in the tail of
    the kernel that produces the sums), or None: no group / world of one / no peer memory."""
    if (group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
            and device.type == "cuda"):
        return _peer_exchange(group, device)
    return None


def global_batch_of(local_batch: int, global_batch, group) -> int:
    """The batch size behind the mean denominators: ``global_batch`` if given, else B_local * world size
    (so that a sharded call never mixes globally reduced sums with a local denominator)."""
    if global_batch is not None:
        return int(global_batch)
    world, _ = group_size_rank(group)
    return int(local_batch) * world


def group_size_rank(group):
    """(world size, rank) of `group`; (1, 0) without an initialised process group."""
    if group is not None and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def all_gather_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenation over ranks (rank order) of the equally shaped, contiguous `x` along dim 0."""
    world, _ = group_size_rank(group)
    if world == 1:
        return x
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def shard_bounds(global_batch: int, rank: int, world_size: int):
    """Contiguous sample range of `rank`; requires world_size <= global_batch (else: replicas only)."""
    if world_size > global_batch:
        raise ValueError(f"cannot shard a batch of {global_batch} samples over {world_size} ranks: run replicas instead")
    base, extra = divmod(global_batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
