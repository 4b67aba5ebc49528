"""Drop-in replacement for the reference's ``code/utils/dycon_losses.py`` on B200.

Same names, signatures, defaults and autograd behaviour as the reference module
(rogeliorjr/DyCON_Paper_Replication), so ``from utils import dycon_losses`` call sites
(code/train_DyCON_BraTS19.py:286-288,295,346-351 and the Pancreas / ISLES22 twins) work
unchanged; the arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/dycon_b200.h``.  There is no CPU or eager-PyTorch fallback: CPU tensors raise.

  UnCLoss().forward(s_logits, t_logits, beta)                       dycon_losses.py:50-118
  FeCLoss(device, temperature, gamma, use_focal, rampup_epochs,
          lambda_cross).forward(feat, mask, teacher_feat,
          gambling_uncertainty, epoch)                              dycon_losses.py:120-235
  adaptive_beta / sigmoid_rampup / gambling_softmax                 dycon_losses.py:8-47
  update_ema_variables(model, ema_model, alpha, global_step)        train_DyCON_BraTS19.py:155-164

Extensions that the reference does not have (all optional, keyword-only):
  * ``FeCLoss(..., precision="fp16"|"bf16"|"fp32")`` -- similarity arithmetic: tcgen05 tiles with
    fp16 (default) or bf16 operands and fp32 accumulation, or exact fp32 SIMT tiles.
  * ``process_group=`` on both modules -- the batch is sharded over ranks; partial sums and the
    batch-global hard-negative count are all-reduced (one 4-double message per step, see
    ``dycon_paper_replication_b200.sharded``).
"""
from __future__ import annotations

import ctypes
import math
import os

import torch
import torch.nn as nn

from . import _lib, sharded

__all__ = ["UnCLoss", "FeCLoss", "adaptive_beta", "sigmoid_rampup", "gambling_softmax",
           "update_ema_variables", "StepLosses", "sgd_clip_ema_step", "loss_is_finite_flag", "pooled_label_mask"]


# =========================================================================== host scalars
def adaptive_beta(epoch, total_epochs, max_beta=5.0, min_beta=0.5):
    """beta decays geometrically from max_beta (epoch 0) to min_beta (dycon_losses.py:8-12)."""
    return max_beta * ((min_beta / max_beta) ** (epoch / total_epochs))


def sigmoid_rampup(current_epoch, total_rampup_epochs, min_threshold, max_threshold, steepness=5.0):
    """exp(-steepness*(1-e/E)^2) ramp from min to max threshold (dycon_losses.py:28-47)."""
    if total_rampup_epochs == 0:
        return max_threshold
    e = max(0.0, min(float(current_epoch), total_rampup_epochs))
    phase = 1.0 - (e / total_rampup_epochs)
    return min_threshold + (max_threshold - min_threshold) * math.exp(-steepness * (phase ** 2))


def gambling_softmax(logits):
    """Un-stabilised softmax over dim 1 with a 1e-18 guard (dycon_losses.py:14-26).  Only referenced
    from commented-out code in the reference; kept as a thin PyTorch function for API parity."""
    ex = torch.exp(logits)
    return ex / (ex.sum(dim=1, keepdim=True) + 1e-18)


# =========================================================================== plumbing
_workspaces = {}
_timer = None     # {name: [(start_event, end_event), ...]} while a kernel_timer() is active


class kernel_timer:
    """Measurement aid for bench.py: CUDA events (on the launching stream) around every C-ABI call
    made while the context is active.  ``ms()`` -> {call name: [milliseconds per call]}."""

    def __enter__(self):
        global _timer
        self.records = {}
        _timer = self.records
        return self

    def __exit__(self, *exc):
        global _timer
        _timer = None
        return False

    def ms(self):
        torch.cuda.synchronize()
        return {k: [a.elapsed_time(b) for a, b in v] for k, v in self.records.items()}


def _tick():
    if _timer is None:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _tock(name, start):
    if start is not None and _timer is not None:
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        _timer.setdefault(name, []).append((start, end))


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _workspace(device, kind, nbytes):
    """Zero-filled reduction workspace, one per (device, stream, kind); kernels leave it zeroed."""
    key = (device.index, _stream_ptr(device), kind)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(int(nbytes), 16), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _require_cuda_fp32(name, x):
    if not torch.is_tensor(x):
        raise TypeError(f"{name} must be a tensor")
    if not x.is_cuda:
        raise RuntimeError(f"{name} is on {x.device}: the DyCON B200 kernels have no CPU fallback")
    if x.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {x.dtype}); the reference runs fp32 end to end")


def _ptr(x):
    return ctypes.c_void_p(x.data_ptr()) if x is not None else None


def _dense_strides(x):
    """x's strides if x is a dense (B, N, D) tensor whose rows or columns are contiguous, else None."""
    B, N, D = x.shape
    st = tuple(x.stride())
    if st in ((N * D, D, 1), (N * D, 1, N)):
        return st
    return None


def _scalar_grad(go):
    go = go.reshape(()).to(torch.float32)
    return go.contiguous()


# =========================================================================== UnCL
class _UnCLFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s_logits, t_logits, beta, process_group, global_batch):
        _require_cuda_fp32("s_logits", s_logits)
        _require_cuda_fp32("t_logits", t_logits)
        if s_logits.shape != t_logits.shape or s_logits.dim() < 3:
            raise ValueError(f"UnCLoss expects equal (B, C, spatial...) shapes, got {tuple(s_logits.shape)} "
                             f"and {tuple(t_logits.shape)}")
        if s_logits.device != t_logits.device:
            raise RuntimeError("s_logits and t_logits are on different devices")
        dev = s_logits.device
        s = s_logits.contiguous()
        t = t_logits.detach().contiguous()
        B, Cn = s.shape[0], s.shape[1]
        V = s[0, 0].numel()
        gb = sharded.global_batch_of(B, global_batch, process_group)
        inv_count = 1.0 / (gb * V)
        with torch.cuda.device(dev):
            _lib.require_b200(dev.index)
            L = _lib.lib()
            ws = _workspace(dev, "uncl", L.dycon_uncl_workspace_bytes())
            stash = torch.empty(B * V, dtype=torch.float32, device=dev) if Cn == 2 else None
            total = torch.empty(1, dtype=torch.float64, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            ex = sharded.fused_exchange(dev, process_group)
            t0 = _tick()
            args = (_ptr(s), _ptr(t), B, Cn, V, float(beta), inv_count, _ptr(stash), _ptr(total), _ptr(loss), _ptr(ws),
                    ws.numel())
            if ex is not None:      # the exchange of the partial sum runs in the tail of the forward kernel
                _lib.check(L.dycon_uncl_fwd_sharded(*args, *ex.abi_args(), _stream_ptr(dev)), "dycon_uncl_fwd_sharded")
            else:
                _lib.check(L.dycon_uncl_fwd(*args, _stream_ptr(dev)), "dycon_uncl_fwd")
            _tock("uncl_fwd", t0)
            if ex is None and process_group is not None:
                loss = sharded.reduce_uncl(total, inv_count, process_group)
        ctx.dims = (B, Cn, V, float(beta), inv_count)
        ctx.shape = s_logits.shape
        if Cn == 2:
            ctx.save_for_backward(stash)
        else:
            ctx.save_for_backward(s, t)
        return loss

    @staticmethod
    def backward(ctx, go):
        B, Cn, V, beta, inv_count = ctx.dims
        saved = ctx.saved_tensors
        dev = saved[0].device
        grad = torch.empty(ctx.shape, dtype=torch.float32, device=dev)
        go = _scalar_grad(go)
        with torch.cuda.device(dev):
            L = _lib.lib()
            t0 = _tick()
            if Cn == 2:
                rc = L.dycon_uncl_bwd(None, None, _ptr(saved[0]), B, Cn, V, beta, inv_count, _ptr(go), _ptr(grad),
                                      _stream_ptr(dev))
            else:
                rc = L.dycon_uncl_bwd(_ptr(saved[0]), _ptr(saved[1]), None, B, Cn, V, beta, inv_count, _ptr(go),
                                      _ptr(grad), _stream_ptr(dev))
            _lib.check(rc, "dycon_uncl_bwd")
            _tock("uncl_bwd", t0)
        return grad, None, None, None, None


class UnCLoss(nn.Module):
    """Uncertainty-aware student/teacher consistency (reference: dycon_losses.py:50-118).

    ``forward(s_logits, t_logits, beta)`` with logits of shape (B, C, H, W, D) returns the 0-dim
    fp32 loss.  Only ``s_logits`` receives a gradient: the teacher forward runs under
    ``torch.no_grad()`` in every reference call site (train_DyCON_BraTS19.py:305-306); a
    ``t_logits`` that requires grad raises instead of being silently dropped.
    """

    def __init__(self, process_group=None, global_batch=None):
        super().__init__()
        self.process_group = process_group
        self.global_batch = global_batch

    def forward(self, s_logits, t_logits, beta):
        if torch.is_tensor(t_logits) and t_logits.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("UnCLoss: t_logits requires grad; the teacher branch is a constant in DyCON "
                               "(detach it, as the reference's no_grad teacher forward does)")
        if torch.is_tensor(beta):
            beta = float(beta)
        return _UnCLFunction.apply(s_logits, t_logits, beta, self.process_group, self.global_batch)


# =========================================================================== fused step losses
class _StepLossesFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s_logits, t_logits, labels, labeled_bs, beta):
        dev = s_logits.device
        s = s_logits.contiguous()
        t = t_logits.detach().contiguous()
        B, Cn = s.shape[0], s.shape[1]
        V = s[0, 0].numel()
        with torch.cuda.device(dev):
            _lib.require_b200(dev.index)
            L = _lib.lib()
            ws = _workspace(dev, "segcons", L.dycon_segcons_workspace_bytes())
            sums = torch.empty(6, dtype=torch.float64, device=dev)
            out = torch.empty(4, dtype=torch.float32, device=dev)
            t0 = _tick()
            _lib.check(L.dycon_segcons_fwd(_ptr(s), _ptr(t), _ptr(labels), B, labeled_bs, Cn, V, float(beta), _ptr(sums),
                                           _ptr(out), _ptr(ws), ws.numel(), _stream_ptr(dev)), "dycon_segcons_fwd")
            _tock("segcons_fwd", t0)
        ctx.save_for_backward(s, t, sums) if labels is None else ctx.save_for_backward(s, t, sums, labels)
        ctx.cfg = (B, Cn, V, labeled_bs, float(beta), labels is not None)
        ctx.shape = s_logits.shape
        return out[0], out[1], out[2], out[3]

    @staticmethod
    def backward(ctx, g_u, g_ce, g_dice, g_cons):
        B, Cn, V, labeled_bs, beta, has_labels = ctx.cfg
        saved = ctx.saved_tensors
        s, t, sums = saved[0], saved[1], saved[2]
        labels = saved[3] if has_labels else None
        dev = s.device
        go = torch.stack([g.reshape(()).to(torch.float32) for g in (g_u, g_ce, g_dice, g_cons)]).contiguous()
        grad = torch.empty(ctx.shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            t0 = _tick()
            _lib.check(_lib.lib().dycon_segcons_bwd(_ptr(s), _ptr(t), _ptr(labels), B, labeled_bs, Cn, V, beta, _ptr(sums),
                                                    _ptr(go), _ptr(grad), _stream_ptr(dev)), "dycon_segcons_bwd")
            _tock("segcons_bwd", t0)
        return grad, None, None, None, None


class StepLosses(nn.Module):
    """The four voxel-wise losses of the reference step loop from ONE pass over the logits (an optional entry point
    beside the unchanged ``UnCLoss``; reference: train_DyCON_BraTS19.py:308-314,351-352 with utils/losses.py:8-16,65-82):

        u_loss, loss_seg, loss_seg_dice, consistency_loss = StepLosses()(stud_logits, ema_logits, label_batch, labeled_bs, beta)

    equal to ``uncl_criterion(stud_logits, ema_logits, beta)``, ``F.cross_entropy(stud_logits[:lb], label_batch[:lb])``,
    ``losses.dice_loss(stud_probs[:lb, 1], label_batch[:lb] == 1)`` and
    ``losses.softmax_mse_loss(stud_probs[lb:], ema_probs[lb:]).mean()`` (the reference passes probabilities to a
    function that takes the softmax again; reproduced).  Gradients flow to ``stud_logits`` only.  Two classes run in
    the fused kernel pair; any other class count composes the UnCL kernels with the stock PyTorch ops.  Edge cases
    the scripts never reach: ``labeled_bs == 0`` gives ``loss_seg = 0`` and ``labeled_bs == B`` gives
    ``consistency_loss = 0`` where the reference's means over empty tensors are NaN."""

    def forward(self, stud_logits, ema_logits, label_batch, labeled_bs, beta):
        _require_cuda_fp32("stud_logits", stud_logits)
        _require_cuda_fp32("ema_logits", ema_logits)
        if stud_logits.shape != ema_logits.shape or stud_logits.dim() < 3:
            raise ValueError("StepLosses expects equal (B, C, spatial...) logits")
        if torch.is_tensor(ema_logits) and ema_logits.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("StepLosses: ema_logits requires grad; the teacher branch is a constant in DyCON")
        B = stud_logits.shape[0]
        labeled_bs = int(labeled_bs)
        if not 0 <= labeled_bs <= B:
            raise ValueError(f"labeled_bs={labeled_bs} outside [0, {B}]")
        if torch.is_tensor(beta):
            beta = float(beta)
        labels = None
        if labeled_bs > 0:
            if not torch.is_tensor(label_batch) or not label_batch.is_cuda:
                raise RuntimeError("label_batch must be a CUDA tensor (no CPU fallback)")
            if label_batch.dtype != torch.int64:
                raise TypeError(f"label_batch must be int64 class indices like F.cross_entropy's target (got {label_batch.dtype})")
            if tuple(label_batch.shape[1:]) != tuple(stud_logits.shape[2:]) or label_batch.shape[0] < labeled_bs:
                raise ValueError("label_batch must be (>= labeled_bs, spatial...) matching the logits")
            labels = label_batch[:labeled_bs].contiguous()
        if stud_logits.shape[1] == 2:
            return _StepLossesFunction.apply(stud_logits, ema_logits, labels, labeled_bs, beta)
        # C != 2: the UnCL kernels + the stock terms as PyTorch ops (they stay on PyTorch in the reference design)
        import torch.nn.functional as F
        u = _UnCLFunction.apply(stud_logits, ema_logits, beta, None, None)
        ps, pt = F.softmax(stud_logits, dim=1), F.softmax(ema_logits.detach(), dim=1)
        zero = stud_logits.new_zeros(())
        ce = F.cross_entropy(stud_logits[:labeled_bs], labels) if labeled_bs > 0 else zero
        if labeled_bs > 0:
            score, target = ps[:labeled_bs, 1], (labels == 1).float()
            dice = 1 - (2 * (score * target).sum() + 1e-5) / ((score * score).sum() + (target * target).sum() + 1e-5)
        else:
            dice = zero + 1 - 1e-5 / 1e-5
        cons = ((F.softmax(ps[labeled_bs:], dim=1) - F.softmax(pt[labeled_bs:], dim=1)) ** 2).mean() if labeled_bs < B else zero
        return u, ce, dice, cons


# =========================================================================== FeCL
_PRECISIONS = {"fp32": _lib.FECL_FP32, "bf16": _lib.FECL_BF16, "fp16": _lib.FECL_FP16}


def default_fecl_precision():
    return os.environ.get("DYCON_FECL_PRECISION", "fp16")


class _FeCLFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, labels, teacher, row_weight, inv_tau, gamma, use_focal, cross_thresh,
                lambda_cross, precision, process_group, global_batch):
        dev = feat.device
        B, N, D = feat.shape
        gb = sharded.global_batch_of(B, global_batch, process_group)
        inv_rows = 1.0 / (gb * N)
        has_teacher = teacher is not None
        with torch.cuda.device(dev):
            _lib.require_b200(dev.index)
            L = _lib.lib()
            sbytes = L.dycon_fecl_state_bytes(B, N, D, int(has_teacher), precision)
            wbytes = L.dycon_fecl_workspace_bytes(B, N, D, precision)
            state = torch.empty(max(sbytes, 16), dtype=torch.uint8, device=dev)
            ws = _workspace(dev, f"fecl{precision}", wbytes)
            sums = torch.empty(3, dtype=torch.float64, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            ts = teacher.stride() if has_teacher else (0, 0, 0)
            ex = sharded.fused_exchange(dev, process_group)
            t0 = _tick()
            args = (_ptr(feat), *feat.stride(), _ptr(teacher), *ts, _ptr(labels), _ptr(row_weight), B, N, D, inv_tau, gamma,
                    int(use_focal), cross_thresh, lambda_cross, inv_rows, precision, _ptr(state), state.numel(), _ptr(sums),
                    _ptr(loss), _ptr(ws), ws.numel())
            if ex is not None:      # the three sums are exchanged in the tail of the loss sweep: `sums` / `loss` are global
                _lib.check(L.dycon_fecl_fwd_sharded(*args, *ex.abi_args(), _stream_ptr(dev)), "dycon_fecl_fwd_sharded")
            else:
                _lib.check(L.dycon_fecl_fwd(*args, _stream_ptr(dev)), "dycon_fecl_fwd")
            _tock("fecl_fwd", t0)
            if ex is None and process_group is not None:
                loss = sharded.reduce_fecl(sums, inv_rows, lambda_cross, has_teacher, process_group)
        ctx.save_for_backward(state, labels, sums)
        ctx.cfg = (B, N, D, has_teacher, inv_tau, gamma, int(use_focal), row_weight is not None, cross_thresh,
                   lambda_cross, precision)
        # the gradient is written in feat's own layout when that is dense ((D*N, 1, N) for the caller's
        # normalize(transpose(view)), train_DyCON_BraTS19.py:316-323): autograd then hands it on as is
        ctx.grad_strides = _dense_strides(feat)
        return loss

    @staticmethod
    def backward(ctx, go):
        state, labels, sums = ctx.saved_tensors
        B, N, D, has_teacher, inv_tau, gamma, use_focal, has_rw, cross_thresh, lambda_cross, precision = ctx.cfg
        dev = state.device
        if ctx.grad_strides is None:
            grad = torch.empty((B, N, D), dtype=torch.float32, device=dev)
        else:
            grad = torch.empty_strided((B, N, D), ctx.grad_strides, dtype=torch.float32, device=dev)
        go = _scalar_grad(go)
        with torch.cuda.device(dev):
            t0 = _tick()
            _lib.check(_lib.lib().dycon_fecl_bwd(_ptr(state), state.numel(), _ptr(labels), B, N, D, int(has_teacher),
                                                 inv_tau, gamma, use_focal, int(has_rw), cross_thresh, lambda_cross,
                                                 precision, ctypes.c_void_p(sums.data_ptr() + 16), _ptr(go),
                                                 _ptr(grad), *grad.stride(), _stream_ptr(dev)),
                       "dycon_fecl_bwd")
            _tock("fecl_bwd", t0)
        return (grad,) + (None,) * 11


class _FeCLGlobalFunction(torch.autograd.Function):
    """FeCL with global negatives: the samples of ALL ranks are contrasted as one sample of B_all*N rows (the
    reference FeCLoss on feat.reshape(1, B_all*N, D)); every rank owns the rows of its local samples.  See
    include/dycon_b200.h, "FeCL with global negatives", for the phase protocol."""

    @staticmethod
    def forward(ctx, feat, labels, teacher, row_weight, inv_tau, gamma, use_focal, cross_thresh, lambda_cross,
                precision, process_group):
        dev = feat.device
        B, N, D = feat.shape
        world, rank = sharded.group_size_rank(process_group)
        B_all, M = B * world, B * world * N
        has_teacher = teacher is not None
        gather = lambda x: sharded.all_gather_rows(x.contiguous(), process_group)
        feat_all = gather(feat.detach())                        # (B_all, N, D) row-major
        teacher_all = gather(teacher) if has_teacher else None
        labels_all = gather(labels)
        rw_all = gather(row_weight) if row_weight is not None else None
        lo, hi = rank * B * N, (rank + 1) * B * N
        with torch.cuda.device(dev):
            _lib.require_b200(dev.index)
            L = _lib.lib()
            sbytes = L.dycon_fecl_gn_state_bytes(B_all, N, D, int(has_teacher), precision)
            if sbytes == 0:
                raise RuntimeError("FeCLoss(cross_gpu_negatives=True) needs precision 'fp16' or 'bf16' and D % 4 == 0, D <= 256")
            off = (ctypes.c_size_t * 6)()
            _lib.check(L.dycon_fecl_gn_layout(B_all, N, D, int(has_teacher), precision, off), "dycon_fecl_gn_layout")
            state = torch.zeros(sbytes, dtype=torch.uint8, device=dev)
            plane = lambda k: state[off[k]:off[k] + 4 * M].view(torch.float32)
            ws = _workspace(dev, f"fecl{precision}", L.dycon_fecl_workspace_bytes(1, M, D, precision))
            sums = torch.empty(3, dtype=torch.float64, device=dev)
            ts = teacher_all.stride() if has_teacher else (0, 0, 0)

            def phases(mask):
                _lib.check(L.dycon_fecl_gn_fwd(mask, _ptr(feat_all), *feat_all.stride(), _ptr(teacher_all), *ts,
                                               _ptr(labels_all), _ptr(rw_all), B_all, N, D, inv_tau, gamma,
                                               int(use_focal), cross_thresh, lambda_cross, precision, _ptr(state),
                                               state.numel(), lo, hi, _ptr(sums), _ptr(ws), ws.numel(),
                                               _stream_ptr(dev)), "dycon_fecl_gn_fwd")

            t0 = _tick()
            phases(1 | 2)                                       # pack, row max of the own rows
            m = plane(1)
            m.copy_(sharded.all_gather_rows(m[lo:hi].clone(), process_group))
            phases(4 | 8)                                       # negative sums, loss terms of the own rows
            _tock("fecl_fwd", t0)
            loss = sharded.reduce_fecl(sums, 1.0 / M, lambda_cross, has_teacher, process_group)
            if ctx.needs_input_grad[0]:
                # the backward reads n, kappa and A of EVERY row (the transposed term): consolidate the own rows'
                # split partials of A into plane 0 (fixed order), then all-gather the three planes
                a_planes = state[off[4]:off[4] + 8 * 4 * M].view(torch.float32).view(8, M)
                own = torch.stack([plane(2)[lo:hi], plane(3)[lo:hi], a_planes[:, lo:hi].sum(dim=0)])
                full = sharded.all_gather_rows(own.unsqueeze(0), process_group)      # (world, 3, B*N)
                plane(2).copy_(full[:, 0].reshape(-1))
                plane(3).copy_(full[:, 1].reshape(-1))
                a_planes[0].copy_(full[:, 2].reshape(-1))
                state[off[0] + 4:off[0] + 8].view(torch.float32).fill_(1.0)           # header: one A plane (no H2D copy: graph capturable)
        ctx.save_for_backward(state, labels_all, sums)
        ctx.cfg = (B, N, D, B_all, has_teacher, inv_tau, gamma, int(use_focal), row_weight is not None, cross_thresh,
                   lambda_cross, precision, lo, hi)
        ctx.grad_strides = _dense_strides(feat)
        return loss

    @staticmethod
    def backward(ctx, go):
        state, labels_all, sums = ctx.saved_tensors
        (B, N, D, B_all, has_teacher, inv_tau, gamma, use_focal, has_rw, cross_thresh, lambda_cross, precision, lo,
         hi) = ctx.cfg
        dev = state.device
        if ctx.grad_strides is None:
            grad = torch.empty((B, N, D), dtype=torch.float32, device=dev)
        else:
            grad = torch.empty_strided((B, N, D), ctx.grad_strides, dtype=torch.float32, device=dev)
        go = _scalar_grad(go)
        with torch.cuda.device(dev):
            t0 = _tick()
            _lib.check(_lib.lib().dycon_fecl_gn_bwd(_ptr(state), state.numel(), _ptr(labels_all), B_all, N, D,
                                                    int(has_teacher), inv_tau, gamma, use_focal, int(has_rw),
                                                    cross_thresh, lambda_cross, precision, lo, hi,
                                                    ctypes.c_void_p(sums.data_ptr() + 16), _ptr(go), _ptr(grad),
                                                    *grad.stride(), _stream_ptr(dev)), "dycon_fecl_gn_bwd")
            _tock("fecl_bwd", t0)
        return (grad,) + (None,) * 10


class _FeCLFeaturesFunction(torch.autograd.Function):
    """FeCL on the RAW feature volumes: ``F.normalize(features.view(B, C, -1).transpose(1, 2), dim=-1)``
    of student and teacher (train_DyCON_BraTS19.py:316-323) is folded into FeCL's operand staging as a per-row
    scale, and its Jacobian is applied to FeCL's gradient by one kernel -- the normalised fp32 embeddings and the
    autograd chain through them never exist."""

    @staticmethod
    def forward(ctx, features, teacher_features, labels, row_weight, inv_tau, gamma, use_focal, cross_thresh,
                lambda_cross, precision, process_group, global_batch):
        dev = features.device
        x = features.contiguous()
        B, D = x.shape[0], x.shape[1]
        N = x[0, 0].numel()
        xv = x.view(B, D, N).transpose(1, 2)                    # (B, N, D) with strides (D*N, 1, N), no copy
        has_teacher = teacher_features is not None
        tv = teacher_features.contiguous().view(B, D, N).transpose(1, 2) if has_teacher else None
        gb = sharded.global_batch_of(B, global_batch, process_group)
        inv_rows = 1.0 / (gb * N)
        with torch.cuda.device(dev):
            _lib.require_b200(dev.index)
            L = _lib.lib()
            t0 = _tick()
            inv_s = torch.empty(B * N, dtype=torch.float32, device=dev)
            _lib.check(L.dycon_row_inv_norm(_ptr(xv), *xv.stride(), B, N, D, _ptr(inv_s), _stream_ptr(dev)), "dycon_row_inv_norm")
            inv_t = None
            if has_teacher:
                inv_t = torch.empty(B * N, dtype=torch.float32, device=dev)
                _lib.check(L.dycon_row_inv_norm(_ptr(tv), *tv.stride(), B, N, D, _ptr(inv_t), _stream_ptr(dev)),
                           "dycon_row_inv_norm")
            sbytes = L.dycon_fecl_state_bytes(B, N, D, int(has_teacher), precision)
            wbytes = L.dycon_fecl_workspace_bytes(B, N, D, precision)
            state = torch.empty(max(sbytes, 16), dtype=torch.uint8, device=dev)
            ws = _workspace(dev, f"fecl{precision}", wbytes)
            sums = torch.empty(3, dtype=torch.float64, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            ts = tv.stride() if has_teacher else (0, 0, 0)
            ex = sharded.fused_exchange(dev, process_group)
            xargs = ex.abi_args() if ex is not None else (None, 0, 1, None, -1.0)
            _lib.check(L.dycon_fecl_fwd_scaled(_ptr(xv), *xv.stride(), _ptr(tv), *ts, _ptr(inv_s), _ptr(inv_t), _ptr(labels),
                                               _ptr(row_weight), B, N, D, inv_tau, gamma, int(use_focal), cross_thresh,
                                               lambda_cross, inv_rows, precision, _ptr(state), state.numel(), _ptr(sums),
                                               _ptr(loss), _ptr(ws), ws.numel(), *xargs, _stream_ptr(dev)),
                       "dycon_fecl_fwd_scaled")
            _tock("fecl_fwd", t0)
            if ex is None and process_group is not None:
                loss = sharded.reduce_fecl(sums, inv_rows, lambda_cross, has_teacher, process_group)
        ctx.save_for_backward(state, labels, sums, x, inv_s)
        ctx.cfg = (B, N, D, has_teacher, inv_tau, gamma, int(use_focal), row_weight is not None, cross_thresh,
                   lambda_cross, precision)
        ctx.shape = features.shape
        return loss

    @staticmethod
    def backward(ctx, go):
        state, labels, sums, x, inv_s = ctx.saved_tensors
        B, N, D, has_teacher, inv_tau, gamma, use_focal, has_rw, cross_thresh, lambda_cross, precision = ctx.cfg
        dev = state.device
        xv = x.view(B, D, N).transpose(1, 2)
        g = torch.empty_strided((B, N, D), (N * D, 1, N), dtype=torch.float32, device=dev)     # gradient w.r.t. the unit rows
        dx = torch.empty_like(x)
        dxv = dx.view(B, D, N).transpose(1, 2)
        go = _scalar_grad(go)
        with torch.cuda.device(dev):
            L = _lib.lib()
            t0 = _tick()
            _lib.check(L.dycon_fecl_bwd(_ptr(state), state.numel(), _ptr(labels), B, N, D, int(has_teacher), inv_tau, gamma,
                                        use_focal, int(has_rw), cross_thresh, lambda_cross, precision,
                                        ctypes.c_void_p(sums.data_ptr() + 16), _ptr(go), _ptr(g), *g.stride(),
                                        _stream_ptr(dev)), "dycon_fecl_bwd")
            _lib.check(L.dycon_normalize_bwd(_ptr(xv), *xv.stride(), _ptr(g), *g.stride(), _ptr(inv_s), B, N, D, _ptr(dxv),
                                             *dxv.stride(), _stream_ptr(dev)), "dycon_normalize_bwd")
            _tock("fecl_bwd", t0)
        return (dx.view(ctx.shape),) + (None,) * 11


def pooled_label_mask(label_batch, feature_spatial):
    """``(F.avg_pool3d(label.float(), k, k) > 0.5).float().reshape(B, -1)`` with per-axis kernels
    k = label extent // feature extent (train_DyCON_BraTS19.py:326-330; train_DyCON_ISLES22.py:268-281), one kernel,
    straight from the int64 / uint8 / float32 label volume.  Returns (B, N) fp32 in {0, 1}."""
    if not torch.is_tensor(label_batch) or not label_batch.is_cuda:
        raise RuntimeError("label_batch must be a CUDA tensor (no CPU fallback)")
    if label_batch.dim() != 4 or len(feature_spatial) != 3:
        raise ValueError("label_batch must be (B, H, W, D) and feature_spatial three extents")
    codes = {torch.int64: _lib.LABEL_INT64, torch.float32: _lib.LABEL_FLOAT32, torch.uint8: _lib.LABEL_UINT8}
    if label_batch.dtype not in codes:
        raise TypeError(f"label_batch must be int64, uint8 or float32 (got {label_batch.dtype})")
    lab = label_batch.contiguous()
    B, H, W, Dz = lab.shape
    k = [ext // f for ext, f in zip((H, W, Dz), feature_spatial)]
    if min(k) < 1 or any(ext // kk != f for ext, kk, f in zip((H, W, Dz), k, feature_spatial)):
        raise ValueError(f"label volume {(H, W, Dz)} does not pool onto the feature grid {tuple(feature_spatial)}")
    dev = lab.device
    out = torch.empty(B, feature_spatial[0] * feature_spatial[1] * feature_spatial[2], dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.require_b200(dev.index)
        _lib.check(_lib.lib().dycon_pool_mask(_ptr(lab), codes[lab.dtype], B, H, W, Dz, k[0], k[1], k[2], _ptr(out),
                                              _stream_ptr(dev)), "dycon_pool_mask")
    return out


class FeCLoss(nn.Module):
    """Feature contrastive loss with focal positives and the teacher hard-negative branch
    (reference: dycon_losses.py:120-235; constructor :141-148, forward :150-235).

    feat (B,N,D) student embeddings (any strides), mask (B,1,N) labels, teacher_feat (B,N,D) or
    None, gambling_uncertainty (B,N) or None (a per-row weight that also disables the focal
    weights, :209-211), epoch -> the sigmoid_rampup thresholds.  Returns the 0-dim fp32 loss;
    the gradient flows to ``feat`` only.
    """

    def __init__(self, device, temperature=0.6, gamma=2.0, use_focal=False, rampup_epochs=2000,
                 lambda_cross=1.0, *, precision=None, process_group=None, global_batch=None,
                 cross_gpu_negatives=False):
        super().__init__()
        self.device = device
        self.temperature = temperature
        self.gamma = gamma
        self.use_focal = use_focal
        self.rampup_epochs = rampup_epochs
        self.lambda_cross = lambda_cross
        self.precision = precision or default_fecl_precision()
        if self.precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self.process_group = process_group
        self.global_batch = global_batch
        # extension (BASELINE config 5): contrast every row against the rows of ALL samples of ALL ranks
        self.cross_gpu_negatives = bool(cross_gpu_negatives)

    def forward(self, feat, mask, teacher_feat=None, gambling_uncertainty=None, epoch=0):
        _require_cuda_fp32("feat", feat)
        if feat.dim() != 3:
            raise ValueError(f"feat must be (B, N, D), got {tuple(feat.shape)}")
        B, N, D = feat.shape
        grad_on = torch.is_grad_enabled()
        if teacher_feat is not None:
            _require_cuda_fp32("teacher_feat", teacher_feat)
            if teacher_feat.shape != feat.shape:
                raise ValueError("teacher_feat must have the shape of feat")
            if teacher_feat.requires_grad and grad_on:
                raise RuntimeError("FeCLoss: teacher_feat requires grad; the teacher is a constant in DyCON")
            teacher_feat = teacher_feat.detach()
        if not torch.is_tensor(mask) or not mask.is_cuda:
            raise RuntimeError("mask must be a CUDA tensor (no CPU fallback)")
        if mask.is_floating_point() and mask.requires_grad and grad_on:
            raise RuntimeError("FeCLoss: mask requires grad; labels are constants")
        if mask.numel() != B * N:
            raise ValueError(f"mask must be (B, 1, N) = ({B}, 1, {N}), got {tuple(mask.shape)}")
        labels = mask.detach().reshape(B, N).to(torch.float32).contiguous()
        rw = None
        if gambling_uncertainty is not None:
            if gambling_uncertainty.requires_grad and grad_on:
                raise RuntimeError("FeCLoss: gambling_uncertainty requires grad; it is a constant weight")
            if gambling_uncertainty.numel() != B * N:
                raise ValueError("gambling_uncertainty must be (B, N)")
            rw = gambling_uncertainty.detach().reshape(B, N).to(device=feat.device, dtype=torch.float32).contiguous()
        # host scalars, recomputed every call like the reference (:199-200, :222)
        cross_thresh = sigmoid_rampup(epoch, self.rampup_epochs, min_threshold=0.3, max_threshold=0.5)
        if self.cross_gpu_negatives:
            return _FeCLGlobalFunction.apply(feat, labels, teacher_feat, rw, 1.0 / float(self.temperature),
                                             float(self.gamma), bool(self.use_focal), float(cross_thresh),
                                             float(self.lambda_cross), _PRECISIONS[self.precision], self.process_group)
        return _FeCLFunction.apply(feat, labels, teacher_feat, rw, 1.0 / float(self.temperature), float(self.gamma),
                                   bool(self.use_focal), float(cross_thresh), float(self.lambda_cross),
                                   _PRECISIONS[self.precision], self.process_group, self.global_batch)


def _fecl_from_features(self, stud_features, label_batch, ema_features=None, gambling_uncertainty=None, epoch=0):
    """``FeCLoss.from_features``: the loss straight from what the network returns and the loader delivers.

    Equivalent to the reference's caller-side preparation followed by ``forward`` (train_DyCON_BraTS19.py:316-350)::

        emb  = F.normalize(stud_features.view(B, C, -1).transpose(1, 2), dim=-1)       # same for ema_features
        mask = (F.avg_pool3d(label_batch.float(), k, k) > 0.5).float().reshape(B, -1).unsqueeze(1)
        loss = fecl(feat=emb, mask=mask, teacher_feat=ema_emb, gambling_uncertainty=..., epoch=epoch)

    with k = label extent // feature extent per axis.  The gradient flows to ``stud_features``."""
    _require_cuda_fp32("stud_features", stud_features)
    if stud_features.dim() != 5:
        raise ValueError(f"stud_features must be (B, C, h, w, d), got {tuple(stud_features.shape)}")
    B, C = stud_features.shape[:2]
    N = stud_features[0, 0].numel()
    grad_on = torch.is_grad_enabled()
    if ema_features is not None:
        _require_cuda_fp32("ema_features", ema_features)
        if ema_features.shape != stud_features.shape:
            raise ValueError("ema_features must have the shape of stud_features")
        if ema_features.requires_grad and grad_on:
            raise RuntimeError("FeCLoss: ema_features requires grad; the teacher is a constant in DyCON")
        ema_features = ema_features.detach()
    labels = pooled_label_mask(label_batch, tuple(stud_features.shape[2:]))
    rw = None
    if gambling_uncertainty is not None:
        if gambling_uncertainty.requires_grad and grad_on:
            raise RuntimeError("FeCLoss: gambling_uncertainty requires grad; it is a constant weight")
        if gambling_uncertainty.numel() != B * N:
            raise ValueError("gambling_uncertainty must be (B, N)")
        rw = gambling_uncertainty.detach().reshape(B, N).to(device=stud_features.device, dtype=torch.float32).contiguous()
    if self.cross_gpu_negatives:
        raise RuntimeError("FeCLoss.from_features does not combine with cross_gpu_negatives (use forward)")
    cross_thresh = sigmoid_rampup(epoch, self.rampup_epochs, min_threshold=0.3, max_threshold=0.5)
    return _FeCLFeaturesFunction.apply(stud_features, ema_features, labels, rw, 1.0 / float(self.temperature),
                                       float(self.gamma), bool(self.use_focal), float(cross_thresh),
                                       float(self.lambda_cross), _PRECISIONS[self.precision], self.process_group,
                                       self.global_batch)


FeCLoss.from_features = _fecl_from_features


# =========================================================================== EMA
class _EmaPlan:
    """ctypes pointer tables for one (student, teacher) parameter set, rebuilt only if a pointer moves."""

    def __init__(self):
        self.key = None
        self.arrays = None

    def tables(self, ema_params, params):
        key = tuple(p.data_ptr() for p in ema_params) + tuple(p.data_ptr() for p in params)
        if key != self.key:
            n = len(params)
            e = (ctypes.c_void_p * n)(*[p.data_ptr() for p in ema_params])
            s = (ctypes.c_void_p * n)(*[p.data_ptr() for p in params])
            c = (ctypes.c_int64 * n)(*[p.numel() for p in params])
            self.key, self.arrays = key, (e, s, c, n)
        return self.arrays


_ema_plans = {}


def _dense_like(a, b):
    """Same shape and strides, and dense in memory (contiguous or channels-last)."""
    if a.shape != b.shape or a.stride() != b.stride():
        return False
    if a.is_contiguous():
        return True
    if a.dim() == 4:
        return a.is_contiguous(memory_format=torch.channels_last)
    if a.dim() == 5:
        return a.is_contiguous(memory_format=torch.channels_last_3d)
    return False


def update_ema_variables(model, ema_model, alpha, global_step):
    """Mean-teacher update, one multi-tensor launch (reference: train_DyCON_BraTS19.py:155-164).

    alpha = min(1 - 1/(global_step+1), alpha); ema = ema*alpha + (1-alpha)*param over
    ``parameters()`` only (buffers are not averaged), unwrapping ``.module`` like the reference.
    """
    alpha = min(1 - 1 / (global_step + 1), alpha)
    src = model.module if hasattr(model, "module") else model
    dst = ema_model.module if hasattr(ema_model, "module") else ema_model
    params = [p.data for p in src.parameters()]
    ema_params = [p.data for p in dst.parameters()]
    n = min(len(params), len(ema_params))          # zip() semantics of the reference loop
    params, ema_params = params[:n], ema_params[:n]
    if n == 0:
        return
    dev = ema_params[0].device
    for e, p in zip(ema_params, params):
        _require_cuda_fp32("ema parameter", e)
        _require_cuda_fp32("parameter", p)
        if e.device != dev or p.device != dev:
            raise RuntimeError("update_ema_variables: parameters span several devices")
        if not _dense_like(e, p):
            raise RuntimeError("update_ema_variables: student/teacher parameters must be dense with equal strides")
    plan = _ema_plans.setdefault((id(src), id(dst)), _EmaPlan())
    e_arr, p_arr, c_arr, n = plan.tables(ema_params, params)
    with torch.cuda.device(dev):
        _lib.require_b200(dev.index)
        t0 = _tick()
        _lib.check(_lib.lib().dycon_ema_multi(e_arr, p_arr, c_arr, n, float(alpha), float(1 - alpha),
                                              _stream_ptr(dev)), "dycon_ema_multi")
        _tock("ema", t0)


# =========================================================================== clip + SGD + EMA, finite flag (8 f3/f4)
def loss_is_finite_flag(*losses, counter=None):
    """Device-side replacement of ``if torch.isnan(loss) or torch.isinf(loss): continue``
    (train_DyCON_BraTS19.py:360-362): returns an int32 device tensor that is 1 when any of the given 0-dim fp32
    CUDA tensors is NaN / +-inf -- no ``.item()``, no host synchronisation.  Hand it to ``sgd_clip_ema_step`` as
    ``skip_flag``: a raised flag turns the whole update into a no-op, which is what the reference's ``continue``
    does.  ``counter`` (int64 device tensor of one element, optional) counts the skipped steps for later logging."""
    import ctypes
    if not 1 <= len(losses) <= 16:
        raise ValueError("1..16 loss scalars")
    vals = []
    for x in losses:
        _require_cuda_fp32("loss", x)
        if x.numel() != 1:
            raise ValueError("loss_is_finite_flag takes 0-dim / one-element tensors")
        vals.append(x.detach().reshape(1).contiguous())
    dev = vals[0].device
    flag = torch.empty(1, dtype=torch.int32, device=dev)
    table = (ctypes.c_void_p * len(vals))(*[v.data_ptr() for v in vals])
    with torch.cuda.device(dev):
        _lib.require_b200(dev.index)
        _lib.check(_lib.lib().dycon_finite_check(table, len(vals), _ptr(flag), _ptr(counter), _stream_ptr(dev)),
                   "dycon_finite_check")
    flag._dycon_keepalive = vals          # the kernel reads them asynchronously
    return flag


def sgd_clip_ema_step(optimizer, model, ema_model, max_norm, ema_decay, global_step, *, skip_flag=None,
                      scale_grads=False):
    """``clip_grad_norm_(params, max_norm); optimizer.step(); update_ema_variables(model, ema_model, ema_decay,
    global_step)`` (train_DyCON_BraTS19.py:366-372) as TWO multi-tensor launches: the gradient norm, then one pass
    that reads every parameter, gradient, momentum buffer and teacher tensor once.

    ``optimizer`` must be a ``torch.optim.SGD`` (dampening 0, not maximize); its hyper-parameters are read from
    ``param_groups`` every call (the scripts change ``lr`` per iteration) and its momentum buffers live in
    ``optimizer.state[p]['momentum_buffer']`` exactly as after ``optimizer.step()``, so ``state_dict()`` stays
    interchangeable.  ``ema_model`` may be None (no teacher).  Returns the total gradient norm (0-dim device
    tensor, like ``clip_grad_norm_``).  ``skip_flag``: see ``loss_is_finite_flag``.  ``scale_grads=True`` also leaves
    the clipped values in ``.grad`` as ``clip_grad_norm_`` does (the scripts zero them right after)."""
    import ctypes
    if not isinstance(optimizer, torch.optim.SGD):
        raise TypeError("sgd_clip_ema_step drives torch.optim.SGD (the reference's optimizer)")
    alpha = min(1 - 1 / (global_step + 1), ema_decay)
    src = model.module if hasattr(model, "module") else model
    ema_of = {}
    if ema_model is not None:
        dst = ema_model.module if hasattr(ema_model, "module") else ema_model
        for e, p in zip(dst.parameters(), src.parameters()):       # zip() semantics of the reference loop
            ema_of[id(p)] = e
    seen = set()
    groups = []
    dev = None
    for group in optimizer.param_groups:
        if group.get("dampening", 0) != 0 or group.get("maximize", False):
            raise ValueError("sgd_clip_ema_step: dampening / maximize are not supported (the reference uses neither)")
        fresh, warm = ([], [], [], []), ([], [], [], [])     # (params, grads, buffers, teachers): first step of a buffer / later steps
        for p in group["params"]:
            seen.add(id(p))
            _require_cuda_fp32("parameter", p.data)
            dev = dev or p.device
            if p.device != dev:
                raise RuntimeError("sgd_clip_ema_step: parameters span several devices")
            g = p.grad
            if g is not None:
                _require_cuda_fp32("gradient", g)
                if g.is_sparse or not _dense_like(g, p.data):
                    raise RuntimeError("sgd_clip_ema_step: gradients must be dense with the parameter's strides")
            buf, is_first = None, False
            if g is not None and group["momentum"] != 0:
                st = optimizer.state[p]
                buf = st.get("momentum_buffer")
                is_first = buf is None
                if is_first:                                  # torch.optim.SGD: buf = clone(grad) on the first step
                    # zeros, not empty: if skip_flag drops this very step, the next one runs buf = 0 * mu + g = g
                    buf = st["momentum_buffer"] = torch.zeros_like(p.data)
            e = ema_of.get(id(p))
            if e is not None:
                _require_cuda_fp32("ema parameter", e.data)
                if not _dense_like(e.data, p.data):
                    raise RuntimeError("sgd_clip_ema_step: student/teacher parameters must be dense with equal strides")
            dstl = fresh if is_first else warm
            dstl[0].append(p.data); dstl[1].append(g); dstl[2].append(buf); dstl[3].append(None if e is None else e.data)
        for lists, first in ((fresh, True), (warm, False)):
            if lists[0]:
                groups.append((group, *lists, first))
    # teacher tensors whose student parameter is not in the optimizer still follow it (plain EMA)
    rest = [(p, ema_of[id(p)]) for p in src.parameters() if id(p) in ema_of and id(p) not in seen]
    if dev is None:
        return torch.zeros(())
    tab = lambda xs: (ctypes.c_void_p * len(xs))(*[None if x is None else x.data_ptr() for x in xs])
    with torch.cuda.device(dev):
        _lib.require_b200(dev.index)
        L = _lib.lib()
        all_g = [g for _, _, gs, _, _, _ in groups for g in gs if g is not None]
        clip = torch.empty(2, dtype=torch.float32, device=dev)
        ws = _workspace(dev, "gradnorm", L.dycon_grad_norm_workspace_bytes())
        t0 = _tick()
        counts = (ctypes.c_int64 * max(len(all_g), 1))(*[g.numel() for g in all_g])
        _lib.check(L.dycon_grad_norm(tab(all_g), counts, len(all_g), float(max_norm if max_norm is not None else 0.0),
                                     _ptr(clip), _ptr(ws), ws.numel(), _stream_ptr(dev)), "dycon_grad_norm")
        for group, ps, gs, bs, es, first in groups:
            if not ps:
                continue
            n = (ctypes.c_int64 * len(ps))(*[p.numel() for p in ps])
            _lib.check(L.dycon_sgd_ema_step(tab(ps), tab(gs), tab(bs), tab(es), n, len(ps), float(group["lr"]),
                                            float(group["momentum"]), float(group["weight_decay"]),
                                            int(bool(group.get("nesterov", False))), int(first), float(alpha),
                                            float(1 - alpha), _ptr(clip), _ptr(skip_flag), int(bool(scale_grads)),
                                            _stream_ptr(dev)), "dycon_sgd_ema_step")
        if rest:
            ps, es = [p.data for p, _ in rest], [e.data for _, e in rest]
            n = (ctypes.c_int64 * len(ps))(*[p.numel() for p in ps])
            _lib.check(L.dycon_sgd_ema_step(tab(ps), None, None, tab(es), n, len(ps), 0.0, 0.0, 0.0, 0, 0, float(alpha),
                                            float(1 - alpha), None, _ptr(skip_flag), 0, _stream_ptr(dev)),
                       "dycon_sgd_ema_step")
        _tock("sgd_ema_step", t0)
    return clip[0]
