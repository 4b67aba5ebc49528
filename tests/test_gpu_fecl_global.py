"""FeCL with global negatives (BASELINE config 5; extension, not in the reference): every row is contrasted
against the rows of all samples of the (global) batch.  Oracle: the reference FeCL on
feat.reshape(1, B*N, D), mask.reshape(1, 1, B*N).

(1) the module with a world of one rank; (2) the C-ABI phase protocol driven for TWO virtual ranks on one GPU
(each owns half of the samples and its own state; the all-gathers between the phases are slice copies) -- the
real multi-process run is tools/multi_gpu_check.py."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import normwise
from oracle import closed_form, torch_port

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]
CTOR = dict(temperature=0.6, gamma=2.0, use_focal=True, rampup_epochs=1500)


def inputs(b, n, d, seed):
    g = torch.Generator().manual_seed(seed)
    mask = (torch.rand(b, 1, n, generator=g) < 0.3).float()
    base = torch.randn(1, 1, d, generator=g)
    f = torch.nn.functional.normalize(base + 0.8 * torch.randn(b, n, d, generator=g) + 0.5 * mask.transpose(1, 2), dim=-1)
    t = torch.nn.functional.normalize(f + 0.1 * torch.randn(b, n, d, generator=g), dim=-1)
    return f, mask, t


def oracle(f, mask, t, epoch, go):
    b, n, d = f.shape
    thr = torch_port.ramp_threshold(epoch, 1500, 0.3, 0.5)
    return closed_form.fecl(f.reshape(1, b * n, d).numpy(), mask.reshape(1, 1, b * n).numpy(),
                            t.reshape(1, b * n, d).numpy(), None, inv_tau=1 / 0.6, gamma=2.0, use_focal=True,
                            cross_thresh=thr, go=go, ambiguity=5e-4)


def strict(got, ref, f, t):
    """Threshold flips predicted from the fp16-rounded operands, never fitted (closed_form.fecl_grad_error_strict)."""
    b, n, d = f.shape
    thr = torch_port.ramp_threshold(100, 1500, 0.3, 0.5)
    st = closed_form.fecl_grad_error_strict(got, ref, f.reshape(1, b * n, d).numpy(), t.reshape(1, b * n, d).numpy(), "fp16", thr)
    assert st["outside_flips"] == 0 and st["flipped"] + st["free"] <= st["window_pairs"], st
    assert st["err"] <= 2e-3, st


@pytest.mark.parametrize("shape", [(3, 80, 32), (2, 300, 256)])
def test_module_world_of_one_matches_the_merged_reference(shape):
    from dycon_paper_replication_b200 import FeCLoss
    f, mask, t = inputs(*shape, seed=21)
    crit = FeCLoss("cuda", precision="fp16", cross_gpu_negatives=True, **CTOR)
    x = f.cuda().requires_grad_(True)
    loss = crit(x, mask.cuda(), t.cuda(), None, 100)
    (0.5 * loss).backward()
    ref = oracle(f, mask, t, 100, 0.5)
    assert abs(loss.item() - ref["loss"]) <= 2e-3 * abs(ref["loss"])
    got = x.grad.cpu().numpy().reshape(ref["grad"].shape)
    strict(got, ref, f, t)


def test_two_virtual_ranks_through_the_c_abi():
    from dycon_paper_replication_b200 import _lib, dycon_losses
    L = _lib.lib()
    B_all, N, D, W = 4, 96, 64, 2
    M, Bl = B_all * N, B_all // W
    f, mask, t = inputs(B_all, N, D, seed=5)
    fa, ta, la = f.cuda(), t.cuda(), mask.reshape(B_all, N).cuda().contiguous()
    prec = _lib.FECL_FP16
    thr = dycon_losses.sigmoid_rampup(100, 1500, 0.3, 0.5)
    P = lambda x: ctypes.c_void_p(x.data_ptr())
    stream = torch.cuda.current_stream().cuda_stream
    sbytes = L.dycon_fecl_gn_state_bytes(B_all, N, D, 1, prec)
    off = (ctypes.c_size_t * 6)()
    assert L.dycon_fecl_gn_layout(B_all, N, D, 1, prec, off) == 0
    states = [torch.zeros(sbytes, dtype=torch.uint8, device="cuda") for _ in range(W)]
    wss = [torch.zeros(L.dycon_fecl_workspace_bytes(1, M, D, prec), dtype=torch.uint8, device="cuda") for _ in range(W)]
    sums = [torch.zeros(3, dtype=torch.float64, device="cuda") for _ in range(W)]
    rows = [(r * Bl * N, (r + 1) * Bl * N) for r in range(W)]
    plane = lambda st, k: st[off[k]:off[k] + 4 * M].view(torch.float32)

    def fwd(r, mask_bits):
        lo, hi = rows[r]
        _lib.check(L.dycon_fecl_gn_fwd(mask_bits, P(fa), *fa.stride(), P(ta), *ta.stride(), P(la), None, B_all, N, D,
                                       1 / 0.6, 2.0, 1, thr, 1.0, prec, P(states[r]), sbytes, lo, hi, P(sums[r]),
                                       P(wss[r]), wss[r].numel(), stream), "gn_fwd")

    for r in range(W):
        fwd(r, 1 | 2)
    m_full = torch.cat([plane(states[r], 1)[rows[r][0]:rows[r][1]] for r in range(W)])     # "all-gather" of m
    for r in range(W):
        plane(states[r], 1).copy_(m_full)
        fwd(r, 4 | 8)
    total = sums[0] + sums[1]                                                              # "all-reduce"
    loss = (total[0] / M + total[1] / (total[2] + 1e-18)).item()
    own = []
    for r in range(W):
        lo, hi = rows[r]
        a8 = states[r][off[4]:off[4] + 8 * 4 * M].view(torch.float32).view(8, M)
        own.append(torch.stack([plane(states[r], 2)[lo:hi], plane(states[r], 3)[lo:hi], a8[:, lo:hi].sum(dim=0)]))
    full = torch.cat(own, dim=1)                                                           # (3, M)
    go = torch.full((), 0.5, device="cuda")
    grads = []
    for r in range(W):
        lo, hi = rows[r]
        plane(states[r], 2).copy_(full[0])
        plane(states[r], 3).copy_(full[1])
        states[r][off[4]:off[4] + 4 * M].view(torch.float32).copy_(full[2])
        states[r][off[0]:off[0] + 8].view(torch.float32)[1] = 1.0
        g = torch.empty(Bl, N, D, device="cuda")
        _lib.check(L.dycon_fecl_gn_bwd(P(states[r]), sbytes, P(la), B_all, N, D, 1, 1 / 0.6, 2.0, 1, 0, thr, 1.0, prec,
                                       lo, hi, ctypes.c_void_p(total.data_ptr() + 16), P(go), P(g), *g.stride(),
                                       stream), "gn_bwd")
        grads.append(g)
    ref = oracle(f, mask, t, 100, 0.5)
    assert abs(loss - ref["loss"]) <= 2e-3 * abs(ref["loss"]), (loss, ref["loss"])
    got = torch.cat(grads).cpu().numpy().reshape(ref["grad"].shape)
    strict(got, ref, f, t)
