"""Property tests (hypothesis, CPU): the two independent restatements of the reference -- the op-for-op torch
port with autograd gradients and the float64 closed forms with hand-derived gradients -- agree on random
shapes, label patterns and options, including the edge cases of SURVEY.md section 8(d): rows whose class has a
single member (P_i = 1), an all-background sample, no hard negatives (cnt = 0), a row weight, C > 2."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import closed_form, torch_port


@settings(max_examples=25, deadline=None)
@given(b=st.integers(1, 3), n=st.integers(2, 24), d=st.integers(2, 12), p_fg=st.sampled_from([0.0, 0.1, 0.5]),
       focal=st.booleans(), teacher=st.booleans(), weight=st.booleans(), epoch=st.sampled_from([0, 100, 1500]),
       seed=st.integers(0, 10_000))
def test_fecl_closed_form_equals_autograd_of_the_port(b, n, d, p_fg, focal, teacher, weight, epoch, seed):
    g = torch.Generator().manual_seed(seed)
    mask = (torch.rand(b, 1, n, generator=g) < p_fg).double()
    if n > 2:
        mask[0, 0, 0] = 1.0 - mask[0, 0, 1]            # make sure some pair of different labels exists
    f = torch.nn.functional.normalize(torch.randn(b, n, d, generator=g, dtype=torch.float64) + 0.7, dim=-1)
    t = torch.nn.functional.normalize(f + 0.2 * torch.randn(b, n, d, generator=g, dtype=torch.float64), dim=-1) if teacher else None
    w = torch.rand(b, n, generator=g, dtype=torch.float64) if weight else None
    kw = dict(temperature=0.6, gamma=2.0, use_focal=focal, rampup_epochs=1500, lambda_cross=0.7)
    loss, grad = torch_port.fecl_fwd_bwd(f, mask, t, w, epoch, go=0.5, dtype=torch.float64, **kw)
    if not torch.isfinite(loss):                        # cs rounded above 1: the reference is NaN there too
        return
    thr = torch_port.ramp_threshold(epoch, 1500, 0.3, 0.5)
    ref = closed_form.fecl(f.numpy(), mask.numpy(), None if t is None else t.numpy(), None if w is None else w.numpy(),
                           inv_tau=1 / 0.6, gamma=2.0, use_focal=focal, cross_thresh=thr, lambda_cross=0.7, go=0.5)
    assert abs(ref["loss"] - loss.item()) <= 1e-10 * max(1.0, abs(loss.item()))
    scale = max(np.abs(grad.numpy()).max(), 1e-30)
    assert np.abs(ref["grad"] - grad.numpy()).max() <= 1e-9 * scale + 1e-15


@settings(max_examples=25, deadline=None)
@given(b=st.integers(1, 3), c=st.integers(2, 4), v=st.integers(1, 30), beta=st.sampled_from([0.5, 1.58, 5.0]),
       scale=st.sampled_from([0.5, 2.0, 8.0]), seed=st.integers(0, 10_000))
def test_uncl_closed_form_equals_autograd_of_the_port(b, c, v, beta, scale, seed):
    g = torch.Generator().manual_seed(seed)
    s = scale * torch.randn(b, c, v, 1, 1, generator=g, dtype=torch.float64)
    t = s + 0.5 * torch.randn(b, c, v, 1, 1, generator=g, dtype=torch.float64)
    loss, grad = torch_port.uncl_fwd_bwd(s, t, beta, go=0.5, dtype=torch.float64)
    ref = closed_form.uncl(s.numpy(), t.numpy(), beta, go=0.5)
    assert abs(ref["loss"] - loss.item()) <= 1e-12 * max(1.0, abs(loss.item()))
    # (confident voxels have gradients ~1e-6 that come from cancellation: absolute floor of a few float64 ulps of 1)
    assert np.abs(ref["grad"] - grad.numpy()).max() <= 1e-9 * np.abs(grad.numpy()).max() + 1e-14
    # softmax Jacobian: the channel gradients of a voxel sum to zero
    assert np.abs(ref["grad"].sum(axis=1)).max() <= 1e-9 * np.abs(ref["grad"]).max() + 1e-14
