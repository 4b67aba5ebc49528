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


@settings(max_examples=15, deadline=None)
@given(n=st.integers(3, 70), d=st.integers(2, 12), p_fg=st.sampled_from([0.0, 0.2, 0.5]), focal=st.booleans(),
       teacher=st.booleans(), weight=st.booleans(), block=st.sampled_from([1, 7, 16, 64]), seed=st.integers(0, 10_000))
def test_blocked_closed_form_equals_the_dense_one(n, d, p_fg, focal, teacher, weight, block, seed):
    """fecl_blocked (row-block sweeps, the transposed term from per-row statistics: what the GPU tests use at
    N = 9216 and for merged batches) reproduces closed_form.fecl, which is pinned to the reference."""
    rng = np.random.default_rng(seed)
    y = (rng.random((1, n)) < p_fg).astype(np.float32)
    y[0, 0] = 1.0 - y[0, 1]
    f = rng.standard_normal((1, n, d)) + 0.7
    f /= np.linalg.norm(f, axis=-1, keepdims=True)
    t = None
    if teacher:
        t = f + 0.2 * rng.standard_normal((1, n, d))
        t /= np.linalg.norm(t, axis=-1, keepdims=True)
    w = rng.random((1, n)) if weight else None
    kw = dict(inv_tau=1 / 0.6, gamma=2.0, use_focal=focal, cross_thresh=0.31, lambda_cross=0.7, go=0.5, rows_global=3 * n)
    a = closed_form.fecl(f, y, t, w, cnt_global=123.0 if teacher else None, **kw)
    b = closed_form.fecl_blocked(f, y, t, w, cnt_global=123.0 if teacher else None, block=block, **kw)
    if not np.isfinite(a["loss"]):
        return
    assert abs(a["loss"] - b["loss"]) <= 1e-12 * max(1.0, abs(a["loss"]))
    scale = max(np.abs(a["grad"]).max(), 1e-30)
    assert np.abs(a["grad"][0] - b["grad"]).max() <= 1e-10 * scale
    lo, hi = n // 3, max(n // 3 + 1, 2 * n // 3)
    c = closed_form.fecl_blocked(f, y, t, w, cnt_global=123.0 if teacher else None, block=block, grad_rows=(lo, hi), **kw)
    assert np.abs(a["grad"][0][lo:hi] - c["grad"]).max() <= 1e-10 * scale


def test_strict_comparator_predicts_flips_from_rounded_operands():
    """fecl_grad_error_strict on a synthetic 'kernel': the closed form evaluated on fp16-rounded operands stands in
    for the GPU.  Its threshold flips must be predicted (err small), and a gradient with a WRONG membership for one
    boundary pair must be rejected -- the fitted comparator (fecl_grad_error) would have accepted it."""
    rng = np.random.default_rng(5)
    b, n, d = 1, 160, 16
    y = (rng.random((b, n)) < 0.4).astype(np.float32)
    f = rng.standard_normal((b, n, d)).astype(np.float32) + 0.9
    f /= np.linalg.norm(f, axis=-1, keepdims=True)
    t = f + 0.25 * rng.standard_normal((b, n, d)).astype(np.float32)
    t /= np.linalg.norm(t, axis=-1, keepdims=True)
    f, t = f.astype(np.float32), t.astype(np.float32)
    cs = np.einsum("bid,bjd->bij", f.astype(np.float64), t.astype(np.float64))
    thr = float(np.median(cs))                       # many pairs sit near the threshold
    kw = dict(inv_tau=1 / 0.6, gamma=2.0, use_focal=True, cross_thresh=thr, go=0.5)
    ref = closed_form.fecl(f, y, t, None, ambiguity=5e-4, **kw)
    fq, tq = closed_form.round_operand(f, "fp16"), closed_form.round_operand(t, "fp16")
    kern = closed_form.fecl(fq, y, tq, None, **kw)          # what an exact fp16-operand kernel returns
    st = closed_form.fecl_grad_error_strict(kern["grad"], ref, f, t, "fp16", thr)
    assert st["outside_flips"] == 0 and st["flipped"] > 0
    assert st["err"] <= 2e-3, st
    # flip ONE boundary pair the wrong way in the 'kernel' output: strict must see it
    bb, ii, jj, c, hard = next(a for a in ref["ambiguous"] if abs(float(fq[a[0], a[1]] @ tq[a[0], a[2]]) - thr) > 1e-5)
    wrong = kern["grad"].copy()
    sign = -1.0 if (float(fq[bb, ii] @ tq[bb, jj]) > thr) else 1.0
    wrong[bb, ii] += sign * 0.5 * t[bb, jj].astype(np.float64) / ((1.0 - c) * kern["cnt"])
    bad = closed_form.fecl_grad_error_strict(wrong, ref, f, t, "fp16", thr)
    assert bad["err"] > 5 * st["err"] and bad["err"] > 2e-3, (bad, st)
