"""NumPy float64 closed forms + hand-derived gradients -- TEST INFRASTRUCTURE, NOT PRODUCT.

Independent of ``oracle.torch_port`` (no autograd): this is the specification the
CUDA kernels implement.  Derivations are in SURVEY.md section 0 / DESIGN.md; each
function cites the reference lines whose value it must reproduce.  Pinned against
outputs of the unmodified reference by ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import numpy as np

EPS_UNCL = 1e-6     # dycon_losses.py:95
EPS_FECL = 1e-18    # dycon_losses.py:187,189,192,229


# --------------------------------------------------------------------------- UnCL
def _softmax_c(x):
    x = x - x.max(axis=1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=1, keepdims=True)


def uncl(s, t, beta, go=1.0, count=None):
    """UnCL value and d(go*loss)/ds  (dycon_losses.py:94-118 and its autograd backward).

    s, t: (B, C, ...) arrays.  ``count`` overrides the mean denominator B*V (used by
    the sharded path, where it is the *global* voxel count).
    Returns dict(loss, grad, sum) with ``sum`` = sum_v L_v.
    """
    s = np.asarray(s, np.float64)
    t = np.asarray(t, np.float64)
    b, c = s.shape[:2]
    s2 = s.reshape(b, c, -1)
    t2 = t.reshape(b, c, -1)
    cnt = float(count if count is not None else b * s2.shape[2])
    ps, pt = _softmax_c(s2), _softmax_c(t2)
    ls, lt = np.log(ps + EPS_UNCL), np.log(pt + EPS_UNCL)
    hs, ht = -(ps * ls).sum(1), -(pt * lt).sum(1)              # :99-105
    es, et = np.exp(beta * hs), np.exp(beta * ht)              # :108-109
    w = 1.0 / (es + et)
    q = ((ps - pt) ** 2).sum(1)
    per_voxel = q * w + beta * (hs + ht)                       # :113,:116 (mean of the broadcast)
    total = per_voxel.sum()
    # backward
    dl_dhs = beta - q * beta * es * w * w
    g = 2.0 * (ps - pt) * w[:, None] - dl_dhs[:, None] * (ls + ps / (ps + EPS_UNCL))
    grad = (go / cnt) * ps * (g - (ps * g).sum(1, keepdims=True))
    return {"loss": total / cnt, "sum": total, "grad": grad.reshape(s.shape)}


# --------------------------------------------------------------------------- FeCL
def fecl(feat, mask, teacher=None, row_weight=None, *, inv_tau, gamma=2.0, use_focal=False,
         cross_thresh=0.5, lambda_cross=1.0, go=1.0, rows_global=None, cnt_global=None, ambiguity=0.0):
    """FeCL value, gradient and the per-row statistics the kernels keep.

    feat/teacher (B,N,D), mask (B,N) labels, row_weight (B,N)|None (the reference's
    ``gambling_uncertainty``, which also switches focal weighting off --
    dycon_losses.py:209-211).  ``cross_thresh`` is the host scalar
    ``sigmoid_rampup(epoch, rampup, 0.3, 0.5)`` (dycon_losses.py:222).
    ``rows_global`` / ``cnt_global`` override the two batch-global normalisers
    (B*N of the mean at :193/:206/:211 and the hard-negative count at :229) for
    the sharded path.

    ``ambiguity`` > 0 additionally reports the negative pairs whose cross similarity lies within
    that distance of ``cross_thresh``: membership of the hard set is a step function of ``cs``
    (dycon_losses.py:223), so an implementation that rounds ``cs`` differently may legitimately
    flip exactly those pairs (see ``fecl_grad_error``).
    """
    f = np.asarray(feat, np.float64)
    b, n, _ = f.shape
    y = np.asarray(mask).reshape(b, n)
    rows = float(rows_global if rows_global is not None else b * n)
    focal = bool(use_focal) and row_weight is None
    r = np.ones((b, n)) if row_weight is None else np.asarray(row_weight, np.float64).reshape(b, n)

    pos = (y[:, :, None] == y[:, None, :])                     # :172
    neg = ~pos
    offd = ~np.eye(n, dtype=bool)[None]
    lg = np.einsum("bid,bjd->bij", f, f) * inv_tau             # :175
    lg = np.where(offd, lg, 0.0)                               # :178
    m = lg.max(axis=1)                                         # :180  (per column j)
    e = np.exp(lg - m[:, None, :])                             # :181-183
    nsum = (e * neg).sum(-1)                                   # :184
    tt = e + nsum[:, :, None]
    d = e / (tt + EPS_FECL)                                    # :186-187
    pm = pos & offd
    nl = -np.log(d + EPS_FECL)                                 # :189
    wgt = (1.0 - d) ** gamma if focal else np.ones_like(d)     # :201-202 (positives only)
    c = 1.0 / (pos.sum(-1) - 1 + EPS_FECL)                     # :192
    row_loss = c * (nl * wgt * pm).sum(-1)
    student_sum = (r * row_loss).sum()

    kappa = r * c / rows
    with np.errstate(divide="ignore", invalid="ignore"):
        if focal:
            dphi = gamma * (1.0 - d) ** (gamma - 1.0) * np.log(d + EPS_FECL) - (1.0 - d) ** gamma / (d + EPS_FECL)
        else:
            dphi = -1.0 / (d + EPS_FECL)
    dphi = np.where(pm, dphi, 0.0)
    a = (dphi * d / (tt + EPS_FECL)).sum(-1)
    dl = kappa[:, :, None] * (dphi * d * (1.0 - d) - neg * e * a[:, :, None])
    g = np.where(offd, dl, 0.0) * inv_tau
    grad = np.einsum("bij,bjd->bid", g + g.transpose(0, 2, 1), f)
    grad_student = grad
    cross_unnorm, ambiguous = None, []

    cross_sum, cnt = 0.0, 0.0
    if teacher is not None:
        tf = np.asarray(teacher, np.float64)
        cs = np.einsum("bid,bjd->bij", f, tf)                  # :217
        hard = neg & (cs > cross_thresh)                       # :223
        cnt = float(hard.sum())
        with np.errstate(divide="ignore", invalid="ignore"):
            cross_sum = float((-np.log(1.0 - cs + EPS_FECL))[hard].sum()) if cnt else 0.0
        cg = float(cnt_global if cnt_global is not None else cnt)
        with np.errstate(divide="ignore", invalid="ignore"):
            gc = np.where(hard, 1.0 / ((1.0 - cs + EPS_FECL) * (cg + EPS_FECL)), 0.0)
        grad = grad + lambda_cross * np.einsum("bij,bjd->bid", gc, tf)
        if ambiguity > 0:
            with np.errstate(divide="ignore", invalid="ignore"):
                cross_unnorm = np.einsum("bij,bjd->bid", np.where(hard, 1.0 / (1.0 - cs + EPS_FECL), 0.0), tf)
            for bb, ii, jj in zip(*np.nonzero(neg & (np.abs(cs - cross_thresh) <= ambiguity))):
                ambiguous.append((int(bb), int(ii), int(jj), float(cs[bb, ii, jj]), bool(hard[bb, ii, jj])))
    else:
        cg = 0.0

    loss = student_sum / rows + lambda_cross * (cross_sum / (cg + EPS_FECL) if teacher is not None else 0.0)
    return {"loss": loss, "grad": go * grad, "m": m, "n": nsum, "A": a, "kappa": kappa,
            "student_sum": student_sum, "cross_sum": cross_sum, "cnt": cnt,
            "grad_student": go * grad_student, "cross_unnorm": cross_unnorm, "ambiguous": ambiguous,
            "go": go, "lambda_cross": lambda_cross}


def fecl_grad_error(grad, ref, teacher, max_per_row=14):
    """max|grad - oracle| / max|oracle|, minimised over the admissible states of the threshold-boundary
    pairs listed in ``ref["ambiguous"]`` (needs ``fecl(..., ambiguity=delta)``).

    The cross gradient is (lambda/(cnt+1e-18)) * sum_hard t_j/(1-cs_ij): a boundary pair (i,j) only moves
    row i (plus the global count), so the admissible state is chosen row by row from the cached
    un-normalised sum (all 2^k states of the k boundary pairs of a row are evaluated at once); the
    oracle is not re-run.  Rows with more than ``max_per_row`` boundary pairs keep the fp64 membership.
    """
    g = np.asarray(grad, np.float64)
    scale = np.abs(ref["grad"]).max()

    def err(x):
        return float(np.abs(g - x).max() / scale)

    amb = ref["ambiguous"]
    plain = err(ref["grad"])
    if not amb or ref["cross_unnorm"] is None:
        return plain
    tf = np.asarray(teacher, np.float64)
    k = ref["go"] * ref["lambda_cross"]
    rows = {}
    for b, i, j, cs, hard in amb:
        rows.setdefault((b, i), []).append((j, cs, hard))
    u = ref["cross_unnorm"].copy()
    cnt = ref["cnt"]
    net = 0.0
    for (b, i), pairs in rows.items():
        if len(pairs) > max_per_row:
            continue      # too many to enumerate: keep the fp64 membership (each flip only weighs 1/cnt)
        signs = np.array([-1.0 if hard else 1.0 for _, _, hard in pairs])
        contrib = np.stack([sg * tf[b, j] / (1.0 - cs + EPS_FECL) for sg, (j, cs, _) in zip(signs, pairs)])
        bits = ((np.arange(1 << len(pairs))[:, None] >> np.arange(len(pairs))[None, :]) & 1).astype(np.float64)
        cand = ref["grad_student"][b, i][None, :] + k * (u[b, i][None, :] + bits @ contrib) / (cnt + EPS_FECL)
        best = int(np.abs(cand - g[b, i][None, :]).max(axis=1).argmin())
        u[b, i] = u[b, i] + bits[best] @ contrib
        net += float(bits[best] @ signs)
    return min(plain, err(ref["grad_student"] + k * u / (cnt + net + EPS_FECL)))


# --------------------------------------------------------------------------- EMA
def ema(ema_params, params, alpha, global_step):
    """Out-of-place fp32 EMA with the reference's rounding order
    (train_DyCON_BraTS19.py:157,164): t = rn(ema*a); out = t + rn(1-a)*p."""
    a = min(1 - 1 / (global_step + 1), alpha)
    a32, oma32 = np.float32(a), np.float32(1 - a)
    out = []
    for e, p in zip(ema_params, params):
        t = (np.asarray(e, np.float32) * a32).astype(np.float32)
        out.append((t + oma32 * np.asarray(p, np.float32)).astype(np.float32))
    return out
