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
            "go": go, "lambda_cross": lambda_cross, "labels": y}


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


def fecl_blocked(feat, mask, teacher=None, row_weight=None, *, inv_tau, gamma=2.0, use_focal=False,
                 cross_thresh=0.5, lambda_cross=1.0, go=1.0, rows_global=None, cnt_global=None, block=1024,
                 grad_rows=None, ambiguity=0.0):
    """The same value and gradient as ``fecl`` for ONE sample (feat (N,D) or (1,N,D)) without ever holding an
    (N,N) array: four sweeps over row blocks (row max; negative sums; loss and A; gradient), the transposed
    gradient term G_ji evaluated from the per-row statistics of row j -- the structure the kernels use.  For
    shapes where ``fecl`` does not fit (ISLES22 N=9216, merged batches of config 5).  ``grad_rows`` = (lo, hi)
    restricts the gradient sweep to those rows (the loss always covers all rows).  float64 throughout, on torch
    CPU tensors so that the element-wise passes use all host cores.  ``ambiguity`` > 0 also returns what
    ``fecl_grad_error_strict`` needs (boundary pairs, un-normalised cross sums, the student part), shaped for a
    batch of one.  dycon_losses.py:150-235."""
    import torch
    f = torch.as_tensor(np.asarray(feat, np.float64)).reshape(-1, np.asarray(feat).shape[-1])
    n = f.shape[0]
    y = torch.as_tensor(np.asarray(mask, np.float64)).reshape(n)
    tf = None if teacher is None else torch.as_tensor(np.asarray(teacher, np.float64)).reshape(n, -1)
    r = torch.ones(n, dtype=torch.float64) if row_weight is None else torch.as_tensor(np.asarray(row_weight, np.float64)).reshape(n)
    rows = float(rows_global if rows_global is not None else n)
    focal = bool(use_focal) and row_weight is None
    blocks = [(a, min(a + block, n)) for a in range(0, n, block)]
    idx = torch.arange(n)

    def logits(a, b):                                          # rows a..b x all columns, diagonal zeroed (:175-178)
        lg = (f[a:b] @ f.T) * inv_tau
        lg[idx[a:b] - a, idx[a:b]] = 0.0
        return lg

    def dphi_of(d):
        if focal:
            return gamma * (1.0 - d) ** (gamma - 1.0) * torch.log(d + EPS_FECL) - (1.0 - d) ** gamma / (d + EPS_FECL)
        return -1.0 / (d + EPS_FECL)

    m = torch.empty(n, dtype=torch.float64)
    for a, b in blocks:                                        # :180 column max == row max (the matrix is symmetric)
        m[a:b] = logits(a, b).max(dim=1).values
    nsum, pcount = torch.empty(n, dtype=torch.float64), torch.empty(n, dtype=torch.float64)
    for a, b in blocks:                                        # :183-184
        pos = y[a:b, None] == y[None, :]
        e = torch.exp(logits(a, b) - m[None, :])
        nsum[a:b] = (e * ~pos).sum(-1)
        pcount[a:b] = pos.sum(-1).double()
    c = 1.0 / (pcount - 1 + EPS_FECL)                          # :192
    kappa = r * c / rows
    amat = torch.empty(n, dtype=torch.float64)
    student_sum, cross_sum, cnt = 0.0, 0.0, 0.0

    def pair_terms(a, b):
        pos = y[a:b, None] == y[None, :]
        pm = pos.clone()
        pm[idx[a:b] - a, idx[a:b]] = False
        e = torch.exp(logits(a, b) - m[None, :])
        tt = e + nsum[a:b, None]
        d = e / (tt + EPS_FECL)
        dphi = torch.where(pm, dphi_of(d), torch.zeros((), dtype=torch.float64))
        return pos, pm, e, tt, d, dphi

    for a, b in blocks:                                        # :186-206 and the backward row scalar A
        pos, pm, e, tt, d, dphi = pair_terms(a, b)
        nl = -torch.log(d + EPS_FECL)
        wgt = (1.0 - d) ** gamma if focal else 1.0
        student_sum += float((r[a:b] * c[a:b] * (nl * wgt * pm).sum(-1)).sum())
        amat[a:b] = (dphi * d / (tt + EPS_FECL)).sum(-1)
        if tf is not None:                                     # :217-229
            cs = f[a:b] @ tf.T
            hard = ~pos & (cs > cross_thresh)
            cnt += float(hard.sum())
            cross_sum += float((-torch.log(1.0 - cs + EPS_FECL))[hard].sum())
    cg = float(cnt_global if cnt_global is not None else cnt) if tf is not None else 0.0
    loss = student_sum / rows + lambda_cross * (cross_sum / (cg + EPS_FECL) if tf is not None else 0.0)

    lo, hi = grad_rows if grad_rows is not None else (0, n)
    grad = torch.zeros((hi - lo, f.shape[1]), dtype=torch.float64)
    grad_student = torch.zeros_like(grad) if ambiguity > 0 else None
    cross_unnorm = torch.zeros_like(grad) if (ambiguity > 0 and tf is not None) else None
    ambiguous = []
    zero = torch.zeros((), dtype=torch.float64)
    for a, b in [(max(a, lo), min(b, hi)) for a, b in blocks if min(b, hi) > max(a, lo)]:
        pos, pm, e, tt, d, dphi = pair_terms(a, b)
        gij = kappa[a:b, None] * (dphi * d * (1.0 - d) - ~pos * e * amat[a:b, None])
        # G_ji for the same (i, j): row j's statistics, column max m_i;  l_ji = l_ij
        eT = torch.exp(logits(a, b) - m[a:b, None])
        ttT = eT + nsum[None, :]
        dT = eT / (ttT + EPS_FECL)
        dphiT = torch.where(pm, dphi_of(dT), zero)
        gji = kappa[None, :] * (dphiT * dT * (1.0 - dT) - ~pos * eT * amat[None, :])
        h = (gij + gji) * inv_tau
        h[idx[a:b] - a, idx[a:b]] = 0.0
        g = h @ f
        if grad_student is not None:
            grad_student[a - lo:b - lo] = g
        if tf is not None:
            cs = f[a:b] @ tf.T
            hard = ~pos & (cs > cross_thresh)
            gc = torch.where(hard, 1.0 / ((1.0 - cs + EPS_FECL) * (cg + EPS_FECL)), zero)
            g = g + lambda_cross * (gc @ tf)
            if cross_unnorm is not None:
                cross_unnorm[a - lo:b - lo] = torch.where(hard, 1.0 / (1.0 - cs + EPS_FECL), zero) @ tf
                near = ~pos & ((cs - cross_thresh).abs() <= ambiguity)
                for ii, jj in zip(*torch.nonzero(near, as_tuple=True)):
                    ambiguous.append((0, int(ii) + a - lo, int(jj), float(cs[ii, jj]), bool(hard[ii, jj])))
        grad[a - lo:b - lo] = g
    out = {"loss": loss, "grad": (go * grad).numpy(), "m": m.numpy(), "n": nsum.numpy(), "A": amat.numpy(),
           "kappa": kappa.numpy(), "student_sum": student_sum, "cross_sum": cross_sum, "cnt": cnt}
    if ambiguity > 0:
        out.update(grad=out["grad"][None], grad_student=(go * grad_student).numpy()[None],
                   cross_unnorm=None if cross_unnorm is None else cross_unnorm.numpy()[None], ambiguous=ambiguous,
                   go=go, lambda_cross=lambda_cross, labels=y.numpy()[None, lo:hi])
    return out


def round_operand(x, mode):
    """fp32 array -> the values the 16-bit tensor-core modes feed to the MMA (round to nearest even)."""
    x = np.asarray(x, np.float32)
    if mode == "fp16":
        return x.astype(np.float16).astype(np.float64)
    if mode == "bf16":
        u = x.view(np.uint32).astype(np.uint64)
        u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
        return u.astype(np.uint32).view(np.float32).astype(np.float64)
    return x.astype(np.float64)


def fecl_grad_error_strict(grad, ref, feat, teacher, mode, cross_thresh, accum_window=3e-6, max_per_row=14):
    """The teacher-branch comparison WITHOUT fitting the flips to the output: hard-negative membership of the
    threshold-boundary pairs (``ref["ambiguous"]``) is PREDICTED from cs recomputed on the operands as the kernel
    rounds them (``round_operand``); only pairs whose rounded-operand cs lies within ``accum_window`` of the
    threshold (the fp32 accumulation order is the kernel's own) stay free.  Returns
    dict(err, plain, flipped, free, window_pairs, outside_flips):
      err            max|grad - predicted oracle| / max|oracle|
      plain          the same against the fp64 membership
      flipped        boundary pairs whose predicted membership differs from the fp64 one
      free           pairs left to the fit (inside the accumulation window)
      outside_flips  negative pairs OUTSIDE the ambiguity window whose rounded-operand membership differs from
                     fp64 (must be 0: it validates the window)."""
    g = np.asarray(grad, np.float64)
    scale = np.abs(ref["grad"]).max()
    err = lambda x: float(np.abs(g - x).max() / scale)
    plain = err(ref["grad"])
    out = {"err": plain, "plain": plain, "flipped": 0, "free": 0, "window_pairs": len(ref["ambiguous"]), "outside_flips": 0}
    if teacher is None or ref["cross_unnorm"] is None:
        return out
    fq, tq = round_operand(feat, mode), round_operand(teacher, mode)
    f64, t64 = np.asarray(feat, np.float64), np.asarray(teacher, np.float64)
    b, n, _ = f64.shape
    y = None
    amb = {(bb, ii, jj): (cs, hard) for bb, ii, jj, cs, hard in ref["ambiguous"]}
    # window validation: every negative pair outside the window keeps its membership under operand rounding
    labels = ref.get("labels")
    if labels is not None:
        for bb in range(b):
            csq = fq[bb] @ tq[bb].T
            cs64 = f64[bb] @ t64[bb].T
            neg = labels[bb][:, None] != labels[bb][None, :]
            diff = neg & ((csq > cross_thresh) != (cs64 > cross_thresh))
            for ii, jj in zip(*np.nonzero(diff)):
                if (bb, int(ii), int(jj)) not in amb:
                    out["outside_flips"] += 1
    k = ref["go"] * ref["lambda_cross"]
    u = ref["cross_unnorm"].copy()
    cnt, net = ref["cnt"], 0.0
    rows = {}
    for (bb, ii, jj), (cs, hard) in amb.items():
        rows.setdefault((bb, ii), []).append((jj, cs, hard))
    free_rows = {}
    for (bb, ii), pairs in rows.items():
        for jj, cs, hard in pairs:
            csq = float(fq[bb, ii] @ tq[bb, jj])
            if abs(csq - cross_thresh) <= accum_window:
                free_rows.setdefault((bb, ii), []).append((jj, cs, hard))
                out["free"] += 1
                continue
            want = csq > cross_thresh
            if want != hard:
                sg = 1.0 if want else -1.0
                u[bb, ii] += sg * t64[bb, jj] / (1.0 - cs + EPS_FECL)
                net += sg
                out["flipped"] += 1
    for (bb, ii), pairs in free_rows.items():
        if len(pairs) > max_per_row:
            continue
        signs = np.array([-1.0 if hard else 1.0 for _, _, hard in pairs])
        contrib = np.stack([sg * t64[bb, jj] / (1.0 - cs + EPS_FECL) for sg, (jj, cs, _) in zip(signs, pairs)])
        bits = ((np.arange(1 << len(pairs))[:, None] >> np.arange(len(pairs))[None, :]) & 1).astype(np.float64)
        cand = ref["grad_student"][bb, ii][None, :] + k * (u[bb, ii][None, :] + bits @ contrib) / (cnt + net + EPS_FECL)
        best = int(np.abs(cand - g[bb, ii][None, :]).max(axis=1).argmin())
        u[bb, ii] = u[bb, ii] + bits[best] @ contrib
        net += float(bits[best] @ signs)
    out["err"] = err(ref["grad_student"] + k * u / (cnt + net + EPS_FECL))
    return out


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
