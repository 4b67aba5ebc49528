"""Torch (CPU) restatement of the DyCON losses -- TEST INFRASTRUCTURE, NOT PRODUCT.

Operation-for-operation restatement of the reference's algorithm so that its
fp32 rounding behaviour and its CPU cost profile are representative; the
gradients come from autograd exactly as they do in the reference.  Pinned
against outputs of the unmodified reference by ``tests/test_oracle_golden.py``
(fixtures written by ``oracle/make_golden.py``).

Reference locations (rogeliorjr/DyCON_Paper_Replication):
  * UnCL            code/utils/dycon_losses.py:94-118
  * FeCL            code/utils/dycon_losses.py:150-235
  * sigmoid_rampup  code/utils/dycon_losses.py:28-47
  * adaptive_beta   code/utils/dycon_losses.py:8-12
  * EMA             code/train_DyCON_BraTS19.py:155-164
"""
from __future__ import annotations

import math

import torch


# --------------------------------------------------------------------------- host scalars
def beta_schedule(epoch, total_epochs, max_beta=5.0, min_beta=0.5):
    """Geometric decay max_beta -> min_beta (dycon_losses.py:8-12)."""
    return max_beta * (min_beta / max_beta) ** (epoch / total_epochs)


def ramp_threshold(epoch, ramp_epochs, lo, hi, steepness=5.0):
    """Gaussian-shaped ramp lo -> hi (dycon_losses.py:28-47)."""
    if ramp_epochs == 0:
        return hi
    e = max(0.0, min(float(epoch), ramp_epochs))
    return lo + (hi - lo) * math.exp(-steepness * (1.0 - e / ramp_epochs) ** 2)


# --------------------------------------------------------------------------- UnCL
def uncl_loss(s_logits: torch.Tensor, t_logits: torch.Tensor, beta: float) -> torch.Tensor:
    """Uncertainty-weighted consistency (dycon_losses.py:94-118).

    Keeps the reference's broadcasting: the class-summed term is (B,H,W,D) while
    the entropy regulariser is (B,1,H,W,D), so their sum is (B,B,H,W,D) before the
    mean (dycon_losses.py:116).
    """
    eps = 1e-6                                               # :95
    ps = torch.softmax(s_logits, dim=1)                      # :98
    hs = -(ps * torch.log(ps + eps)).sum(dim=1, keepdim=True)    # :99-100
    pt = torch.softmax(t_logits, dim=1)                      # :103
    ht = -(pt * torch.log(pt + eps)).sum(dim=1, keepdim=True)    # :104-105
    es = torch.exp(beta * hs)                                # :108
    et = torch.exp(beta * ht)                                # :109
    weighted = (ps - pt) ** 2 / (es + et)                    # :113
    return torch.mean(weighted.sum(dim=1) + beta * (hs + ht)).mean()   # :116-118


# --------------------------------------------------------------------------- FeCL
def fecl_loss(feat, mask, teacher_feat=None, gambling_uncertainty=None, epoch=0, *,
              temperature=0.6, gamma=2.0, use_focal=False, rampup_epochs=2000,
              lambda_cross=1.0) -> torch.Tensor:
    """Focal / teacher-augmented voxel-feature contrastive loss (dycon_losses.py:150-235).

    feat (B,N,D), mask (B,1,N), teacher_feat (B,N,D)|None, gambling_uncertainty (B,N)|None.
    """
    n = feat.shape[1]
    same = (mask == mask.transpose(1, 2)).to(feat.dtype)      # :172
    diff = 1 - same                                            # :173
    off_diag = 1 - torch.eye(n, dtype=feat.dtype, device=feat.device)   # :176-177

    logits = feat @ feat.transpose(1, 2) / temperature         # :175
    logits = logits * off_diag                                 # :178  (diag := 0, not -inf)
    col_max = logits.max(dim=1, keepdim=True).values           # :180
    logits = logits - col_max.detach()                         # :181
    ex = torch.exp(logits)                                     # :183
    neg_total = (ex * diff).sum(dim=-1)                        # :184
    ratio = ex / (ex + neg_total.unsqueeze(-1) + 1e-18)        # :186-187
    per_pair = -torch.log(ratio + 1e-18) * same * off_diag     # :189-190
    pos_count = same.sum(dim=-1) - 1 + 1e-18                   # :192
    student = (per_pair.sum(dim=-1) / pos_count).mean()        # :192-193

    if use_focal:                                              # :196-206
        w = torch.ones_like(ratio)
        pos_thr = ramp_threshold(epoch, rampup_epochs, 1.3, 1.5)
        neg_thr = ramp_threshold(epoch, rampup_epochs, 0.3, 0.5)
        hard_pos = same.bool() & (ratio < pos_thr)
        w[hard_pos] = (1 - ratio[hard_pos]).pow(gamma)
        hard_neg = diff.bool() & (ratio > neg_thr)
        w[hard_neg] = ratio[hard_neg].pow(gamma)
        student = ((per_pair * w).sum(dim=-1) / pos_count).mean()

    if gambling_uncertainty is not None:                       # :209-211
        student = (per_pair.sum(dim=-1) / pos_count * gambling_uncertainty).mean()

    cross = 0.0
    if teacher_feat is not None:                               # :213-231
        cs = feat @ teacher_feat.transpose(1, 2)               # :217 (no temperature)
        thr = ramp_threshold(epoch, rampup_epochs, 0.3, 0.5)   # :222
        hard = diff.bool() & (cs > thr)                        # :223
        if hard.sum() > 0:                                     # :226
            hardf = hard.to(feat.dtype)
            cross = (-torch.log(1 - cs + 1e-18) * hardf).sum() / (hardf.sum() + 1e-18)   # :227-229
    return student + lambda_cross * cross                      # :234


# --------------------------------------------------------------------------- EMA
def ema_update(ema_params, params, alpha, global_step):
    """Mean-teacher EMA over parameter lists, in place (train_DyCON_BraTS19.py:155-164)."""
    a = min(1 - 1 / (global_step + 1), alpha)
    with torch.no_grad():
        for e, p in zip(ema_params, params):
            e.mul_(a).add_(p, alpha=1 - a)


# --------------------------------------------------------------------------- helpers for tests/bench
def uncl_fwd_bwd(s, t, beta, go=1.0, dtype=None):
    """Returns (loss, grad_s) as detached tensors; ``go`` is the upstream scalar."""
    s = s.detach().to(dtype or s.dtype).clone().requires_grad_(True)
    t = t.detach().to(dtype or t.dtype)
    loss = uncl_loss(s, t, beta)
    (loss * go).backward()
    return loss.detach(), s.grad.detach()


def fecl_fwd_bwd(feat, mask, teacher=None, unc=None, epoch=0, go=1.0, dtype=None, **kw):
    """Returns (loss, grad_feat) as detached tensors."""
    dt = dtype or feat.dtype
    f = feat.detach().to(dt).clone().requires_grad_(True)
    loss = fecl_loss(f, mask.to(dt), None if teacher is None else teacher.detach().to(dt),
                     None if unc is None else unc.to(dt), epoch, **kw)
    (loss * go).backward()
    return loss.detach(), f.grad.detach()


# --------------------------------------------------------------------------- stock step-loop losses
def dice_loss(score, target):
    """Soft Dice on one class (code/utils/losses.py:8-16)."""
    target = target.to(score.dtype)
    smooth = 1e-5
    inter = torch.sum(score * target)
    return 1 - (2 * inter + smooth) / (torch.sum(score * score) + torch.sum(target * target) + smooth)


def softmax_mse(input_logits, target_logits):
    """Element-wise squared difference of the two softmaxes over dim 1 (code/utils/losses.py:65-82)."""
    return (torch.softmax(input_logits, dim=1) - torch.softmax(target_logits, dim=1)) ** 2


def step_losses(stud_logits, ema_logits, label_batch, labeled_bs, beta):
    """The four voxel-wise losses of one step, as the loop computes them (code/train_DyCON_BraTS19.py:308-314,
    351-352): (u_loss, loss_seg, loss_seg_dice, consistency_loss).  The consistency term is handed the
    PROBABILITIES (:352), so softmax_mse takes the softmax of a softmax -- restated as is."""
    stud_probs = torch.softmax(stud_logits, dim=1)                             # :308
    ema_probs = torch.softmax(ema_logits, dim=1)                               # :309
    loss_seg = torch.nn.functional.cross_entropy(stud_logits[:labeled_bs], label_batch[:labeled_bs])          # :313
    loss_seg_dice = dice_loss(stud_probs[:labeled_bs, 1], label_batch[:labeled_bs] == 1)                      # :314
    u_loss = uncl_loss(stud_logits, ema_logits, beta)                          # :351
    consistency = softmax_mse(stud_probs[labeled_bs:], ema_probs[labeled_bs:]).mean()                         # :352
    return u_loss, loss_seg, loss_seg_dice, consistency


# --------------------------------------------------------------------------- caller-side preparation
def prep_embeddings(features):
    """(B, C, h, w, d) -> unit rows (B, N, C) (code/train_DyCON_BraTS19.py:316-319)."""
    b, c = features.shape[:2]
    return torch.nn.functional.normalize(torch.transpose(features.reshape(b, c, -1), 1, 2), dim=-1)


def prep_mask(label_batch, feature_spatial):
    """(B, H, W, D) labels -> (B, 1, N) float mask (train_DyCON_BraTS19.py:326-330; per-axis kernels as in
    train_DyCON_ISLES22.py:268-281)."""
    k = tuple(ext // f for ext, f in zip(label_batch.shape[1:], feature_spatial))
    m = torch.nn.functional.avg_pool3d(label_batch.float(), kernel_size=k, stride=k)
    m = (m > 0.5).float()
    return m.reshape(label_batch.shape[0], -1).unsqueeze(1)


def fecl_from_features(stud_features, label_batch, ema_features=None, epoch=0, **ctor):
    """Preparation + FeCL, as the step loop chains them (train_DyCON_BraTS19.py:316-350)."""
    emb = prep_embeddings(stud_features)
    temb = None if ema_features is None else prep_embeddings(ema_features)
    mask = prep_mask(label_batch, stud_features.shape[2:]).to(emb.dtype)
    return fecl_loss(emb, mask, temb, None, epoch, **ctor)
