"""Golden vectors for the step-loop rows of SURVEY.md section 8(f) -- TEST INFRASTRUCTURE.

    python -m oracle.make_golden_step          (build container only: needs /root/reference)

Executes the UNMODIFIED reference functions -- ``UnCLoss`` (code/utils/dycon_losses.py), ``dice_loss`` and
``softmax_mse_loss`` (code/utils/losses.py), ``F.cross_entropy`` -- exactly as the step loop combines them
(code/train_DyCON_BraTS19.py:308-314,351-357) on small seeded inputs in fp32 and fp64, and writes inputs, the four
losses and the gradient of their weighted sum to tests/golden/step.npz.  Kept apart from oracle/make_golden.py so
that the round-1 fixtures are never rewritten.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
WEIGHTS = (0.5, 1.0, 1.0, 0.37)        # u_weight, l_weight (ce), l_weight (dice), a consistency weight


def step_case(ref, stock, shape, labeled_bs, beta, scale, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    s32 = scale * torch.randn(shape, generator=g)
    t = (s32 + 0.5 * torch.randn(shape, generator=g)).to(dtype)       # inputs are fp32 values in both runs
    s = s32.to(dtype).requires_grad_(True)
    label = (torch.rand((shape[0],) + shape[2:], generator=g) < 0.3).long()
    probs, eprobs = F.softmax(s, dim=1), F.softmax(t, dim=1)                       # train_DyCON_BraTS19.py:308-309
    ce = F.cross_entropy(s[:labeled_bs], label[:labeled_bs])                       # :313
    dice = stock.dice_loss(probs[:labeled_bs, 1], label[:labeled_bs] == 1)         # :314
    u = ref.UnCLoss()(s, t, beta)                                                  # :351
    cons = stock.softmax_mse_loss(probs[labeled_bs:], eprobs[labeled_bs:]).mean()  # :352
    total = WEIGHTS[0] * u + WEIGHTS[1] * ce + WEIGHTS[2] * dice + WEIGHTS[3] * cons
    total.backward()
    return dict(s=s.detach().float().numpy(), t=t.float().numpy(), label=label.numpy(),
                losses=np.array([u.item(), ce.item(), dice.item(), cons.item()], np.float64), grad=s.grad.numpy())


def prep_case(ref, feat_shape, label_shape, seed, dtype, epoch=100, teacher=True):
    """The caller-side preparation of code/train_DyCON_BraTS19.py:316-330 (restated line for line: the script itself
    cannot be imported) in front of the UNMODIFIED FeCLoss."""
    g = torch.Generator().manual_seed(seed)
    b, c = feat_shape[:2]
    base = torch.randn(1, c, 1, 1, 1, generator=g)
    x32 = base + 0.8 * torch.randn(feat_shape, generator=g)
    label = torch.zeros(label_shape, dtype=torch.long)
    k = tuple(le // fe for le, fe in zip(label_shape[1:], feat_shape[2:]))
    coarse = torch.rand((b,) + tuple(feat_shape[2:]), generator=g) < 0.35          # blocky labels, some cells half full
    label = coarse.repeat_interleave(k[0], 1).repeat_interleave(k[1], 2).repeat_interleave(k[2], 3).long()
    label = label * (torch.rand(label_shape, generator=g) < 0.8).long()
    x32 = x32 + 0.5 * coarse.float().unsqueeze(1)
    t32 = x32 + 0.1 * torch.randn(feat_shape, generator=g)
    stud_features = x32.to(dtype).clone().requires_grad_(True)
    ema_features = t32.to(dtype)
    stud_embedding = stud_features.view(b, c, -1)                                  # :317
    stud_embedding = torch.transpose(stud_embedding, 1, 2)                         # :318
    stud_embedding = F.normalize(stud_embedding, dim=-1)                           # :319
    ema_embedding = F.normalize(torch.transpose(ema_features.view(b, c, -1), 1, 2), dim=-1)       # :321-323
    mask_con = F.avg_pool3d(label.float(), kernel_size=k, stride=k)                # :326-327
    mask_con = (mask_con > 0.5).float().reshape(b, -1).unsqueeze(1).to(dtype)      # :328-330
    torch.set_default_dtype(dtype)
    try:
        crit = ref.FeCLoss(device="cpu", temperature=0.6, gamma=2.0, use_focal=True, rampup_epochs=1500)
        loss = crit(feat=stud_embedding, mask=mask_con, teacher_feat=ema_embedding if teacher else None,
                    gambling_uncertainty=None, epoch=epoch)                        # :346-350
    finally:
        torch.set_default_dtype(torch.float32)
    (0.5 * loss).backward()
    return dict(features=x32.numpy(), ema_features=t32.numpy(), label=label.numpy(), mask=mask_con.float().numpy(),
                loss=np.float64(loss.item()), grad=stud_features.grad.numpy())


def main():
    ref, stock = ref_loader.dycon_losses(), ref_loader.stock_losses()
    prep = {}
    for name, (fs, ls, seed, teacher) in {"cubic_k4": ((2, 16, 3, 3, 2), (2, 12, 12, 8), 21, True),
                                          "anisotropic": ((2, 8, 4, 2, 3), (2, 8, 6, 12), 22, True),
                                          "no_teacher_k2": ((3, 12, 2, 3, 4), (3, 4, 6, 8), 23, False)}.items():
        r32 = prep_case(ref, fs, ls, seed, torch.float32, teacher=teacher)
        r64 = prep_case(ref, fs, ls, seed, torch.float64, teacher=teacher)
        rec = dict(features=r32["features"], ema_features=r32["ema_features"], label=r32["label"], mask=r32["mask"],
                   teacher=np.int64(teacher), epoch=np.int64(100), go=np.float64(0.5), loss32=r32["loss"], grad32=r32["grad"],
                   loss64=r64["loss"], grad64=r64["grad"])
        prep.update({f"{name}/{k}": v for k, v in rec.items()})
    np.savez_compressed(os.path.join(OUT, "prep.npz"), **prep)
    print("wrote prep.npz")
    specs = {
        "half_labelled": ((4, 2, 6, 5, 4), 2, 1.58, 2.0, 11),
        "three_of_four": ((4, 2, 5, 4, 3), 3, 5.0, 2.0, 12),
        "one_labelled": ((3, 2, 4, 4, 4), 1, 0.5, 2.0, 13),
        "confident": ((4, 2, 3, 7, 5), 2, 1.58, 9.0, 14),
    }
    flat = {}
    for name, (shape, lb, beta, scale, seed) in specs.items():
        r32 = step_case(ref, stock, shape, lb, beta, scale, seed, torch.float32)
        r64 = step_case(ref, stock, shape, lb, beta, scale, seed, torch.float64)
        assert np.array_equal(r32["label"], r64["label"])
        rec = dict(weights=np.array(WEIGHTS, np.float64), s=r32["s"], t=r32["t"], label=r32["label"], labeled_bs=np.int64(lb), beta=np.float64(beta),
                   losses32=r32["losses"], grad32=r32["grad"], losses64=r64["losses"], grad64=r64["grad"])
        flat.update({f"{name}/{k}": v for k, v in rec.items()})
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "step.npz"), **flat)
    print("wrote step.npz", sorted(specs))


if __name__ == "__main__":
    main()
