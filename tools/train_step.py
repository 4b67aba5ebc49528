#!/usr/bin/env python
"""BASELINE config 3: a full DyCON mean-teacher train step with the reference's own 3D U-Net as context.

    python tools/train_step.py [--shape pancreas] [--batch 8] [--steps 50] [--timed 20] [--json out.json]
    python bench.py --train-step ...                                  (same thing)

Three arms run the step loop of code/train_DyCON_Pancreas.py:198-272 (twin of train_DyCON_BraTS19.py:298-374) on
identical synthetic batches, from identical initial weights, with identical per-step seeds:

  reference  the UNMODIFIED reference loss modules (code/utils/dycon_losses.py, code/utils/losses.py from
             /root/reference or the git-ignored baseline/_ref), the per-tensor EMA loop, clip_grad_norm_ + SGD
  dropin     the two-line swap of INTEGRATION.md: this repo's UnCLoss / FeCLoss / update_ema_variables behind the
             same call sites, everything else PyTorch
  fused      every section-8(f) entry point as well: StepLosses (UnCL + CE + Dice + consistency in one pass),
             FeCLoss.from_features, loss_is_finite_flag + sgd_clip_ema_step (no host synchronisation in the step)

The model is the reference's UNet3D (code/networks/UNet3D_contrastive.py:207-316, `net_factory_3d("unet_3D",
in_chns=1, class_num=2, scaler=2)`) -- context, not product: its convolutions stay on PyTorch / cuDNN.  Output: the
loss trajectories of the three arms (must agree: the arms only differ in who computes the same numbers) and a
step-time breakdown from CUDA events.  TEST / BENCH INFRASTRUCTURE: imports the reference through oracle/ref_loader.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# args of the run scripts (code/run_Panc.sh, train_DyCON_Pancreas.py:24-69)
HP = dict(base_lr=0.01, ema_decay=0.99, l_weight=1.0, u_weight=0.5, temp=0.6, gamma=2.0, use_focal=1, feature_scaler=2,
          beta_min=0.5, beta_max=5.0, consistency=0.1, consistency_rampup=200.0, max_iterations=20000)


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="pancreas")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--labeled-bs", type=int, default=4)
    ap.add_argument("--steps", type=int, default=50, help="steps of the trajectory comparison")
    ap.add_argument("--timed", type=int, default=20, help="steps of the timed breakdown (after the trajectory)")
    ap.add_argument("--arms", default="reference,dropin,fused")
    ap.add_argument("--json", default=None)
    ap.add_argument("--patch", default=None, help="H,W,D override (e.g. 32,32,32 for a quick check)")
    return ap.parse_args(argv)


def consistency_weight(it):            # get_current_consistency_weight, train_DyCON_Pancreas.py:91-93 + ramps.sigmoid_rampup
    import numpy as np
    cur = np.clip(it // 150, 0.0, HP["consistency_rampup"])
    phase = 1.0 - cur / HP["consistency_rampup"]
    return HP["consistency"] * float(np.exp(-5.0 * phase * phase))


def make_batches(patch, batch, n, device):
    """n synthetic batches {image (B,1,H,W,D) fp32, label (B,H,W,D) int64}: a soft ellipsoid + noise."""
    import torch
    from dycon_paper_replication_b200.synthetic import make_blob_labels
    out = []
    for k in range(n):
        g = torch.Generator().manual_seed(1337 + k)
        label = make_blob_labels(batch, patch, g, lo=0.02, hi=0.10).long()
        image = 0.8 * label.float().unsqueeze(1) + 0.6 * torch.randn(batch, 1, *patch, generator=g)
        out.append((image.to(device), label.to(device)))
    return out


class Arm:
    def __init__(self, name, model, device, max_epoch):
        import torch
        from oracle import ref_loader
        self.name, self.device = name, device
        self.model = copy.deepcopy(model)
        self.ema = copy.deepcopy(model)
        for p in self.ema.parameters():
            p.detach_()                                     # create_model(ema=True), train_DyCON_Pancreas.py:128-133
        self.model.train()
        self.ema.train()
        self.opt = torch.optim.SGD(self.model.parameters(), lr=HP["base_lr"], momentum=0.9, weight_decay=0.0001)
        self.iter = 0
        self.max_epoch = max_epoch
        ctor = dict(temperature=HP["temp"], gamma=HP["gamma"], use_focal=bool(HP["use_focal"]), rampup_epochs=1500)
        if name == "reference":
            ref = ref_loader.dycon_losses()
            self.stock = ref_loader.stock_losses()
            self.uncl, self.fecl, self.beta_fn = ref.UnCLoss(), ref.FeCLoss(device=device, **ctor), ref.adaptive_beta
        else:
            import dycon_paper_replication_b200 as ours
            self.ours = ours
            self.uncl, self.fecl, self.beta_fn = ours.UnCLoss(), ours.FeCLoss(device=device, **ctor), ours.adaptive_beta
            self.steplosses = ours.StepLosses()
            self.skipped = torch.zeros(1, dtype=torch.int64, device=device)

    # ---- one iteration of the loop; `ev` collects CUDA events when timing ----
    def step(self, image, label, labeled_bs, epoch, ev=None):
        import torch
        import torch.nn.functional as F
        mark = (lambda k: ev.setdefault(k, []).append(_event())) if ev is not None else (lambda k: None)
        beta = self.beta_fn(epoch=epoch, total_epochs=self.max_epoch, max_beta=HP["beta_max"], min_beta=HP["beta_min"])
        torch.manual_seed(10_000 + self.iter)               # same dropout masks and input noise in every arm
        mark("start")
        noise = torch.clamp(torch.randn_like(image) * 0.1, -0.2, 0.2)           # :201-202
        _, stud_logits, stud_features = self.model(image)                        # :204
        with torch.no_grad():
            _, ema_logits, ema_features = self.ema(image + noise)                # :205-206
        mark("forward")
        cw = consistency_weight(self.iter)
        if self.name == "fused":
            u_loss, loss_seg, loss_seg_dice, cons = self.steplosses(stud_logits, ema_logits, label, labeled_bs, beta)
            f_loss = self.fecl.from_features(stud_features, label, ema_features, None, epoch)
        else:
            stud_probs, ema_probs = F.softmax(stud_logits, dim=1), F.softmax(ema_logits, dim=1)        # :208-209
            loss_seg = F.cross_entropy(stud_logits[:labeled_bs], label[:labeled_bs])                   # :216
            dice = self.stock.dice_loss if self.name == "reference" else _dice
            loss_seg_dice = dice(stud_probs[:labeled_bs, 1], label[:labeled_bs] == 1)                  # :217
            B, C = stud_features.shape[:2]
            stud_emb = F.normalize(torch.transpose(stud_features.view(B, C, -1), 1, 2), dim=-1)        # :219-222
            ema_emb = F.normalize(torch.transpose(ema_features.view(B, C, -1), 1, 2), dim=-1)          # :224-226
            k = HP["feature_scaler"] * 4
            mask = (F.avg_pool3d(label.float(), kernel_size=k, stride=k) > 0.5).float().reshape(B, -1).unsqueeze(1)   # :229-232
            f_loss = self.fecl(feat=stud_emb, mask=mask, teacher_feat=ema_emb, gambling_uncertainty=None, epoch=epoch)
            u_loss = self.uncl(stud_logits, ema_logits, beta)                                           # :254
            mse = self.stock.softmax_mse_loss if self.name == "reference" else _softmax_mse
            cons = mse(stud_probs[labeled_bs:], ema_probs[labeled_bs:]).mean()                          # :255
        loss = HP["l_weight"] * (loss_seg + loss_seg_dice) + cw * cons + HP["u_weight"] * (f_loss + u_loss)     # :258
        mark("losses")
        if self.name == "fused":
            flag = self.ours.loss_is_finite_flag(loss.detach(), counter=self.skipped)      # no host sync (:261-263)
            self.opt.zero_grad()
            loss.backward()
            mark("backward")
            self.ours.sgd_clip_ema_step(self.opt, self.model, self.ema, 1.0, HP["ema_decay"], self.iter, skip_flag=flag)
        else:
            if torch.isnan(loss) or torch.isinf(loss):                                                  # :261-263 (host sync)
                return None
            self.opt.zero_grad()
            loss.backward()                                                                             # :266
            mark("backward")
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=1.0)                      # :269
            self.opt.step()                                                                             # :270
            if self.name == "reference":
                alpha = min(1 - 1 / (self.iter + 1), HP["ema_decay"])                                   # :105-109
                for ep, p in zip(self.ema.parameters(), self.model.parameters()):
                    ep.data.mul_(alpha).add_(p.data, alpha=1 - alpha)
            else:
                self.ours.update_ema_variables(self.model, self.ema, HP["ema_decay"], self.iter)       # :272
        mark("update")
        self.iter += 1
        lr = HP["base_lr"] * (1.0 - self.iter / HP["max_iterations"]) ** 0.9
        for g in self.opt.param_groups:
            g["lr"] = lr
        return [x.detach() for x in (loss, loss_seg, loss_seg_dice, u_loss, f_loss, cons)]


def _event():
    import torch
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _dice(score, target):            # the stock Dice as plain PyTorch ops in the drop-in arm (it stays on PyTorch there)
    import torch
    target = target.float()
    return 1 - (2 * torch.sum(score * target) + 1e-5) / (torch.sum(score * score) + torch.sum(target * target) + 1e-5)


def _softmax_mse(a, b):
    import torch.nn.functional as F
    return (F.softmax(a, dim=1) - F.softmax(b, dim=1)) ** 2


def run(args):
    import numpy as np
    import torch
    from dycon_paper_replication_b200.synthetic import SHAPES
    from oracle import ref_loader
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True              # train_DyCON_Pancreas.py:114-116
    dev = torch.device("cuda", 0)
    patch = tuple(int(x) for x in args.patch.split(",")) if args.patch else SHAPES[args.shape][0]
    torch.manual_seed(1337)
    UNet3D = ref_loader.unet3d()
    model = UNet3D(in_channels=1, n_classes=2, scale_factor=HP["feature_scaler"], use_aspp=False).to(dev)
    n_batches = 4
    batches = make_batches(patch, args.batch, n_batches, dev)
    max_epoch = 300
    arms = [Arm(n, model, dev, max_epoch) for n in args.arms.split(",")]
    names = ["loss", "loss_ce", "loss_dice", "u_loss", "f_loss", "consistency"]
    traj = {a.name: [] for a in arms}
    for it in range(args.steps):
        image, label = batches[it % n_batches]
        for a in arms:
            out = a.step(image, label, args.labeled_bs, epoch=it // 2)
            traj[a.name].append([float(x) for x in out])
    torch.cuda.synchronize()
    result = {"config": f"{args.shape} train step: reference UNet3D (scaler 2) student + teacher, patch {patch}, B={args.batch} "
                        f"({args.labeled_bs} labelled), {args.steps} steps", "loss_names": names, "trajectory": {}, "timing": {}}
    base = np.array(traj[arms[0].name])
    for a in arms:
        t = np.array(traj[a.name])
        result["trajectory"][a.name] = {"first": t[0].tolist(), "last": t[-1].tolist(),
                                        "max_rel_dev_vs_" + arms[0].name: (np.abs(t - base) / np.maximum(np.abs(base), 1e-6)).max(axis=0).tolist()}
    # ---- timed breakdown ----
    for a in arms:
        ev = {}
        for it in range(3):
            a.step(*batches[it % n_batches], args.labeled_bs, epoch=30)
        torch.cuda.synchronize()
        for it in range(args.timed):
            a.step(*batches[it % n_batches], args.labeled_bs, epoch=30, ev=ev)
        torch.cuda.synchronize()
        seg = lambda x, y: float(np.median([p.elapsed_time(q) for p, q in zip(ev[x], ev[y])]))
        result["timing"][a.name] = {"ms_per_step": seg("start", "update"), "forward_ms": seg("start", "forward"),
                                    "losses_ms": seg("forward", "losses"), "backward_ms": seg("losses", "backward"),
                                    "update_ms": seg("backward", "update"),
                                    "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}
    dev_ok = all(max(v["max_rel_dev_vs_" + arms[0].name]) <= 2e-2 for v in result["trajectory"].values())
    result["ok"] = bool(dev_ok)
    result["tolerance"] = "max over steps and loss terms of |arm - reference| / |reference| <= 2e-2 (fp16 FeCL operands, cuDNN float order)"
    return result


def main(argv=None):
    args = parse(argv)
    res = run(args)
    text = json.dumps(res, indent=1)
    if args.json:
        open(args.json, "w").write(text)
    print(json.dumps(res))
    return 0 if res["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
