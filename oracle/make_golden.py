"""Generate tests/golden/*.npz by running the UNMODIFIED reference -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

The reference ships no golden vectors, so parity is pinned
by executing ``/root/reference/code/utils/dycon_losses.py`` (imported by path; it
needs only ``math`` and ``torch``) on small seeded inputs, in fp32 and in fp64,
and committing inputs + outputs.  The EMA fixture executes the reference's
per-parameter loop restated verbatim in behaviour from
``code/train_DyCON_BraTS19.py:155-164`` on ``nn.Module`` parameters (the script
itself cannot be imported: it parses argv and needs absent packages).
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("DYCON_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _run(fn, leaf, *args, dtype, go=1.0, **kw):
    x = leaf.to(dtype).clone().requires_grad_(True)
    conv = [a.to(dtype) if torch.is_tensor(a) and a.is_floating_point() else a for a in args]
    ckw = {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in kw.items()}
    loss = fn(x, *conv, **ckw)
    (loss * go).backward()
    return loss.detach().numpy(), x.grad.detach().numpy()


def uncl_cases(ref):
    g = torch.Generator().manual_seed(1337)
    crit = ref.UnCLoss()
    cases = {}
    specs = {
        "c2_beta5": ((2, 2, 6, 5, 4), 5.0, 2.0),
        "c2_beta05": ((2, 2, 6, 5, 4), 0.5, 2.0),
        "c3_beta158": ((2, 3, 4, 4, 3), 1.58, 2.0),
        "c2_confident": ((3, 2, 5, 3, 7), 1.58, 9.0),     # p < 1e-6: the EPS inside the log matters
        "c4_odd": ((1, 4, 3, 3, 5), 0.8, 1.0),
    }
    for name, (shape, beta, scale) in specs.items():
        s = scale * torch.randn(shape, generator=g)
        t = s + 0.5 * torch.randn(shape, generator=g)
        l32, g32 = _run(lambda a, b: crit(a, b, beta), s, t, dtype=torch.float32, go=0.5)
        l64, g64 = _run(lambda a, b: crit(a, b, beta), s, t, dtype=torch.float64, go=0.5)
        cases[name] = dict(s=s.numpy(), t=t.numpy(), beta=np.float64(beta), go=np.float64(0.5),
                           loss32=l32, grad32=g32, loss64=l64, grad64=g64)
    return cases


def _structured(b, n, d, g, mask, tn=0.3):
    c = F.normalize(torch.randn(d, generator=g), dim=0)
    proto = F.normalize(torch.randn(2, d, generator=g), dim=-1)
    z = torch.randn(b, n, d, generator=g) / d ** 0.5
    x = c + 0.6 * proto[mask.reshape(b, n).long()] + 1.2 * z
    tx = x + tn * torch.randn(b, n, d, generator=g) / d ** 0.5
    return F.normalize(x, dim=-1), F.normalize(tx, dim=-1)


def fecl_cases(ref, legacy):
    g = torch.Generator().manual_seed(1337)
    cases = {}

    def add(name, feat, mask, teacher, unc, epoch, go=0.5, **ctor):
        crit = ref.FeCLoss(device="cpu", **ctor)
        out = {}
        for tag, dt in (("32", torch.float32), ("64", torch.float64)):
            crit_dt = crit
            # the reference builds torch.eye() in default dtype; set it so fp64 stays fp64
            torch.set_default_dtype(dt)
            try:
                l, gr = _run(lambda f, m, t, u: crit_dt(feat=f, mask=m, teacher_feat=t,
                                                       gambling_uncertainty=u, epoch=epoch),
                             feat, mask, teacher, unc, dtype=dt, go=go)
            finally:
                torch.set_default_dtype(torch.float32)
            out["loss" + tag], out["grad" + tag] = l, gr
        rec = dict(feat=feat.numpy(), mask=mask.numpy(), epoch=np.int64(epoch), go=np.float64(go),
                   temperature=np.float64(crit.temperature), gamma=np.float64(crit.gamma),
                   use_focal=np.int64(crit.use_focal), rampup_epochs=np.int64(crit.rampup_epochs),
                   lambda_cross=np.float64(crit.lambda_cross), **out)
        if teacher is not None:
            rec["teacher"] = teacher.numpy()
        if unc is not None:
            rec["unc"] = unc.numpy()
        cases[name] = rec

    b, n, d = 2, 80, 32
    mask = (torch.rand(b, 1, n, generator=g) < 0.3).float()
    f, t = _structured(b, n, d, g, mask)
    add("plain", f, mask, None, None, 0, temperature=0.6, gamma=2.0, use_focal=False, rampup_epochs=1500)
    add("focal_teacher_e100", f, mask, t, None, 100, temperature=0.6, gamma=2.0, use_focal=True,
        rampup_epochs=1500)
    add("focal_teacher_e1500", f, mask, t, None, 1500, temperature=0.6, gamma=2.0, use_focal=True,
        rampup_epochs=1500)
    add("nofocal_teacher_lam2", f, mask, t, None, 700, temperature=0.3, gamma=2.0, use_focal=False,
        rampup_epochs=1500, lambda_cross=2.0)
    add("focal_gamma3", f, mask, None, None, 10, temperature=0.6, gamma=3.0, use_focal=True, rampup_epochs=1500)
    unc = torch.rand(b, n, generator=g)
    add("gambling_overrides_focal", f, mask, t, unc, 100, temperature=0.6, gamma=2.0, use_focal=True,
        rampup_epochs=1500)

    # edge: sample 0 all background (no negatives at all), sample 1 has exactly one foreground
    # voxel (a row with P_i = 1) -- SURVEY.md section 8(d)
    m2 = torch.zeros(b, 1, n)
    m2[1, 0, 17] = 1.0
    f2, t2 = _structured(b, n, d, g, m2)
    add("edge_empty_and_single", f2, m2, t2, None, 100, temperature=0.6, gamma=2.0, use_focal=True,
        rampup_epochs=1500)

    # iid features: D=16 stresses -log(1-cs); D=64 with epoch at ramp end -> cnt = 0
    b3, n3 = 2, 72
    m3 = (torch.rand(b3, 1, n3, generator=g) < 0.4).float()
    f3 = F.normalize(torch.randn(b3, n3, 16, generator=g), dim=-1)
    t3 = F.normalize(f3 + 0.3 * torch.randn(b3, n3, 16, generator=g), dim=-1)
    add("iid_d16", f3, m3, t3, None, 100, temperature=0.6, gamma=2.0, use_focal=True, rampup_epochs=1500)
    f4 = F.normalize(torch.randn(b3, n3, 64, generator=g), dim=-1)
    t4 = F.normalize(torch.randn(b3, n3, 64, generator=g), dim=-1)
    add("iid_d64_cnt0", f4, m3, t4, None, 1500, temperature=0.6, gamma=2.0, use_focal=True,
        rampup_epochs=1500)
    # strided caller layout (D*N, 1, N): values identical, exercises the boundary's stride handling
    f5 = f.transpose(1, 2).contiguous().transpose(1, 2)
    assert f5.stride() == (n * d, 1, n)
    add("strided_layout", f5, mask, t.transpose(1, 2).contiguous().transpose(1, 2), None, 100,
        temperature=0.6, gamma=2.0, use_focal=True, rampup_epochs=1500)

    # legacy losses.FeCLoss(device, temperature).forward(feat, mask)   (code/utils/losses.py:221-250)
    if legacy is not None:
        crit = legacy.FeCLoss(device="cpu", temperature=0.6)
        l32, g32 = _run(lambda x, m: crit(x, m), f, mask, dtype=torch.float32, go=1.0)
        cases["legacy_plain"] = dict(feat=f.numpy(), mask=mask.numpy(), loss32=l32, grad32=g32,
                                     temperature=np.float64(0.6))
    return cases


def scalar_cases(ref):
    rows = []
    for e, tot in ((0, 300), (1, 300), (150, 300), (300, 300), (7, 41)):
        rows.append(("adaptive_beta", e, tot, 5.0, 0.5, ref.adaptive_beta(e, tot, 5.0, 0.5)))
    for e, tot, lo, hi in ((0, 1500, 0.3, 0.5), (100, 1500, 0.3, 0.5), (1500, 1500, 0.3, 0.5),
                           (2000, 1500, 1.3, 1.5), (5, 0, 0.3, 0.5), (-3, 10, 0.1, 0.9)):
        rows.append(("sigmoid_rampup", e, tot, lo, hi, ref.sigmoid_rampup(e, tot, lo, hi)))
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 3, 4, generator=g)
    return {"names": np.array([r[0] for r in rows]), "args": np.array([r[1:5] for r in rows], np.float64),
            "values": np.array([r[5] for r in rows], np.float64),
            "gs_in": x.numpy(), "gs_out": ref.gambling_softmax(x).numpy()}


def ema_cases():
    """Reference loop (train_DyCON_BraTS19.py:155-164) executed on real nn.Modules."""
    def reference_loop(model, ema_model, alpha, global_step):
        alpha = min(1 - 1 / (global_step + 1), alpha)
        for ep, p in zip(ema_model.parameters(), model.parameters()):
            ep.data.mul_(alpha).add_(p.data, alpha=1 - alpha)

    torch.manual_seed(1337)
    def net():
        return torch.nn.Sequential(torch.nn.Conv3d(1, 5, 3), torch.nn.BatchNorm3d(5),
                                   torch.nn.Conv3d(5, 2, 1), torch.nn.Linear(7, 3, bias=False))
    out = {}
    for step in (0, 1, 50, 10000):
        m, e = net(), net()
        before = [p.detach().clone().numpy() for p in e.parameters()]
        reference_loop(m, e, 0.99, step)
        out[f"step{step}_student"] = np.concatenate([p.detach().numpy().ravel() for p in m.parameters()])
        out[f"step{step}_before"] = np.concatenate([x.ravel() for x in before])
        out[f"step{step}_after"] = np.concatenate([p.detach().numpy().ravel() for p in e.parameters()])
    out["sizes"] = np.array([p.numel() for p in net().parameters()], np.int64)
    return out


def main():
    ref = _load(os.path.join(REF, "code", "utils", "dycon_losses.py"), "_ref_dycon_losses")
    try:
        legacy = _load(os.path.join(REF, "code", "utils", "losses.py"), "_ref_losses")
    except Exception as exc:      # noqa: BLE001 - losses.py imports numpy/torch only, but be explicit
        print("legacy losses.py not importable:", exc, file=sys.stderr)
        legacy = None
    os.makedirs(OUT, exist_ok=True)
    for fam, cases in (("uncl", uncl_cases(ref)), ("fecl", fecl_cases(ref, legacy))):
        flat = {f"{c}/{k}": v for c, rec in cases.items() for k, v in rec.items()}
        np.savez_compressed(os.path.join(OUT, f"{fam}.npz"), **flat)
        print(fam, sorted(cases))
    np.savez_compressed(os.path.join(OUT, "scalars.npz"), **scalar_cases(ref))
    np.savez_compressed(os.path.join(OUT, "ema.npz"), **ema_cases())
    print("wrote", OUT, {f: os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)})


if __name__ == "__main__":
    main()
