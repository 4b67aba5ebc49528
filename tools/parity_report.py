#!/usr/bin/env python
"""Parity numbers of the FeCL CUDA path in every precision mode (run on a B200; test infrastructure).

For every golden fixture and seeded shape: loss error, gradient error against the fp64 closed form (plain), against
the flip-tolerant comparator (fitted) and against the STRICT comparator that predicts the threshold flips from the
rounded operands, plus the error against the closed form evaluated ON the rounded operands (what a 16-bit mode
computes exactly, up to fp32 accumulation).  tests/test_gpu_fecl.py asserts bounds derived from this table.

    python tools/parity_report.py [--json out.json]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_golden, normwise  # noqa: E402
from oracle import closed_form, torch_port  # noqa: E402

AMBIGUITY = {"fp32": 2e-6, "fp16": 5e-4, "bf16": 4e-3}


def run(feat, mask, teacher, unc, epoch, go, mode, **ctor):
    from dycon_paper_replication_b200 import FeCLoss
    crit = FeCLoss(device="cuda", precision=mode, **ctor)
    f = feat.cuda().requires_grad_(True)
    loss = crit(feat=f, mask=mask.cuda(), teacher_feat=None if teacher is None else teacher.cuda(),
                gambling_uncertainty=None if unc is None else unc.cuda(), epoch=epoch)
    (loss * go).backward()
    return loss.detach().cpu().double().item(), f.grad.detach().cpu().numpy()


def cases():
    from dycon_paper_replication_b200.synthetic import make_inputs
    fecl = load_golden("fecl")
    for name in sorted(k for k in fecl if k != "legacy_plain"):
        rec = fecl[name]
        t = lambda k: torch.from_numpy(rec[k]) if k in rec else None
        ctor = dict(temperature=float(rec["temperature"]), gamma=float(rec["gamma"]), use_focal=bool(rec["use_focal"]),
                    rampup_epochs=int(rec["rampup_epochs"]), lambda_cross=float(rec["lambda_cross"]))
        yield "golden/" + name, torch.from_numpy(rec["feat"]), t("mask"), t("teacher"), t("unc"), int(rec["epoch"]), float(rec["go"]), ctor
    ctor = dict(temperature=0.6, gamma=2.0, use_focal=True, rampup_epochs=1500, lambda_cross=1.0)
    for shape, dim, fk, mk, epoch in [("tiny", 32, "structured", "bernoulli", 100), ("tiny", 64, "iid", "bernoulli", 0),
                                      ("brats19", 256, "structured", "blob", 100), ("brats19", 16, "structured", "bernoulli", 100),
                                      ("brats19", 256, "iid", "blob", 1500)]:
        inp = make_inputs(shape, dim=dim, feat_kind=fk, mask_kind=mk, empty_first=(shape == "brats19"))
        yield f"seeded/{shape}-D{dim}-{fk}-{mk}-e{epoch}", inp.feat, inp.mask, inp.teacher, None, epoch, 0.5, ctor


def main():
    rows = []
    for name, feat, mask, teacher, unc, epoch, go, ctor in cases():
        thr = torch_port.ramp_threshold(epoch, ctor["rampup_epochs"], 0.3, 0.5)
        kw = dict(inv_tau=1.0 / ctor["temperature"], gamma=ctor["gamma"], use_focal=ctor["use_focal"], cross_thresh=thr,
                  lambda_cross=ctor["lambda_cross"], go=go)
        fn, mn = feat.numpy(), mask.numpy()
        tn = None if teacher is None else teacher.numpy()
        un = None if unc is None else unc.numpy()
        for mode in ("fp32", "fp16", "bf16"):
            loss, grad = run(feat, mask, teacher, unc, epoch, go, mode, **ctor)
            ref = closed_form.fecl(fn, mn, tn, un, ambiguity=AMBIGUITY[mode], **kw)
            row = {"case": name, "mode": mode, "loss_err": abs(loss - ref["loss"]) / abs(ref["loss"]),
                   "plain": normwise(grad, ref["grad"]), "cnt": ref["cnt"]}
            if tn is not None:
                row["fitted"] = closed_form.fecl_grad_error(grad, ref, tn)
                st = closed_form.fecl_grad_error_strict(grad, ref, fn, tn, mode, thr)
                row.update(strict=st["err"], flipped=st["flipped"], free=st["free"], window=st["window_pairs"],
                           outside=st["outside_flips"])
            if mode != "fp32":
                fq = closed_form.round_operand(fn, mode)
                tq = None if tn is None else closed_form.round_operand(tn, mode)
                rq = closed_form.fecl(fq, mn, tq, un, ambiguity=3e-6, **kw)
                row["loss_err_q"] = abs(loss - rq["loss"]) / abs(rq["loss"])
                row["grad_q"] = closed_form.fecl_grad_error(grad, rq, tq) if tq is not None else normwise(grad, rq["grad"])
            rows.append(row)
            print(json.dumps(row), flush=True)
    if "--json" in sys.argv:
        json.dump(rows, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
