#!/usr/bin/env python
"""In-stream kernel durations of the bench step (CUPTI through torch.profiler): warm caches, real clocks, programmatic
dependent launches in effect -- what ncu's serialised, cache-flushed replays cannot show.

    python tools/kernel_times.py [--steps 30] [--shape brats19] > profiles/rNN_kernel_times.md        # on a B200
"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--shape", default="brats19")
    ap.add_argument("--batch", type=int, default=None)
    args = ap.parse_args()
    from dycon_paper_replication_b200 import FeCLoss, UnCLoss
    from dycon_paper_replication_b200.synthetic import make_inputs
    inp = make_inputs(args.shape, batch=args.batch, dim=256).to("cuda")
    fecl = FeCLoss("cuda", temperature=0.6, gamma=2.0, use_focal=True, rampup_epochs=1500)
    uncl = UnCLoss()
    f = inp.feat.requires_grad_(True)
    s = inp.s_logits.requires_grad_(True)

    def step():
        f.grad = None
        s.grad = None
        loss = 0.5 * (fecl(f, inp.mask, inp.teacher, None, 100) + uncl(s, inp.t_logits, 1.58))
        loss.backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    first = None
    last = 0
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = ev.name
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    total = sum(v[1] for v in agg.values())
    print(f"# In-stream kernel durations, {args.shape} B={inp.feat.shape[0]}, {args.steps} eager steps (torch.profiler / CUPTI)\n")
    print("| kernel | launches per step | avg us | us per step |")
    print("|---|---:|---:|---:|")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        short = name.replace("dycon::<unnamed>::", "").replace("void ", "")[:90]
        print(f"| `{short}` | {n / args.steps:.1f} | {t / n:.2f} | {t / args.steps:.2f} |")
    print(f"\nsum of kernel durations per step: {total / args.steps:.1f} us (kernels chained with programmatic dependent launch overlap)")


if __name__ == "__main__":
    main()
