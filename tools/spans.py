#!/usr/bin/env python
"""Spans of every CTA of the four kernels of the default FeCL path on the GPU-wide nanosecond timer (measurement aid).

    DYCON_TIMELINE=1 python -m dycon_paper_replication_b200.csrc.build             # _dycon_b200_timeline.so (stamps compiled in)
    DYCON_SO_VARIANT=timeline python tools/spans.py > out.md                       # on a B200

Per kernel: when the first / last CTA entered, reached its main loop (after the programmatic-launch wait), left the loop
and exited, all relative to the entry of the first CTA of the pack kernel of the same step; the distribution of CTA life
times; the SMs used.  Shows launch skew, what the programmatic dependent launches overlap, and the tails."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from dycon_paper_replication_b200 import FeCLoss, UnCLoss, _lib
    from dycon_paper_replication_b200.synthetic import make_inputs
    inp = make_inputs("brats19", batch=4, dim=256).to("cuda")
    crit = FeCLoss("cuda", temperature=0.6, gamma=2.0, use_focal=True, rampup_epochs=1500)
    uncl = UnCLoss()
    f = inp.feat.requires_grad_(True)
    s = inp.s_logits.requires_grad_(True)

    def step():
        f.grad = None
        s.grad = None
        if "--fecl-only" in sys.argv:      # nothing streams through the L2 between the forward and the backward
            loss = 0.5 * crit(f, inp.mask, inp.teacher, None, 100)
        elif "--uncl-first" in sys.argv:
            u = uncl(s, inp.t_logits, 1.58)
            loss = 0.5 * (crit(f, inp.mask, inp.teacher, None, 100) + u)
        else:
            loss = 0.5 * (crit(f, inp.mask, inp.teacher, None, 100) + uncl(s, inp.t_logits, 1.58))
        loss.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        step()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    n1 = 2 * 4 * 64 * 8
    buf = np.zeros(n1 + 4 * 512 * 6, np.uint64)
    n = _lib.lib().dycon_debug_timeline(buf.ctypes.data_as(ctypes.c_void_p), buf.nbytes)
    if n != buf.nbytes:
        print("not a timeline build with spans")
        return
    sp = buf[n1:].reshape(4, 512, 6).astype(np.int64)
    names = ["pack16", "similarity sweep", "row kernel", "backward GEMM"]
    evn = ["entry", "main loop reached", "main loop done", "exit"]
    t0 = sp[0][sp[0][:, 0] > 0][:, 0].min()
    print("# Spans of all CTAs (GPU-wide timer, us since the first CTA of the pack kernel entered), one replay of the captured step\n")
    print("| kernel | CTAs | SMs | event | first | median | last |")
    print("|---|---:|---:|---|---:|---:|---:|")
    for k in range(4):
        d = sp[k][sp[k][:, 0] > 0]
        if d.size == 0:
            continue
        sms = len(set(d[:, 4].tolist()))
        for e in (0, 1, 2, 3):
            v = d[:, e][d[:, e] > 0]
            if v.size == 0:
                continue
            v = (v - t0) / 1e3
            print(f"| {names[k]} | {len(d)} | {sms} | {evn[e]} | {v.min():.2f} | {np.median(v):.2f} | {v.max():.2f} |")
        life = (d[:, 3] - d[:, 0]) / 1e3
        print(f"| {names[k]} | | | CTA life time | {life.min():.2f} | {np.median(life):.2f} | {life.max():.2f} |")
    d = sp[3][sp[3][:, 0] > 0]
    if d.size:
        print("\n## backward GEMM: CTA life time against its work (tiles, tiles with a Gc tile)\n")
        print("| tiles | with Gc | CTAs | loop us (entry -> accumulators complete) median | max | read-out us median |")
        print("|---:|---:|---:|---:|---:|---:|")
        key = d[:, 5]
        for kv in sorted(set(key.tolist())):
            m = d[key == kv]
            loop = (m[:, 1] - m[:, 0]) / 1e3
            ro = (m[:, 3] - m[:, 2]) / 1e3
            print(f"| {kv % 1000} | {kv // 1000} | {len(m)} | {np.median(loop):.2f} | {loop.max():.2f} | {np.median(ro):.2f} |")
    d = sp[2][sp[2][:, 0] > 0]
    if d.size and "--row-detail" in sys.argv:
        print("\n## row kernel: loop time (programmatic-launch wait passed -> rows done) by CTAs per SM and by quarter of the grid\n")
        loop = (d[:, 2] - d[:, 1]) / 1e3
        per_sm = {}
        for k in range(len(d)):
            per_sm.setdefault(int(d[k, 4]), []).append(loop[k])
        by_n = {}
        for sm, v in per_sm.items():
            by_n.setdefault(len(v), []).extend(v)
        for nn, v in sorted(by_n.items()):
            print(f"- SMs holding {nn} CTA(s): {len(v) // nn} SMs, loop {min(v):.2f} / {np.median(v):.2f} / {max(v):.2f} us")
        q = len(d) // 4
        for k in range(4):
            v = loop[k * q:(k + 1) * q]
            print(f"- CTAs {k * q}..{(k + 1) * q - 1}: loop {v.min():.2f} / {np.median(v):.2f} / {v.max():.2f} us")
        order = np.argsort(loop)
        print("- slowest CTAs (cta, sm, loop us): " + ", ".join(f"({int(k)}, {int(d[k, 4])}, {loop[k]:.1f})" for k in order[-8:]))
        print("- fastest CTAs (cta, sm, loop us): " + ", ".join(f"({int(k)}, {int(d[k, 4])}, {loop[k]:.1f})" for k in order[:8]))
    d = sp[1][sp[1][:, 0] > 0]
    if d.size:
        print("\n## similarity sweep: CTA loop time (main loop reached -> done) by sub-tile count\n")
        for kv in sorted(set(d[:, 5].tolist())):
            m = d[d[:, 5] == kv]
            loop = (m[:, 2] - m[:, 1]) / 1e3
            print(f"- {kv} sub-tiles: {len(m)} CTAs, loop {loop.min():.2f} / {np.median(loop):.2f} / {loop.max():.2f} us (min / median / max)")


if __name__ == "__main__":
    main()
