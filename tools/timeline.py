#!/usr/bin/env python
"""Device-clock timeline of one CTA of the FeCL loss sweep (P2) and of the backward (measurement aid).

    DYCON_TIMELINE=1 python -m dycon_paper_replication_b200.csrc.build             # _dycon_b200_timeline.so (stamps compiled in)
    DYCON_SO_VARIANT=timeline python tools/timeline.py [--sorted] > out.md         # on a B200

Columns are SM cycles since the first stamp of the kernel.  Roles: producer (stage obtained / TMA issued), MMA
issuers (operands landed / accumulator free / MMAs committed [/ H landed / MMA2 committed]), one warp per epilogue
team (loop top / accumulator obtained / arithmetic done [/ h_free obtained] / barrier passed)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    if "--sorted" in sys.argv:          # rows packed sorted by label + class-specialised bodies (off by default)
        os.environ["DYCON_FECL_SORT"] = "1"
    os.environ["DYCON_NO_PDL"] = "1"
    from dycon_paper_replication_b200 import FeCLoss, _lib
    from dycon_paper_replication_b200.synthetic import make_inputs
    inp = make_inputs("brats19", batch=4, dim=256).to("cuda")
    crit = FeCLoss("cuda", temperature=0.6, gamma=2.0, use_focal=True, rampup_epochs=1500)
    f = inp.feat.requires_grad_(True)
    for _ in range(3):
        f.grad = None
        loss = crit(f, inp.mask, inp.teacher, None, 100)
        (0.5 * loss).backward()
    torch.cuda.synchronize()
    buf = np.zeros((2, 4, 64, 8), np.uint64)
    n = _lib.lib().dycon_debug_timeline(buf.ctypes.data_as(ctypes.c_void_p), buf.nbytes)
    if n == 0:
        print("not a timeline build: DYCON_TIMELINE=1 python -m dycon_paper_replication_b200.csrc.build, then run with "
              "DYCON_SO_VARIANT=timeline")
        return
    names = {0: ("tensor-core sweep (similarity sweep in the default fp16 path, loss sweep otherwise)", ["tma:stage", "tma:issued", "mma:bfull", "mma:accfree", "mma:commit", "e0:top", "e0:acc", "e0:math",
                              "e0:bar", "e1:top", "e1:acc", "e1:math", "e1:bar"]),
             1: ("backward", ["tma:stage", "mma1:bfull", "mma1:scfree", "mma1:commit", "mma2:hfull", "mma2:commit", "e0:top",
                              "e0:sc", "e0:math", "e0:hfree", "e0:bar", "e1:top", "e1:sc", "e1:math", "e1:hfree", "e1:bar", "cls"])}
    for kern in (0, 1):
        title, cols = names[kern]
        d = buf[kern].astype(np.int64)
        nz = d[:, :62, :7][d[:, :62, :7] > 0]
        if nz.size == 0:
            continue
        t0 = nz.min()
        rel = lambda v: "" if v <= 0 else str(int(v - t0))
        if kern == 1 and not (d[2:4, :62, :5] > 0).any():
            # the pure-GEMM backward (fecl_tc_bwd_gemm_kernel): producer and MMA issuer only, no pair arithmetic
            print(f"\n## backward GEMM: cycles since the first stamp (accumulators complete: {rel(d[2, 62, 1])}, kernel tail: {rel(d[2, 63, 0])})\n")
            print("| 64-column tile | tma: stage free | mma: operands landed | mma: 12 MMAs issued |")
            print("|---:|---:|---:|---:|")
            for t in range(62):
                row = [d[0, t, 0], d[1, t, 0], d[1, t, 2]]
                if any(v > 0 for v in row):
                    print(f"| {t} | " + " | ".join(rel(v) for v in row) + " |")
            continue
        print(f"\n## {title}: cycles since the first stamp (end of loop e0/e1: {rel(d[2, 62, 0])} / {rel(d[3, 62, 0])}, "
              f"df_full e0: {rel(d[2, 62, 1])}, kernel tail: {rel(d[2, 63, 0])})\n")
        print("| t | " + " | ".join(cols) + " |")
        print("|" + "---:|" * (len(cols) + 1))
        for t in range(62):
            if kern == 0:
                team = 2 + (t & 1)
                row = [d[0, t, 0], d[0, t, 1], d[1, t, 0], d[1, t, 1], d[1, t, 2]]
                e = [d[team, t, k] for k in range(4)]
                row += e + [0] * 4 if team == 2 else [0] * 4 + e
            else:
                team = 2 + (t & 1)
                row = [d[0, t, 0], d[1, t, 0], d[1, t, 1], d[1, t, 2], d[1, t, 3], d[1, t, 4]]
                e = [d[team, t, k] for k in range(5)]
                row += e + [0] * 5 if team == 2 else [0] * 5 + e
            if not any(v > 0 for v in row):
                continue
            cells = [rel(v) for v in row]
            if kern == 1:
                cells.append(str(int(d[0, t, 7])))
            print(f"| {t} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
