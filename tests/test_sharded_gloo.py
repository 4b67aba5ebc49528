"""N>1 protocol on CPU: world_size-2 gloo processes run the product's host-side sharding logic
(dycon_paper_replication_b200.sharded) with the oracle standing in for the kernels, and must reproduce
the single-process oracle on the concatenated batch -- loss AND per-shard gradient."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import load_golden, normwise
    from dycon_paper_replication_b200 import sharded
    from oracle import closed_form, torch_port
    try:
        # ---------------- FeCL: shard the 2-sample golden fixture one sample per rank
        rec = load_golden("fecl")["focal_teacher_e100"]
        B, N, _ = rec["feat"].shape
        lo, hi = sharded.shard_bounds(B, rank, world)
        thr = torch_port.ramp_threshold(int(rec["epoch"]), int(rec["rampup_epochs"]), 0.3, 0.5)
        kw = dict(inv_tau=1.0 / float(rec["temperature"]), gamma=float(rec["gamma"]), use_focal=bool(rec["use_focal"]),
                  cross_thresh=thr, lambda_cross=float(rec["lambda_cross"]), go=float(rec["go"]))
        local = closed_form.fecl(rec["feat"][lo:hi], rec["mask"][lo:hi], rec["teacher"][lo:hi], None,
                                 rows_global=B * N, **kw)
        sums = torch.tensor([local["student_sum"], local["cross_sum"], local["cnt"]], dtype=torch.float64)
        sharded.all_reduce_sums(sums, dist.group.WORLD)
        loss = sharded.fecl_loss_from_sums(sums, 1.0 / (B * N), kw["lambda_cross"], True)
        assert abs(loss.item() - float(rec["loss64"])) <= 1e-6 * abs(float(rec["loss64"])), (loss.item(), rec["loss64"])
        # backward: local kernel with the REDUCED count == the rank's slice of the global gradient
        again = closed_form.fecl(rec["feat"][lo:hi], rec["mask"][lo:hi], rec["teacher"][lo:hi], None,
                                 rows_global=B * N, cnt_global=float(sums[2]), **kw)
        assert normwise(again["grad"], rec["grad64"][lo:hi]) <= 1e-9
        # ---------------- UnCL
        urec = load_golden("uncl")["c2_beta5"]
        Bu = urec["s"].shape[0]
        ulo, uhi = sharded.shard_bounds(Bu, rank, world)
        V = int(np.prod(urec["s"].shape[2:]))
        ul = closed_form.uncl(urec["s"][ulo:uhi], urec["t"][ulo:uhi], float(urec["beta"]), go=float(urec["go"]),
                              count=Bu * V)
        tot = torch.tensor([ul["sum"]], dtype=torch.float64)
        sharded.all_reduce_sums(tot, dist.group.WORLD)
        uloss = sharded.uncl_loss_from_sum(tot, 1.0 / (Bu * V))
        assert abs(uloss.item() - float(urec["loss64"])) <= 1e-6 * abs(float(urec["loss64"]))
        assert normwise(ul["grad"], urec["grad64"][ulo:uhi]) <= 1e-10
        # ---------------- host helpers of the global-negatives protocol (all-gather in rank order; the reduced
        #                  loss helpers fall back to the backend all-reduce for CPU tensors)
        assert sharded.group_size_rank(dist.group.WORLD) == (world, rank)
        # with a process group the denominators are always global (never local B with globally reduced sums)
        assert sharded.global_batch_of(3, None, dist.group.WORLD) == 3 * world
        assert sharded.global_batch_of(3, 7, dist.group.WORLD) == 7 and sharded.global_batch_of(3, None, None) == 3
        rows = torch.arange(6, dtype=torch.float32).reshape(3, 2) + 100 * rank
        full = sharded.all_gather_rows(rows, dist.group.WORLD)
        want = torch.cat([torch.arange(6, dtype=torch.float32).reshape(3, 2) + 100 * r for r in range(world)])
        assert torch.equal(full, want)
        s3 = torch.tensor([1.0 + rank, 2.0, 4.0 * (rank + 1)], dtype=torch.float64)
        red = sharded.reduce_fecl(s3, 0.5, 2.0, True, dist.group.WORLD)
        tot3 = [sum(1.0 + r for r in range(world)), 2.0 * world, sum(4.0 * (r + 1) for r in range(world))]
        assert abs(red.item() - (tot3[0] * 0.5 + 2.0 * tot3[1] / tot3[2])) < 1e-6
        u1 = sharded.reduce_uncl(torch.tensor([3.0 * (rank + 1)], dtype=torch.float64), 0.25, dist.group.WORLD)
        assert abs(u1.item() - 0.25 * sum(3.0 * (r + 1) for r in range(world))) < 1e-6
        out.put((rank, "ok"))
    except Exception as exc:       # noqa: BLE001 - report to the parent
        out.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process_oracle():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(out.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == {0: "ok", 1: "ok"}, results


def test_shard_bounds_cover_the_batch_and_reject_oversharding():
    sys.path.insert(0, ROOT)
    from dycon_paper_replication_b200 import sharded
    for b in (1, 4, 7, 8, 32):
        for w in (1, 2, 3, 4, 8):
            if w > b:
                with pytest.raises(ValueError):
                    sharded.shard_bounds(b, 0, w)
                continue
            spans = [sharded.shard_bounds(b, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_loss_assembly_matches_reference_formula():
    sys.path.insert(0, ROOT)
    from dycon_paper_replication_b200 import sharded
    sums = torch.tensor([12.5, 3.0, 4.0], dtype=torch.float64)
    assert abs(sharded.fecl_loss_from_sums(sums, 0.1, 2.0, True).item() - (1.25 + 2.0 * 0.75)) < 1e-6
    assert abs(sharded.fecl_loss_from_sums(sums, 0.1, 2.0, False).item() - 1.25) < 1e-6
    zero = torch.tensor([12.5, 0.0, 0.0], dtype=torch.float64)           # cnt == 0 -> cross term is 0/1e-18 = 0
    assert abs(sharded.fecl_loss_from_sums(zero, 0.1, 1.0, True).item() - 1.25) < 1e-6
