#!/usr/bin/env python
"""Multi-GPU check of the sharded path (run under torchrun on N B200s of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_check.py

1. the NVLink peer-memory exchange (dycon_exchange_sums) against NCCL all_reduce: eager back-to-back
   calls (slot reuse) and CUDA-graph replays;
2. the sharded FeCL + UnCL modules (process_group=, global_batch=) against the float64 closed-form oracle on
   the concatenated batch: every rank must obtain the single-process loss and its own slice of the gradient;
3. FeCL with global negatives (cross_gpu_negatives=True) against the reference on the merged batch.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    from dycon_paper_replication_b200 import FeCLoss, UnCLoss, sharded
    from dycon_paper_replication_b200.synthetic import make_inputs
    from oracle import closed_form, torch_port

    g = dist.group.WORLD
    # ---- 1. exchange vs NCCL ------------------------------------------------------------------
    gen = torch.Generator().manual_seed(7 + rank)
    used_peer = None
    for it in range(200):
        n = 1 + it % 7
        x = torch.randn(n, generator=gen, dtype=torch.float64).to(dev)
        want = x.clone()
        dist.all_reduce(want, group=g)
        got = sharded.all_reduce_sums(x.clone(), g)
        if used_peer is None:
            used_peer = sharded._exchanges.get((id(g), dev.index)) is not None
        assert torch.allclose(got, want, rtol=1e-13, atol=1e-13), (it, got, want)
    # back-to-back without host synchronisation, then graph replays
    xs = [torch.full((3,), float(rank + 1) * (k + 1), dtype=torch.float64, device=dev) for k in range(64)]
    for x in xs:
        sharded.all_reduce_sums(x, g)
    torch.cuda.synchronize()
    tot = sum(range(1, world + 1))
    for k, x in enumerate(xs):
        assert torch.equal(x, torch.full((3,), float(tot * (k + 1)), dtype=torch.float64, device=dev)), (k, x)
    src = torch.full((3,), float(rank + 1), dtype=torch.float64, device=dev)
    buf = torch.empty_like(src)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        buf.copy_(src)
        sharded.all_reduce_sums(buf, g)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        buf.copy_(src)
        sharded.all_reduce_sums(buf, g)
    for _ in range(20):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(buf, torch.full((3,), float(tot), dtype=torch.float64, device=dev)), buf

    # ---- 2. sharded losses vs the single-process oracle -------------------------------------------
    B = 2 * world
    inp = make_inputs("tiny", batch=B, dim=32, mask_kind="bernoulli", seed=99)
    lo, hi = sharded.shard_bounds(B, rank, world)
    f = inp.feat[lo:hi].to(dev).requires_grad_(True)
    s = inp.s_logits[lo:hi].to(dev).requires_grad_(True)
    worst = 0.0
    full_f = inp.feat.to(dev).requires_grad_(True)
    full_s = inp.s_logits.to(dev).requires_grad_(True)
    for mode in ("fp32", "fp16"):
        f.grad = s.grad = full_f.grad = full_s.grad = None
        ctor = dict(temperature=0.6, gamma=2.0, use_focal=True, rampup_epochs=1500, precision=mode)
        # the same kernels on the whole batch in this process: the sharded run must reproduce its loss and its
        # slice of the gradient (same per-sample arithmetic; only the denominators and cnt are global)
        fl0 = FeCLoss(dev, **ctor)(full_f, inp.mask.to(dev), inp.teacher.to(dev), None, 100)
        ul0 = UnCLoss()(full_s, inp.t_logits.to(dev), 1.58)
        (0.5 * (fl0 + ul0)).backward()
        fecl = FeCLoss(dev, process_group=g, global_batch=B, **ctor)
        uncl = UnCLoss(process_group=g, global_batch=B)
        fl = fecl(f, inp.mask[lo:hi].to(dev), inp.teacher[lo:hi].to(dev), None, 100)
        ul = uncl(s, inp.t_logits[lo:hi].to(dev), 1.58)
        (0.5 * (fl + ul)).backward()
        assert abs(fl.item() - fl0.item()) <= 2e-6 * abs(fl0.item()), (mode, fl.item(), fl0.item())
        assert abs(ul.item() - ul0.item()) <= 2e-6 * abs(ul0.item()), (ul.item(), ul0.item())
        eg = ((f.grad - full_f.grad[lo:hi]).abs().max() / full_f.grad.abs().max()).item()
        es = ((s.grad - full_s.grad[lo:hi]).abs().max() / full_s.grad.abs().max()).item()
        assert eg <= 2e-6 and es <= 2e-6, (mode, eg, es)
        worst = max(worst, eg, es)
        if mode == "fp32":      # and the exact mode against the float64 closed form of the reference
            thr = torch_port.ramp_threshold(100, 1500, 0.3, 0.5)
            rf = closed_form.fecl(inp.feat.numpy(), inp.mask.numpy(), inp.teacher.numpy(), None, inv_tau=1 / 0.6,
                                  gamma=2.0, use_focal=True, cross_thresh=thr, go=0.5)
            ru = closed_form.uncl(inp.s_logits.numpy(), inp.t_logits.numpy(), 1.58, go=0.5)
            assert abs(fl.item() - rf["loss"]) <= 1e-5 * abs(rf["loss"]), (fl.item(), rf["loss"])
            assert abs(ul.item() - ru["loss"]) <= 1e-5 * abs(ru["loss"]), (ul.item(), ru["loss"])
            assert np.abs(f.grad.cpu().numpy() - rf["grad"][lo:hi]).max() <= 1e-5 * np.abs(rf["grad"]).max()
            assert np.abs(s.grad.cpu().numpy() - ru["grad"][lo:hi]).max() <= 1e-5 * np.abs(ru["grad"]).max()
    # ---- 2b. the sharded step captured in a CUDA graph (the exchange runs in the tail of the forward kernels):
    #          replays must reproduce the eager result bit for bit
    ctor = dict(temperature=0.6, gamma=2.0, use_focal=True, rampup_epochs=1500, precision="fp16")
    fecl = FeCLoss(dev, process_group=g, global_batch=B, **ctor)
    uncl = UnCLoss(process_group=g, global_batch=B)
    mk, tk, tl = inp.mask[lo:hi].to(dev), inp.teacher[lo:hi].to(dev), inp.t_logits[lo:hi].to(dev)
    # fresh leaves whose first backward runs on the side stream (an AccumulateGrad node bound to the legacy
    # stream cannot be captured)
    fg = inp.feat[lo:hi].to(dev).requires_grad_(True)
    sg = inp.s_logits[lo:hi].to(dev).requires_grad_(True)

    def sharded_step():
        fg.grad = sg.grad = None
        fl = fecl(fg, mk, tk, None, 100)
        ul = uncl(sg, tl, 1.58)
        (0.5 * (fl + ul)).backward()
        return fl, ul

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fl, ul = sharded_step()
        want = (fl.item(), ul.item(), fg.grad.clone(), sg.grad.clone())
        sharded_step()
    torch.cuda.current_stream().wait_stream(side)
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        gfl, gul = sharded_step()
    for _ in range(10):
        graph.replay()
    torch.cuda.synchronize()
    assert gfl.item() == want[0] and gul.item() == want[1], (gfl.item(), want[0], gul.item(), want[1])
    assert torch.equal(fg.grad, want[2]) and torch.equal(sg.grad, want[3])
    ex = sharded._exchanges.get((id(g), dev.index))
    assert ex is None or not ex.timed_out()

    # ---- 3. global negatives (BASELINE config 5): every rank's rows against the rows of ALL ranks ----------
    crit = FeCLoss(dev, precision="fp16", process_group=g, cross_gpu_negatives=True, temperature=0.6, gamma=2.0,
                   use_focal=True, rampup_epochs=1500)
    f.grad = None
    gl = crit(f, inp.mask[lo:hi].to(dev), inp.teacher[lo:hi].to(dev), None, 100)
    (0.5 * gl).backward()
    Bn, Nn, Dn = inp.feat.shape
    thr = torch_port.ramp_threshold(100, 1500, 0.3, 0.5)
    rg = closed_form.fecl(inp.feat.reshape(1, Bn * Nn, Dn).numpy(), inp.mask.reshape(1, 1, Bn * Nn).numpy(),
                          inp.teacher.reshape(1, Bn * Nn, Dn).numpy(), None, inv_tau=1 / 0.6, gamma=2.0, use_focal=True,
                          cross_thresh=thr, go=0.5, ambiguity=5e-4)
    assert abs(gl.item() - rg["loss"]) <= 2e-3 * abs(rg["loss"]), (gl.item(), rg["loss"])
    full = [torch.empty_like(f.grad.contiguous()) for _ in range(world)]
    dist.all_gather(full, f.grad.contiguous(), group=g)
    got = torch.cat(full).cpu().numpy().reshape(rg["grad"].shape)
    gerr = closed_form.fecl_grad_error(got, rg, inp.teacher.reshape(1, Bn * Nn, Dn).numpy())
    assert gerr <= 2e-3, gerr
    dist.barrier()
    if rank == 0:
        print(f"multi_gpu_check ok: world={world} peer_exchange={used_peer} sharded-vs-unsharded max err={worst:.2e} "
              f"global-negatives grad err={gerr:.1e}", flush=True)
    os._exit(0)      # skip the NCCL teardown (it can hang once graphs captured NCCL work)


if __name__ == "__main__":
    main()
