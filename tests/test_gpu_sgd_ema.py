"""Fused clip + SGD + EMA step and the device-side finite flag against the PyTorch calls they
replace -- torch.nn.utils.clip_grad_norm_, torch.optim.SGD.step and the reference's EMA loop
(code/train_DyCON_BraTS19.py:155-164,268,360-372) -- run on identical copies on the same device."""
import copy

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


def reference_ema(model, ema_model, alpha, global_step):          # train_DyCON_BraTS19.py:155-164
    alpha = min(1 - 1 / (global_step + 1), alpha)
    for ep, p in zip(ema_model.parameters(), model.parameters()):
        ep.data.mul_(alpha).add_(p.data, alpha=1 - alpha)


def net():
    torch.manual_seed(3)
    m = torch.nn.Sequential(torch.nn.Conv3d(1, 6, 3), torch.nn.BatchNorm3d(6), torch.nn.Conv3d(6, 33, 3),
                            torch.nn.Flatten(), torch.nn.LazyLinear(5), torch.nn.Linear(5, 1200, bias=False)).cuda()
    m(torch.randn(2, 1, 6, 6, 6, device="cuda"))
    return m


def set_grads(model, seed, scale, frozen=()):
    g = torch.Generator(device="cuda").manual_seed(seed)
    for k, p in enumerate(model.parameters()):
        p.grad = None if k in frozen else scale * torch.randn(p.shape, generator=g, device="cuda")


def close(a, b, tol):
    return (a - b).abs().max().item() <= tol * max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("momentum,wd,nesterov,max_norm", [(0.9, 1e-4, False, 1.0), (0.9, 0.0, True, 0.05),
                                                           (0.0, 1e-4, False, 1e9)])
def test_three_steps_match_the_pytorch_calls(momentum, wd, nesterov, max_norm):
    from dycon_paper_replication_b200 import sgd_clip_ema_step
    a, b = net(), None
    b = copy.deepcopy(a)
    ea, eb = copy.deepcopy(a), copy.deepcopy(a)
    for q in list(ea.parameters()) + list(eb.parameters()):
        q.data.mul_(0.5)
    oa = torch.optim.SGD(a.parameters(), lr=0.01, momentum=momentum, weight_decay=wd, nesterov=nesterov)
    ob = torch.optim.SGD(b.parameters(), lr=0.01, momentum=momentum, weight_decay=wd, nesterov=nesterov)
    for step in range(3):
        frozen = (2,) if step == 1 else ()                 # a parameter without a gradient: not stepped, teacher follows
        set_grads(a, 10 + step, 0.3 * (step + 1), frozen)
        set_grads(b, 10 + step, 0.3 * (step + 1), frozen)
        for grp in oa.param_groups + ob.param_groups:      # the scripts change lr every iteration
            grp["lr"] = 0.01 * (1 - step / 10) ** 0.9
        want_norm = torch.nn.utils.clip_grad_norm_(a.parameters(), max_norm=max_norm)
        oa.step()
        reference_ema(a, ea, 0.99, step)
        got_norm = sgd_clip_ema_step(ob, b, eb, max_norm, 0.99, step)
        assert abs(got_norm.item() - want_norm.item()) <= 2e-6 * want_norm.item()
        for pa, pb in zip(a.parameters(), b.parameters()):
            assert close(pb.data, pa.data, 2e-6), step
        for pa, pb in zip(ea.parameters(), eb.parameters()):
            assert close(pb.data, pa.data, 2e-6), step
        if momentum:
            for pa, pb in zip(a.parameters(), b.parameters()):
                if pa in oa.state and "momentum_buffer" in oa.state[pa] and oa.state[pa]["momentum_buffer"] is not None:
                    assert close(ob.state[pb]["momentum_buffer"], oa.state[pa]["momentum_buffer"], 2e-6), step
    # buffers (BatchNorm running stats) are untouched by the step, as in the reference
    for x, y in zip(a.buffers(), b.buffers()):
        assert torch.equal(x, y)
    # the optimizer state stays interchangeable with torch.optim.SGD
    ob.load_state_dict(oa.state_dict())


def test_unclipped_step_is_bit_identical_to_pytorch():
    """With the clip coefficient exactly 1 the only difference to the PyTorch ops would be rounding order."""
    from dycon_paper_replication_b200 import sgd_clip_ema_step
    a = net()
    b, ea, eb = copy.deepcopy(a), copy.deepcopy(a), copy.deepcopy(a)
    oa = torch.optim.SGD(a.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    ob = torch.optim.SGD(b.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    for step in range(3):
        set_grads(a, 20 + step, 1e-3)
        set_grads(b, 20 + step, 1e-3)
        torch.nn.utils.clip_grad_norm_(a.parameters(), max_norm=1e6)
        oa.step()
        reference_ema(a, ea, 0.99, step + 5)
        sgd_clip_ema_step(ob, b, eb, 1e6, 0.99, step + 5)
    same = sum(torch.equal(x.data, y.data) for x, y in zip(a.parameters(), b.parameters()))
    same_e = sum(torch.equal(x.data, y.data) for x, y in zip(ea.parameters(), eb.parameters()))
    n = len(list(a.parameters()))
    assert same == n and same_e == n, (same, same_e, n)


def test_finite_flag_skips_the_step_without_a_host_sync():
    from dycon_paper_replication_b200 import loss_is_finite_flag, sgd_clip_ema_step
    b = net()
    eb = copy.deepcopy(b)
    ob = torch.optim.SGD(b.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    set_grads(b, 1, 0.1)
    before = [p.detach().clone() for p in b.parameters()]
    before_e = [p.detach().clone() for p in eb.parameters()]
    counter = torch.zeros(1, dtype=torch.int64, device="cuda")
    good = torch.tensor(1.5, device="cuda")
    for bad in (torch.tensor(float("nan"), device="cuda"), torch.tensor(float("inf"), device="cuda")):
        flag = loss_is_finite_flag(good, bad, counter=counter)
        sgd_clip_ema_step(ob, b, eb, 1.0, 0.99, 7, skip_flag=flag)
        assert flag.item() == 1
        for p, q in zip(b.parameters(), before):
            assert torch.equal(p.data, q)
        for p, q in zip(eb.parameters(), before_e):
            assert torch.equal(p.data, q)
    assert counter.item() == 2
    flag = loss_is_finite_flag(good, counter=counter)
    sgd_clip_ema_step(ob, b, eb, 1.0, 0.99, 7, skip_flag=flag)
    assert flag.item() == 0 and counter.item() == 2
    assert not torch.equal(next(b.parameters()).data, before[0])


def test_graph_capture_of_the_whole_update():
    from dycon_paper_replication_b200 import loss_is_finite_flag, sgd_clip_ema_step
    b = net()
    eb = copy.deepcopy(b)
    ob = torch.optim.SGD(b.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    set_grads(b, 4, 0.1)
    loss = torch.tensor(0.7, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        sgd_clip_ema_step(ob, b, eb, 1.0, 0.99, 0, skip_flag=loss_is_finite_flag(loss))     # creates the momentum buffers
    torch.cuda.current_stream().wait_stream(side)
    ref = copy.deepcopy(b)
    oref = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    oref.load_state_dict(copy.deepcopy(ob.state_dict()))
    eref = copy.deepcopy(eb)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        sgd_clip_ema_step(ob, b, eb, 1.0, 0.99, 50, skip_flag=loss_is_finite_flag(loss))
    for _ in range(3):
        graph.replay()
    for p, q in zip(ref.parameters(), b.parameters()):
        p.grad = q.grad.clone()
    for _ in range(3):
        for p, q in zip(ref.parameters(), b.parameters()):
            p.grad = q.grad.clone()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        oref.step()
        reference_ema(ref, eref, 0.99, 50)
    for p, q in zip(ref.parameters(), b.parameters()):
        assert close(q.data, p.data, 5e-6)
    for p, q in zip(eref.parameters(), eb.parameters()):
        assert close(q.data, p.data, 5e-6)
