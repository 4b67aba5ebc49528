"""Fused step losses (UnCL + cross entropy + Dice + consistency from one pass over the logits) vs the
unmodified reference functions (tests/golden/step.npz, oracle/make_golden_step.py) and the fp64 port at full size.
fp32 path: rtol 1e-5 on every loss, max|dg| <= 1e-5 max|g| on the gradient of the weighted sum."""
import numpy as np
import pytest
import torch

from conftest import load_golden, normwise
from oracle import torch_port

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]
STEP = load_golden("step")


def run(s, t, label, lb, beta, w):
    from dycon_paper_replication_b200 import StepLosses
    x = s.cuda().requires_grad_(True)
    out = StepLosses()(x, t.cuda(), label.cuda(), lb, beta)
    sum(float(wk) * o for wk, o in zip(w, out)).backward()
    return np.array([o.item() for o in out], np.float64), x.grad.cpu().numpy()


@pytest.mark.parametrize("case", sorted(STEP))
def test_golden(case):
    rec = STEP[case]
    got, grad = run(torch.from_numpy(rec["s"]), torch.from_numpy(rec["t"]), torch.from_numpy(rec["label"]),
                    int(rec["labeled_bs"]), float(rec["beta"]), rec["weights"])
    assert np.all(np.abs(got - rec["losses64"]) <= 1e-5 * np.abs(rec["losses64"])), (got, rec["losses64"])
    assert normwise(grad, rec["grad64"]) <= 1e-5


def test_full_brats19_shape_vs_fp64_port():
    from dycon_paper_replication_b200.synthetic import make_inputs
    inp = make_inputs("brats19", batch=4)
    g = torch.Generator().manual_seed(3)
    label = (torch.rand((4,) + tuple(inp.s_logits.shape[2:]), generator=g) < 0.1).long()
    w = (0.5, 1.0, 1.0, 0.2)
    got, grad = run(inp.s_logits, inp.t_logits, label, 2, 1.58, w)
    s = inp.s_logits.double().requires_grad_(True)
    out = torch_port.step_losses(s, inp.t_logits.double(), label, 2, 1.58)
    sum(wk * o for wk, o in zip(w, out)).backward()
    ref = np.array([float(o.detach()) for o in out])
    assert np.all(np.abs(got - ref) <= 1e-5 * np.abs(ref)), (got, ref)
    assert normwise(grad, s.grad.numpy()) <= 1e-5


def test_each_output_has_its_own_gradient_and_matches_the_unfused_uncl():
    """Backward with one upstream weight at a time (the others zero), and u_loss equal to the UnCLoss kernels."""
    from dycon_paper_replication_b200 import StepLosses, UnCLoss
    rec = STEP["half_labelled"]
    s, t, label = (torch.from_numpy(rec[k]) for k in ("s", "t", "label"))
    lb, beta = int(rec["labeled_bs"]), float(rec["beta"])
    for k in range(4):
        w = [0.0] * 4
        w[k] = 1.0
        _, grad = run(s, t, label, lb, beta, w)
        x = s.double().requires_grad_(True)
        out = torch_port.step_losses(x, t.double(), label, lb, beta)
        out[k].backward()
        assert normwise(grad, x.grad.numpy()) <= 1e-5, k
    x = s.cuda().requires_grad_(True)
    u = UnCLoss()(x, t.cuda(), beta)
    fused = StepLosses()(s.cuda(), t.cuda(), label.cuda(), lb, beta)[0]
    assert abs(u.item() - fused.item()) <= 2e-6 * abs(u.item())


def test_other_class_counts_compose_the_unfused_path():
    from dycon_paper_replication_b200 import StepLosses
    g = torch.Generator().manual_seed(9)
    s = torch.randn(3, 3, 4, 4, 4, generator=g)
    t = s + 0.5 * torch.randn(3, 3, 4, 4, 4, generator=g)
    label = torch.randint(0, 3, (3, 4, 4, 4), generator=g)
    x = s.cuda().requires_grad_(True)
    out = StepLosses()(x, t.cuda(), label.cuda(), 2, 0.8)
    sum(out).backward()
    y = s.double().requires_grad_(True)
    ref = torch_port.step_losses(y, t.double(), label, 2, 0.8)
    sum(ref).backward()
    for a, b in zip(out, ref):
        assert abs(a.item() - b.item()) <= 1e-5 * abs(b.item())
    assert normwise(x.grad.cpu().numpy(), y.grad.numpy()) <= 1e-5


def test_bad_arguments_raise():
    from dycon_paper_replication_b200 import StepLosses
    s = torch.randn(2, 2, 4, 4, 4)
    lab = torch.zeros(2, 4, 4, 4, dtype=torch.long)
    with pytest.raises(RuntimeError):
        StepLosses()(s, s, lab, 1, 1.0)                       # CPU tensors: no fallback
    with pytest.raises(TypeError):
        StepLosses()(s.cuda(), s.cuda(), lab.float().cuda(), 1, 1.0)
    with pytest.raises(ValueError):
        StepLosses()(s.cuda(), s.cuda(), lab.cuda(), 3, 1.0)
