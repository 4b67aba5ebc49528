"""UnCL CUDA path (through the module -> ctypes -> C ABI) vs the oracle.  Tolerance: the north-star's
fp32 bound, rtol 1e-5 on the loss and max|dg| <= 1e-5 * max|g| on the gradient."""
import numpy as np
import pytest
import torch

from conftest import load_golden, normwise
from oracle import closed_form, torch_port

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

TOL = 1e-5
UNCL = load_golden("uncl")


def run(s, t, beta, go=0.5):
    from dycon_paper_replication_b200 import UnCLoss
    s = s.cuda().requires_grad_(True)
    loss = UnCLoss()(s, t.cuda(), beta)
    (loss * go).backward()
    return loss.detach().cpu().double().item(), s.grad.detach().cpu().numpy()


@pytest.mark.parametrize("case", sorted(UNCL))
def test_golden(case):
    rec = UNCL[case]
    loss, grad = run(torch.from_numpy(rec["s"]), torch.from_numpy(rec["t"]), float(rec["beta"]), float(rec["go"]))
    assert abs(loss - float(rec["loss64"])) <= TOL * abs(float(rec["loss64"]))
    assert abs(loss - float(rec["loss32"])) <= TOL * abs(float(rec["loss32"]))
    assert normwise(grad, rec["grad64"]) <= TOL
    assert normwise(grad, rec["grad32"]) <= TOL


@pytest.mark.parametrize("shape,beta", [((2, 2, 32, 32, 32), 5.0), ((3, 2, 17, 9, 5), 1.58),   # odd V: scalar path
                                        ((2, 3, 16, 16, 8), 0.5), ((1, 5, 7, 7, 7), 2.0), ((2, 2, 4, 4, 4), 0.8)])
def test_seeded_vs_oracle(shape, beta):
    g = torch.Generator().manual_seed(1337)
    s = 2 * torch.randn(shape, generator=g)
    t = s + 0.5 * torch.randn(shape, generator=g)
    loss, grad = run(s, t, beta)
    ref = closed_form.uncl(s.numpy(), t.numpy(), beta, go=0.5)
    assert abs(loss - ref["loss"]) <= TOL * abs(ref["loss"])
    assert normwise(grad, ref["grad"]) <= TOL
    l32, g32 = torch_port.uncl_fwd_bwd(s, t, beta, 0.5)
    assert abs(loss - float(l32)) <= TOL * abs(float(l32))
    assert normwise(grad, g32.numpy()) <= TOL


@pytest.mark.parametrize("beta", [5.0, 1.58, 0.5])
def test_brats19_full_size(beta):
    """BASELINE config 1/2: B=4, C=2, 96^3."""
    from dycon_paper_replication_b200.synthetic import make_logits
    g = torch.Generator().manual_seed(1337)
    s, t = make_logits(4, 2, (96, 96, 96), g)
    loss, grad = run(s, t, beta)
    ref = closed_form.uncl(s.numpy(), t.numpy(), beta, go=0.5)
    assert abs(loss - ref["loss"]) <= TOL * abs(ref["loss"])
    assert normwise(grad, ref["grad"]) <= TOL
    # size-independent properties: channel gradients cancel per voxel (softmax Jacobian) ...
    assert np.abs(grad.sum(axis=1)).max() <= 1e-6 * np.abs(grad).max()
    # ... and the loss is a mean: evaluating two halves of the batch separately averages to the whole
    la, _ = run(s[:2], t[:2], beta)
    lb, _ = run(s[2:], t[2:], beta)
    assert abs(0.5 * (la + lb) - loss) <= 2e-6 * abs(loss)


def test_bitwise_deterministic():
    g = torch.Generator().manual_seed(3)
    s = 2 * torch.randn(2, 2, 40, 40, 40, generator=g)
    t = s + 0.5 * torch.randn(s.shape, generator=g)
    a = run(s, t, 1.58)
    b = run(s, t, 1.58)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


def test_nan_propagates_to_loss():
    s = torch.zeros(1, 2, 8, 8, 8)
    s[0, 1, 3, 3, 3] = float("nan")
    loss, _ = run(s, torch.zeros_like(s), 1.0)
    assert np.isnan(loss)        # the caller's guard (train_DyCON_BraTS19.py:360) must still fire


def test_noncontiguous_and_misaligned_inputs():
    g = torch.Generator().manual_seed(5)
    base = 2 * torch.randn(2, 2, 8, 8, 9, generator=g).cuda()
    s = base[..., 1:]                      # non-contiguous view, odd inner size
    t = (base + 0.3)[..., 1:]
    from dycon_paper_replication_b200 import UnCLoss
    sl = s.detach().clone().requires_grad_(True)
    loss = UnCLoss()(sl, t, 2.0)
    loss.backward()
    ref = closed_form.uncl(s.cpu().numpy(), t.cpu().numpy(), 2.0)
    assert abs(loss.item() - ref["loss"]) <= TOL * abs(ref["loss"])
    assert normwise(sl.grad.cpu().numpy(), ref["grad"]) <= TOL


def test_teacher_requires_grad_raises_and_cpu_raises():
    from dycon_paper_replication_b200 import UnCLoss
    s = torch.randn(1, 2, 4, 4, 4, device="cuda")
    with pytest.raises(RuntimeError):
        UnCLoss()(s, s.clone().requires_grad_(True), 1.0)
    with pytest.raises(RuntimeError):
        UnCLoss()(s.cpu(), s.cpu(), 1.0)
