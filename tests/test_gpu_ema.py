"""Multi-tensor EMA kernel vs the reference per-parameter loop (train_DyCON_BraTS19.py:155-164)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import closed_form, torch_port

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


class Bag(torch.nn.Module):
    def __init__(self, shapes, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes])
        self.register_buffer("running", torch.full((3,), float(seed)))   # buffers must NOT be averaged


def real_shapes():
    from dycon_paper_replication_b200.synthetic import unet3d_param_shapes
    return unet3d_param_shapes()


@pytest.mark.parametrize("step", [0, 1, 50, 10000])
def test_unet3d_parameter_set_matches_reference_loop(step):
    from dycon_paper_replication_b200 import update_ema_variables
    shapes = real_shapes()
    student, teacher = Bag(shapes, 1).cuda(), Bag(shapes, 2).cuda()
    want = Bag(shapes, 2).cuda()
    torch_port.ema_update([p.data for p in want.parameters()], [p.data for p in student.parameters()], 0.99, step)
    update_ema_variables(student, teacher, 0.99, step)
    torch.cuda.synchronize()
    exact = 0
    for got, ref in zip(teacher.parameters(), want.parameters()):
        g, r = got.detach().cpu().numpy(), ref.detach().cpu().numpy()
        assert np.all(np.abs(g - r) <= np.spacing(np.abs(r)))      # <= 1 ulp of the CUDA eager loop
        exact += int(np.array_equal(g, r))
    assert exact == len(shapes), f"only {exact}/{len(shapes)} tensors bit-identical to the eager CUDA loop"
    assert torch.equal(teacher.running, torch.full((3,), 2.0, device="cuda"))
    if step == 0:                                                    # alpha = 0: teacher := student
        for got, src in zip(teacher.parameters(), student.parameters()):
            assert torch.equal(got, src)


def test_golden_reference_loop_cpu_fixture():
    from dycon_paper_replication_b200 import update_ema_variables
    z = np.load(GOLDEN + "/ema.npz")
    sizes = z["sizes"].tolist()
    for step in (0, 1, 50, 10000):
        student, teacher = Bag([(n,) for n in sizes], 0).cuda(), Bag([(n,) for n in sizes], 0).cuda()
        off = 0
        for ps, pt, n in zip(student.parameters(), teacher.parameters(), sizes):
            ps.data.copy_(torch.from_numpy(z[f"step{step}_student"][off:off + n]))
            pt.data.copy_(torch.from_numpy(z[f"step{step}_before"][off:off + n]))
            off += n
        update_ema_variables(student, teacher, 0.99, step)
        got = np.concatenate([p.detach().cpu().numpy() for p in teacher.parameters()])
        want = z[f"step{step}_after"]
        assert np.all(np.abs(got - want) <= np.spacing(np.abs(want)))


def test_module_wrapped_misaligned_and_empty():
    from dycon_paper_replication_b200 import update_ema_variables
    shapes = [(5,), (4099,), (1,), (3, 7, 11)]
    student, teacher = Bag(shapes, 3).cuda(), Bag(shapes, 4).cuda()
    # misaligned views: parameters that start 4 bytes into an allocation
    for m in (student, teacher):
        big = torch.randn(8200, device="cuda")
        m.ps[1].data = big[1:4100]
    want = [p.detach().clone() for p in teacher.parameters()]
    torch_port.ema_update(want, [p.data for p in student.parameters()], 0.99, 7)

    class Wrap(torch.nn.Module):       # DataParallel-style .module wrapper (train_DyCON_BraTS19.py:160-161)
        def __init__(self, m):
            super().__init__()
            self.module = m
    update_ema_variables(Wrap(student), Wrap(teacher), 0.99, 7)
    for got, ref in zip(teacher.parameters(), want):
        assert torch.equal(got.detach(), ref)
    with pytest.raises(RuntimeError):
        update_ema_variables(Bag(shapes, 1), Bag(shapes, 2), 0.99, 1)     # CPU params: no fallback
