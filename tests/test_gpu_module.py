"""The drop-in modules inside an autograd graph, the way the step loop uses them
(train_DyCON_BraTS19.py:346-365), and CUDA-graph capturability (no host syncs)."""
import numpy as np
import pytest
import torch

from conftest import normwise
from oracle import torch_port

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


def test_combined_step_matches_oracle_autograd():
    from dycon_paper_replication_b200 import dycon_losses
    from dycon_paper_replication_b200.synthetic import make_inputs
    inp = make_inputs("tiny", dim=32, mask_kind="bernoulli")
    dev = inp.to("cuda")
    s = dev.s_logits.requires_grad_(True)
    raw = dev.feat.detach().clone().requires_grad_(True)            # a leaf *before* an op, like the network
    feat = raw * 1.0
    fecl = dycon_losses.FeCLoss(device="cuda", temperature=0.6, gamma=2.0, use_focal=True, rampup_epochs=1500,
                                precision="fp32")
    uncl = dycon_losses.UnCLoss()
    beta = dycon_losses.adaptive_beta(epoch=3, total_epochs=30, max_beta=5.0, min_beta=0.5)
    f_loss = fecl(feat=feat, mask=dev.mask, teacher_feat=dev.teacher, gambling_uncertainty=None, epoch=3)
    u_loss = uncl(s, dev.t_logits, beta)
    loss = 0.5 * (f_loss + u_loss)
    assert loss.dim() == 0 and loss.dtype == torch.float32
    assert not (torch.isnan(loss) or torch.isinf(loss))              # the caller's guard works on it
    loss.backward()

    so = inp.s_logits.clone().requires_grad_(True)
    fo = inp.feat.clone().requires_grad_(True)
    ref = 0.5 * (torch_port.fecl_loss(fo, inp.mask, inp.teacher, None, 3, temperature=0.6, gamma=2.0, use_focal=True,
                                      rampup_epochs=1500) + torch_port.uncl_loss(so, inp.t_logits, beta))
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert normwise(s.grad.cpu().numpy(), so.grad.numpy()) <= 1e-5
    assert normwise(raw.grad.cpu().numpy(), fo.grad.numpy()) <= 1e-5


def test_cuda_graph_capture_and_replay():
    from dycon_paper_replication_b200 import FeCLoss, UnCLoss
    from dycon_paper_replication_b200.synthetic import make_inputs
    inp = make_inputs("tiny", dim=32, mask_kind="bernoulli").to("cuda")
    s = inp.s_logits.clone().requires_grad_(True)
    f = inp.feat.clone().requires_grad_(True)
    fecl, uncl = FeCLoss("cuda", use_focal=True, rampup_epochs=1500, precision="fp32"), UnCLoss()

    def step():
        s.grad = None
        f.grad = None
        loss = fecl(f, inp.mask, inp.teacher, None, 100) + uncl(s, inp.t_logits, 1.58)
        loss.backward()
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            eager = step()
    torch.cuda.current_stream().wait_stream(side)
    eager_loss, eager_gs, eager_gf = eager.item(), s.grad.clone(), f.grad.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        captured = step()
    gs, gf = s.grad, f.grad
    gs.zero_(); gf.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert captured.item() == eager_loss
    assert torch.equal(gs, eager_gs) and torch.equal(gf, eager_gf)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("layout", ["caller", "contiguous", "sliced"])
def test_gradient_is_written_in_the_layout_of_feat(layout, precision):
    """feat arrives with strides (D*N, 1, N) (train_DyCON_BraTS19.py:316-323): the gradient must come back in
    the same layout (no re-layout copy in autograd) and hold the same values as for a contiguous feat."""
    from dycon_paper_replication_b200 import FeCLoss
    from dycon_paper_replication_b200.synthetic import make_inputs
    inp = make_inputs("tiny", dim=64, mask_kind="bernoulli").to("cuda")
    fecl = FeCLoss("cuda", use_focal=True, rampup_epochs=1500, precision=precision)
    base = inp.feat.detach().contiguous().clone().requires_grad_(True)
    fecl(base, inp.mask, inp.teacher, None, 100).backward()
    if layout == "caller":
        f = inp.feat.detach().transpose(1, 2).contiguous().transpose(1, 2).requires_grad_(True)
        assert f.stride() == (f.shape[1] * f.shape[2], 1, f.shape[1])
    elif layout == "contiguous":
        f = inp.feat.detach().contiguous().clone().requires_grad_(True)
    else:   # neither rows nor columns dense: falls back to a contiguous gradient
        wide = torch.zeros(f_shape := (inp.feat.shape[0], inp.feat.shape[1], 2 * inp.feat.shape[2]), device="cuda")
        wide[..., ::2] = inp.feat
        f = wide[..., ::2].detach().requires_grad_(True)
        assert f_shape[2] == 2 * f.shape[2]
    fecl(f, inp.mask, inp.teacher, None, 100).backward()
    if layout != "sliced":
        assert f.grad.stride() == f.stride()
    assert torch.equal(f.grad.contiguous(), base.grad)
