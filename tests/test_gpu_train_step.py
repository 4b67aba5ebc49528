"""Drop-in check inside a PyTorch mean-teacher training loop (BASELINE config 3 in miniature).

The step mirrors code/train_DyCON_BraTS19.py:298-374: student forward, no-grad teacher forward on a noised
input, supervised CE on the labelled half, embeddings = normalize(view -> transpose) (the (D*N, 1, N)-strided
layout), mask = avg_pool3d(label) > 0.5, FeCL + UnCL on ALL samples, backward, SGD step, EMA teacher update.
The 3D network is a small stock-PyTorch conv net (the reference's UNet3D is context, not product, and
/root/reference does not exist on the GPU box).  The same loop is run with this package's modules and with the
oracle's op-for-op port of the reference losses / EMA loop (on the same GPU); the loss trajectories and the final
student / teacher parameters must agree."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import torch_port

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

CTOR = dict(temperature=0.6, gamma=2.0, use_focal=True, rampup_epochs=1500)


class TinyNet(nn.Module):
    """(B,1,32,32,32) -> logits (B,2,32,32,32), features (B,32,4,4,4): the two outputs the losses consume."""

    def __init__(self, dim=32):
        super().__init__()
        self.enc = nn.Sequential(nn.Conv3d(1, 8, 3, padding=1), nn.ReLU(), nn.Conv3d(8, 8, 3, padding=1), nn.ReLU())
        self.seg = nn.Conv3d(8, 2, 1)
        self.proj = nn.Conv3d(8, dim, 1)
        # a fixed positional code keeps the embeddings of different voxels apart (an untrained projection maps
        # every voxel to almost the same direction, where 1 - cos rounds to <= 0 and the reference itself is NaN)
        self.register_buffer("pos", torch.randn(1, dim, 4, 4, 4, generator=torch.Generator().manual_seed(3)))

    def forward(self, x):
        h = self.enc(x)
        return self.seg(h), 4.0 * self.proj(F.avg_pool3d(h, 8, 8)) + self.pos


def embed(features):
    b, c = features.shape[:2]
    return F.normalize(features.view(b, c, -1).transpose(1, 2), dim=-1)      # strides (C*N, 1, N)


def run_loop(kind, precision, steps=4):
    from dycon_paper_replication_b200 import dycon_losses
    torch.manual_seed(1337)
    dev = torch.device("cuda")
    student = TinyNet().to(dev)
    teacher = copy.deepcopy(student)
    for p in teacher.parameters():
        p.detach_()
    opt = torch.optim.SGD(student.parameters(), lr=0.05, momentum=0.9)
    gen = torch.Generator(device="cpu").manual_seed(7)
    x = torch.randn(4, 1, 32, 32, 32, generator=gen).to(dev)
    label = torch.zeros(4, 32, 32, 32, dtype=torch.long)
    label[:, 8:24, 4:20, 10:30] = 1
    label = label.to(dev)
    noise = (0.1 * torch.randn(4, 1, 32, 32, 32, generator=gen)).clamp(-0.2, 0.2).to(dev)
    if kind == "ours":
        fecl = dycon_losses.FeCLoss(device=dev, precision=precision, **CTOR)
        uncl = dycon_losses.UnCLoss()
    losses = []
    for it in range(steps):
        epoch = 100 + it
        beta = dycon_losses.adaptive_beta(epoch=epoch, total_epochs=300, max_beta=5.0, min_beta=0.5)
        s_logits, s_feat = student(x)
        with torch.no_grad():
            t_logits, t_feat = teacher(x + noise)
        ce = F.cross_entropy(s_logits[:2], label[:2])
        s_emb, t_emb = embed(s_feat), embed(t_feat)
        mask = (F.avg_pool3d(label.float().unsqueeze(1), 8, 8) > 0.5).float().reshape(4, -1).unsqueeze(1)
        if kind == "ours":
            f_loss = fecl(feat=s_emb, mask=mask, teacher_feat=t_emb, gambling_uncertainty=None, epoch=epoch)
            u_loss = uncl(s_logits, t_logits, beta)
        else:
            f_loss = torch_port.fecl_loss(s_emb, mask, t_emb, None, epoch, **CTOR)
            u_loss = torch_port.uncl_loss(s_logits, t_logits, beta)
        loss = ce + 0.5 * (f_loss + u_loss)
        assert torch.isfinite(loss)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if kind == "ours":
            dycon_losses.update_ema_variables(student, teacher, 0.99, it)
        else:
            torch_port.ema_update([p.data for p in teacher.parameters()], [p.data for p in student.parameters()], 0.99, it)
        losses.append((f_loss.item(), u_loss.item(), loss.item()))
    params = torch.cat([p.detach().flatten() for p in list(student.parameters()) + list(teacher.parameters())])
    return losses, params


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("fp16", 5e-3)])
def test_training_loop_matches_the_reference_losses(precision, tol):
    ours, p_ours = run_loop("ours", precision)
    ref, p_ref = run_loop("oracle", precision)
    for (fo, uo, lo), (fr, ur, lr) in zip(ours, ref):
        assert abs(fo - fr) <= tol * abs(fr), (fo, fr)
        assert abs(uo - ur) <= tol * abs(ur), (uo, ur)
        assert abs(lo - lr) <= tol * abs(lr), (lo, lr)
    assert ((p_ours - p_ref).abs().max() / p_ref.abs().max()).item() <= 10 * tol
