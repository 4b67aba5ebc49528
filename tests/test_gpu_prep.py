"""Caller-side preparation fused into the FeCL boundary: FeCLoss.from_features and its kernels
(pooled mask, row norms folded into the operand staging, Jacobian of the normalisation) against the reference chain
of code/train_DyCON_BraTS19.py:316-350 -- fixtures tests/golden/prep.npz (unmodified FeCLoss behind the restated
preparation lines, oracle/make_golden_step.py) and the fp64 port at the BraTS19 / ISLES22 feature shapes."""
import numpy as np
import pytest
import torch

from conftest import load_golden, normwise
from oracle import torch_port

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]
PREP = load_golden("prep")
TOL = {"fp32": 1e-5, "fp16": 2e-3}
CTOR = dict(temperature=0.6, gamma=2.0, use_focal=True, rampup_epochs=1500)


@pytest.mark.parametrize("dtype", [torch.int64, torch.uint8, torch.float32])
@pytest.mark.parametrize("case", sorted(PREP))
def test_pooled_mask_is_exact(case, dtype):
    from dycon_paper_replication_b200.dycon_losses import pooled_label_mask
    rec = PREP[case]
    label = torch.from_numpy(rec["label"]).to(dtype).cuda()
    got = pooled_label_mask(label, rec["features"].shape[2:]).cpu().numpy()
    assert np.array_equal(got, rec["mask"].reshape(got.shape))


def test_pooled_mask_at_the_training_shapes():
    from dycon_paper_replication_b200.dycon_losses import pooled_label_mask
    g = torch.Generator().manual_seed(4)
    for shape, grid in (((4, 96, 96, 96), (12, 12, 12)), ((2, 96, 96, 64), (24, 24, 16)), ((2, 112, 112, 96), (14, 14, 12))):
        coarse = (torch.rand((shape[0],) + grid, generator=g) < 0.3)
        k = [s // q for s, q in zip(shape[1:], grid)]
        label = coarse.repeat_interleave(k[0], 1).repeat_interleave(k[1], 2).repeat_interleave(k[2], 3).long()
        label = label * (torch.rand(shape, generator=g) < 0.75).long()        # windows between 0 and ~0.75 full
        want = torch_port.prep_mask(label, grid).reshape(shape[0], -1)
        got = pooled_label_mask(label.cuda(), grid).cpu()
        assert torch.equal(got, want)


@pytest.mark.parametrize("mode", ["fp32", "fp16"])
@pytest.mark.parametrize("case", sorted(PREP))
def test_golden(case, mode):
    from dycon_paper_replication_b200 import FeCLoss
    rec = PREP[case]
    x = torch.from_numpy(rec["features"]).cuda().requires_grad_(True)
    t = torch.from_numpy(rec["ema_features"]).cuda() if int(rec["teacher"]) else None
    crit = FeCLoss("cuda", precision=mode, **CTOR)
    loss = crit.from_features(x, torch.from_numpy(rec["label"]).cuda(), t, None, int(rec["epoch"]))
    (float(rec["go"]) * loss).backward()
    assert abs(loss.item() - float(rec["loss64"])) <= TOL[mode] * abs(float(rec["loss64"]))
    # tiny fixtures: a threshold flip of the teacher term would be a large step, so compare with the flip-free bound only
    # where the fp32 mode is exact; the 16-bit mode is checked against its own unfused path below
    if mode == "fp32":
        assert normwise(x.grad.cpu().numpy(), rec["grad64"]) <= 5e-5


@pytest.mark.parametrize("mode", ["fp32", "fp16"])
@pytest.mark.parametrize("shape", ["brats19", "isles22"])
def test_fused_equals_the_unfused_chain_on_device(shape, mode):
    """from_features against the SAME kernels behind the reference's own PyTorch preparation (F.normalize, avg_pool3d)
    -- the normalisation folded into the operand staging and its Jacobian kernel must reproduce autograd's chain."""
    from dycon_paper_replication_b200 import FeCLoss
    from dycon_paper_replication_b200.synthetic import SHAPES, feature_grid, make_blob_labels
    g = torch.Generator().manual_seed(8)
    spatial = SHAPES[shape][0]
    grid = feature_grid(shape)
    b, c = 2, 256
    labels = make_blob_labels(b, spatial, g).long()
    coarse = torch_port.prep_mask(labels, grid).reshape(b, 1, *grid)
    x = (torch.randn(1, c, 1, 1, 1, generator=g) + 0.8 * torch.randn(b, c, *grid, generator=g) + 0.6 * coarse).cuda()
    t = (x.cpu() + 0.1 * torch.randn(b, c, *grid, generator=g)).cuda()
    crit = FeCLoss("cuda", precision=mode, **CTOR)
    xa = x.clone().requires_grad_(True)
    la = crit.from_features(xa, labels.cuda(), t, None, 100)
    (0.5 * la).backward()
    xb = x.clone().requires_grad_(True)
    emb = torch_port.prep_embeddings(xb)
    lb = crit(feat=emb, mask=torch_port.prep_mask(labels, grid).cuda(), teacher_feat=torch_port.prep_embeddings(t),
              gambling_uncertainty=None, epoch=100)
    (0.5 * lb).backward()
    tol = 2e-6 if mode == "fp32" else 5e-4        # 16-bit: x * (1/|x|) and x / |x| round to different halves now and then
    assert abs(la.item() - lb.item()) <= tol * abs(lb.item())
    assert normwise(xa.grad.cpu().numpy(), xb.grad.cpu().numpy()) <= (1e-5 if mode == "fp32" else 2e-3)


def test_brats19_shape_vs_fp64_port():
    from dycon_paper_replication_b200 import FeCLoss
    from dycon_paper_replication_b200.synthetic import make_blob_labels
    g = torch.Generator().manual_seed(12)
    b, c, grid = 1, 64, (12, 12, 12)
    labels = make_blob_labels(b, (96, 96, 96), g).long()
    coarse = torch_port.prep_mask(labels, grid).reshape(b, 1, *grid)
    x = torch.randn(1, c, 1, 1, 1, generator=g) + 0.8 * torch.randn(b, c, *grid, generator=g) + 0.6 * coarse
    xg = x.cuda().requires_grad_(True)
    loss = FeCLoss("cuda", precision="fp32", **CTOR).from_features(xg, labels.cuda(), None, None, 100)
    (0.5 * loss).backward()
    xr = x.double().requires_grad_(True)
    torch.set_default_dtype(torch.float64)
    try:
        ref = torch_port.fecl_from_features(xr, labels, None, 100, **CTOR)
    finally:
        torch.set_default_dtype(torch.float32)
    (0.5 * ref).backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert normwise(xg.grad.cpu().numpy(), xr.grad.numpy()) <= 1e-5
