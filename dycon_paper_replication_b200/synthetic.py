"""Seeded synthetic inputs of the shapes the DyCON step loop feeds the losses.

Shapes and recipes follow SURVEY.md section 8(d); they reproduce the tensor
layouts of the reference caller (code/train_DyCON_BraTS19.py:316-330): the
embeddings are built through ``view -> transpose -> F.normalize`` from a
``(B, D, h, w, d)`` feature volume so they carry the same N-contiguous strides
``(D*N, 1, N)``, and the contrastive mask is ``avg_pool3d(label) > 0.5``.
All tensors are generated on the CPU from a ``torch.Generator`` so a seed gives
the same bits on every box; callers move them to the device.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F

# name -> (patch (H,W,Dz), pooling factor = 4*feature_scaler, default batch)
SHAPES = {
    "brats19": ((96, 96, 96), 8, 4),      # train_DyCON_BraTS19.py:147,37  (BASELINE config 1/2: B=4)
    "pancreas": ((112, 112, 96), 8, 8),   # train_DyCON_Pancreas.py:99,36
    "isles22": ((96, 96, 64), 4, 8),      # train_DyCON_ISLES22.py:70,75   (feature grid x4)
    "tiny": ((16, 16, 16), 4, 2),
}


def feature_grid(shape_name: str):
    (h, w, d), k, _ = SHAPES[shape_name]
    if shape_name == "isles22":
        # scaler 4: bottleneck (H/16) upsampled x4 -> H/4   
        return (h // 4, w // 4, d // 4)
    return (h // k, w // k, d // k)


@dataclass
class LossInputs:
    s_logits: torch.Tensor      # (B,C,H,W,Dz) fp32 contiguous
    t_logits: torch.Tensor      # (B,C,H,W,Dz)
    feat: torch.Tensor          # (B,N,D) fp32, strides (D*N,1,N), L2-normalised
    teacher: torch.Tensor       # (B,N,D) same layout
    mask: torch.Tensor          # (B,1,N) fp32 {0,1}

    @property
    def voxels(self) -> int:
        s = self.s_logits
        return s.shape[0] * s[0, 0].numel()

    def to(self, device, non_blocking=False):
        # as_strided copy keeps the caller's (D*N,1,N) layout on the device
        def mv(x):
            y = torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=device)
            y.copy_(x, non_blocking=non_blocking)
            return y
        return LossInputs(*(mv(getattr(self, k)) for k in
                            ("s_logits", "t_logits", "feat", "teacher", "mask")))


def make_logits(batch, classes, spatial, gen, *, scale=2.0, noise=0.5):
    """Student logits ``scale*randn``; teacher = student + ``noise*randn`` (section 8d)."""
    s = scale * torch.randn(batch, classes, *spatial, generator=gen)
    t = s + noise * torch.randn(s.shape, generator=gen)
    return s.contiguous(), t.contiguous()


def make_blob_labels(batch, spatial, gen, *, lo=0.03, hi=0.15, empty_first=False):
    """One random ellipsoid per sample covering lo..hi of the voxels -> (B,H,W,Dz) {0,1}."""
    h, w, d = spatial
    zz, yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), torch.arange(d), indexing="ij")
    out = torch.zeros(batch, h, w, d)
    for b in range(batch):
        if empty_first and b == 0:
            continue
        frac = lo + (hi - lo) * torch.rand((), generator=gen).item()
        # ellipsoid volume 4/3*pi*a*b*c = frac*h*w*d with random aspect
        asp = 0.7 + 0.6 * torch.rand(3, generator=gen)
        base = (frac * h * w * d * 3.0 / (4.0 * 3.14159265 * asp.prod().item())) ** (1.0 / 3.0)
        ra, rb, rc = (base * asp).tolist()
        cz = h * (0.3 + 0.4 * torch.rand((), generator=gen).item())
        cy = w * (0.3 + 0.4 * torch.rand((), generator=gen).item())
        cx = d * (0.3 + 0.4 * torch.rand((), generator=gen).item())
        inside = ((zz - cz) / ra) ** 2 + ((yy - cy) / rb) ** 2 + ((xx - cx) / rc) ** 2 <= 1.0
        out[b][inside] = 1.0
    return out


def pooled_mask(labels, k):
    """Caller-side mask prep (train_DyCON_BraTS19.py:326-330) -> (B,1,N) float."""
    m = F.avg_pool3d(labels.float().unsqueeze(1), kernel_size=k, stride=k).squeeze(1)
    m = (m > 0.5).float()
    return m.reshape(m.shape[0], -1).unsqueeze(1)


def make_embeddings(mask, grid, dim, gen, *, kind="structured", teacher_noise=0.3):
    """Student / teacher embeddings (B,N,D) with the caller's strides (section 8d).

    structured: x = 1.0*c + 0.6*proto[label] + 1.2*z, z ~ N(0, 1/D) -- positive-pair
    cosine ~0.49, negative cross-similarity ~0.36 at D=256 (most negatives are cross-hard).
    iid: x ~ N(0, I) -- cross term empty at D=256, stresses -log(1-cs) at D=16.
    """
    b, _, n = mask.shape
    lab = mask.reshape(b, n).long()
    if kind == "structured":
        c = F.normalize(torch.randn(dim, generator=gen), dim=0)
        proto = F.normalize(torch.randn(2, dim, generator=gen), dim=-1)
        z = torch.randn(b, n, dim, generator=gen) / dim ** 0.5
        x = 1.0 * c + 0.6 * proto[lab] + 1.2 * z
        tx = x + teacher_noise * torch.randn(b, n, dim, generator=gen) / dim ** 0.5
    elif kind == "iid":
        x = torch.randn(b, n, dim, generator=gen)
        tx = F.normalize(x, dim=-1) + teacher_noise * torch.randn(b, n, dim, generator=gen)
    else:
        raise ValueError(kind)

    def as_caller(v):
        vol = v.transpose(1, 2).contiguous().view(b, dim, *grid)     # network output (B,D,h,w,d)
        emb = torch.transpose(vol.view(b, dim, -1), 1, 2)            # :317-318
        return F.normalize(emb, dim=-1)                              # :319 keeps strides (D*N,1,N)

    return as_caller(x), as_caller(tx)


def make_inputs(shape="brats19", batch=None, dim=256, classes=2, seed=1337, *,
                feat_kind="structured", mask_kind="blob", empty_first=False) -> LossInputs:
    spatial, k, default_b = SHAPES[shape]
    b = batch or default_b
    gen = torch.Generator().manual_seed(seed)          # scripts' default seed, train_DyCON_BraTS19.py:31
    s, t = make_logits(b, classes, spatial, gen)
    grid = feature_grid(shape)
    n = grid[0] * grid[1] * grid[2]
    if mask_kind == "blob":
        labels = make_blob_labels(b, spatial, gen, empty_first=empty_first)
        if shape == "isles22":
            mask = pooled_mask(labels, 4)
        else:
            mask = pooled_mask(labels, k)
    elif mask_kind == "bernoulli":
        mask = (torch.rand(b, 1, n, generator=gen) < 0.1).float()
        if empty_first:
            mask[0] = 0
    else:
        raise ValueError(mask_kind)
    assert mask.shape[-1] == n, (mask.shape, n)
    feat, teacher = make_embeddings(mask, grid, dim, gen, kind=feat_kind)
    return LossInputs(s, t, feat, teacher, mask)


def unet3d_param_shapes():
    """The 48 parameter shapes (6,148,532 fp32 elements) of the reference ``unet_3D``
    with feature_scaler=2 (code/networks/UNet3D_contrastive.py:207-316), enumerated
    in the build container by instantiating the reference model; used as the EMA
    workload."""
    shapes = []
    cin = 1
    for c in (16, 32, 64, 128, 256):                              # encoder + centre (no BN)
        shapes += [(c, cin, 3, 3, 3), (c,), (c, c, 3, 3, 3), (c,)]
        cin = c
    for hi, lo in ((256, 128), (128, 64), (64, 32), (32, 16)):    # decoders
        shapes += [(lo, hi + lo, 3, 3, 3), (lo,), (lo, lo, 3, 3, 3), (lo,)]
    shapes += [(2, 16, 1, 1, 1), (2,), (2, 16, 1, 1, 1), (2,)]    # two 1x1x1 heads
    shapes += [(512, 256, 1, 1, 1), (512,), (512,), (512,),       # projection head conv+BN
               (256, 512, 1, 1, 1), (256,), (256,), (256,)]
    return shapes
