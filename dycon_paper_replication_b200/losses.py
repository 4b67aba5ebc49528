"""Legacy ``losses.FeCLoss(device, temperature).forward(feat, mask)``
(reference: code/utils/losses.py:221-250) -- the use_focal=False, teacher-free special case of
``dycon_losses.FeCLoss``, served by the same kernels.  The other members
of the reference's ``losses.py`` (dice, CE, KL ...) are stock PyTorch terms and out of scope."""
from __future__ import annotations

import torch.nn as nn

from . import dycon_losses


class FeCLoss(nn.Module):
    def __init__(self, device, temperature=0.6, *, precision=None):
        super().__init__()
        self.device = device
        self.temperature = temperature
        self._impl = dycon_losses.FeCLoss(device, temperature=temperature, use_focal=False, precision=precision)

    def forward(self, feat, mask):
        self._impl.temperature = self.temperature
        return self._impl(feat, mask)
