"""B200-native DyCON loss hot path: UnCL, FeCL (forward + backward) and the mean-teacher EMA.

``dycon_paper_replication_b200.dycon_losses`` mirrors the reference module
``code/utils/dycon_losses.py``; ``dycon_paper_replication_b200.losses`` carries the legacy
``FeCLoss(device, temperature)`` of ``code/utils/losses.py:221-250``.
"""
from . import dycon_losses
from .dycon_losses import (FeCLoss, UnCLoss, adaptive_beta, gambling_softmax, sigmoid_rampup,
                           update_ema_variables)

__all__ = ["dycon_losses", "UnCLoss", "FeCLoss", "adaptive_beta", "sigmoid_rampup", "gambling_softmax",
           "update_ema_variables"]
