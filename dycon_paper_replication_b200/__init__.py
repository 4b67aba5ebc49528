"""B200-native DyCON loss hot path: UnCL, FeCL (forward + backward) and the mean-teacher EMA.

``dycon_paper_replication_b200.dycon_losses`` mirrors the reference module
``code/utils/dycon_losses.py``; ``dycon_paper_replication_b200.losses`` carries the legacy
``FeCLoss(device, temperature)`` of ``code/utils/losses.py:221-250``.
"""
from . import dycon_losses
from .dycon_losses import (FeCLoss, StepLosses, UnCLoss, adaptive_beta, gambling_softmax, loss_is_finite_flag,
                           sgd_clip_ema_step, sigmoid_rampup, update_ema_variables)

__all__ = ["dycon_losses", "UnCLoss", "FeCLoss", "StepLosses", "adaptive_beta", "sigmoid_rampup", "gambling_softmax",
           "update_ema_variables", "sgd_clip_ema_step", "loss_is_finite_flag"]
