"""Loss functions of the optional training step (mirror of reference src/utils/losses.py).

The reference takes its losses from ``segmentation_models_pytorch`` (``smp.losses.DiceLoss(mode, smooth)``,
reference src/utils/losses.py:18-19) and keeps a ``CombinedLoss`` (:33-52) that it never constructs; config 5 of
BASELINE.json ("Dice+BCE") is that class over the two parts with the weights of reference src/configs/config.py:61-62.
smp is not a dependency here: ``DiceLoss`` below restates smp's binary Dice loss (from_logits, batch-global sums over
dims (0, 2), ``smooth`` in numerator and denominator, ``eps = 1e-7`` clamp, classes without positives masked out).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


class DiceLoss(nn.Module):
    """``smp.losses.DiceLoss(mode='binary', from_logits=True, smooth=smooth, eps=1e-7, log_loss=False)``."""

    def __init__(self, mode: str = "binary", smooth: float = 0.0, eps: float = 1e-7):
        super().__init__()
        if mode != "binary":
            raise NotImplementedError("the B200 path trains the 1-class mask model: mode='binary' only")
        self.smooth, self.eps = float(smooth), float(eps)

    def forward(self, y_pred: torch.Tensor, y_true: torch.Tensor) -> torch.Tensor:
        bs = y_true.size(0)
        p = F.logsigmoid(y_pred.float()).exp().view(bs, 1, -1)
        t = y_true.view(bs, 1, -1).type_as(p)
        inter = torch.sum(p * t, dim=(0, 2))
        card = torch.sum(p + t, dim=(0, 2))
        dice = (2.0 * inter + self.smooth) / (card + self.smooth).clamp_min(self.eps)
        loss = (1.0 - dice) * (t.sum((0, 2)) > 0).to(p.dtype)
        return loss.mean()


class BCEWithLogits(nn.Module):
    """``nn.BCEWithLogitsLoss()`` on a float view of the {0,1} target (the reference's dataset yields long masks,
    src/utils/dataset.py:119-122, unsqueezed at src/train.py:92-93)."""

    def forward(self, y_pred, y_true):
        return F.binary_cross_entropy_with_logits(y_pred.float(), y_true.type_as(y_pred).float())


def get_loss_function(cfg) -> nn.Module:
    """Mirror of reference src/utils/losses.py:11-31 for the losses this path implements."""
    name = cfg.LOSS.NAME
    mode = getattr(cfg.LOSS, "MODE", "binary")
    smooth = getattr(cfg.LOSS, "SMOOTH", cfg.LOSS.DICE_SMOOTH)
    if name == "DiceLoss":
        return DiceLoss(mode=mode, smooth=smooth)
    if name == "BCEWithLogitsLoss":
        return BCEWithLogits()
    if name in ("JaccardLoss", "FocalLoss", "TverskyLoss", "LovaszLoss"):
        raise NotImplementedError(f"{name}: the B200 training step implements DiceLoss, BCEWithLogitsLoss and their "
                                  "combination (BASELINE config 5)")
    raise ValueError(f"不支持的损失函数: {name}")


class CombinedLoss(nn.Module):
    """reference src/utils/losses.py:33-52."""

    def __init__(self, losses: List[nn.Module], weights: Optional[List[float]] = None):
        super().__init__()
        self.losses = nn.ModuleList(losses)
        self.weights = weights if weights else [1.0] * len(losses)

    def forward(self, pred, target):
        total = 0
        for fn, w in zip(self.losses, self.weights):
            total = total + w * fn(pred, target)
        return total


def dice_bce_from_config(cfg) -> CombinedLoss:
    """BASELINE config 5: DICE_WEIGHT * Dice(smooth = DICE_SMOOTH) + BCE_WEIGHT * BCEWithLogits
    (reference src/configs/config.py:57-62)."""
    return CombinedLoss([DiceLoss("binary", smooth=cfg.LOSS.DICE_SMOOTH), BCEWithLogits()],
                        [cfg.LOSS.DICE_WEIGHT, cfg.LOSS.BCE_WEIGHT])
