"""``DiceCELoss`` with the reference's semantics (vendored MONAI ``monai/losses/dice.py:640-810``, as constructed at
``3_train.py:72``: ``DiceCELoss(to_onehot_y=True, softmax=True)``): soft Dice over the spatial dims per (sample, class)
plus softmax cross-entropy, both reduced with ``mean``.  One fp32 log-softmax feeds both terms, whatever the dtype of the
logits (bf16 under autocast): the 2 x 4 x 128^3 probabilities are the only full-size temporaries.

Part of BASELINE configs[4] (the training step); plain torch ops - the hot kernels are inside the network.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["DiceCELoss"]


class DiceCELoss(nn.Module):
    def __init__(self, include_background: bool = True, to_onehot_y: bool = False, sigmoid: bool = False,
                 softmax: bool = False, other_act=None, squared_pred: bool = False, jaccard: bool = False,
                 reduction: str = "mean", smooth_nr: float = 1e-5, smooth_dr: float = 1e-5, batch: bool = False,
                 weight=None, lambda_dice: float = 1.0, lambda_ce: float = 1.0) -> None:
        super().__init__()
        if sigmoid or other_act is not None or weight is not None:
            raise NotImplementedError("the path uses DiceCELoss(to_onehot_y=True, softmax=True); sigmoid / other_act / "
                                      "class weights are not re-hosted")
        if reduction not in ("mean", "sum"):
            raise ValueError(f"reduction must be 'mean' or 'sum', got {reduction!r}")
        if lambda_dice < 0.0:
            raise ValueError("lambda_dice should be no less than 0.0.")
        if lambda_ce < 0.0:
            raise ValueError("lambda_ce should be no less than 0.0.")
        self.include_background = include_background
        self.to_onehot_y = to_onehot_y
        self.softmax = softmax
        self.squared_pred = squared_pred
        self.jaccard = jaccard
        self.reduction = reduction
        self.smooth_nr = float(smooth_nr)
        self.smooth_dr = float(smooth_dr)
        self.batch = batch
        self.lambda_dice = lambda_dice
        self.lambda_ce = lambda_ce

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if input.dim() != target.dim():
            raise ValueError(f"the number of dimensions for input and target should be the same, got shape {input.shape} "
                             f"and {target.shape}.")
        n_ch = input.shape[1]
        if n_ch == 1:
            raise NotImplementedError("single-channel predictions take MONAI's BCE branch, which is not on this path")
        with torch.autocast(input.device.type, enabled=False):
            logits = input.float()
            logp = F.log_softmax(logits, dim=1)
            # class-index target for the cross-entropy (dice.py:741-761); one-hot target for the Dice term
            if target.shape[1] == 1:
                idx = target[:, 0].long()
                onehot = F.one_hot(idx, n_ch).movedim(-1, 1).to(logp.dtype) if self.to_onehot_y else target.to(logp.dtype)
            else:
                onehot = target.to(logp.dtype)
                idx = target.argmax(dim=1)
            ce = F.nll_loss(logp, idx, reduction=self.reduction)
            pred = logp.exp() if self.softmax else logits
            if onehot.shape != pred.shape:
                raise AssertionError(f"ground truth has different shape ({onehot.shape}) from input ({pred.shape})")
            if not self.include_background:
                pred, onehot = pred[:, 1:], onehot[:, 1:]
            axes = list(range(2, pred.dim()))
            if self.batch:
                axes = [0] + axes
            inter = (onehot * pred).sum(dim=axes)
            if self.squared_pred:
                denom = (onehot ** 2).sum(dim=axes) + (pred ** 2).sum(dim=axes)
            else:
                denom = onehot.sum(dim=axes) + pred.sum(dim=axes)
            if self.jaccard:
                denom = 2.0 * (denom - inter)
            f = 1.0 - (2.0 * inter + self.smooth_nr) / (denom + self.smooth_dr)
            dice = f.mean() if self.reduction == "mean" else f.sum()
            return self.lambda_dice * dice + self.lambda_ce * ce
