"""The two mmdet 2.20 loss classes the SSL detector instantiates (…_ssl.py:106,113-123),
restated so the reference config builds unchanged.  Only CrossEntropyLoss(use_sigmoid=True)
is evaluated on the mse_loss=True path (…_ssl.py:894-895)."""
import torch.nn.functional as F
from torch import nn

from .registry import LOSSES


@LOSSES.register_module()
class SmoothL1Loss(nn.Module):
    def __init__(self, beta=1.0, reduction="mean", loss_weight=1.0):
        super().__init__()
        self.beta, self.reduction, self.loss_weight = beta, reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        assert weight is None and avg_factor is None
        return self.loss_weight * F.smooth_l1_loss(pred, target, beta=self.beta,
                                                   reduction=reduction_override or self.reduction)


@LOSSES.register_module()
class CrossEntropyLoss(nn.Module):
    """use_sigmoid=True: labels one-hot expanded to the logit width, BCE-with-logits, mean."""

    def __init__(self, use_sigmoid=False, use_mask=False, reduction="mean", class_weight=None,
                 ignore_index=None, loss_weight=1.0):
        super().__init__()
        if not use_sigmoid or use_mask or reduction != "mean" or class_weight is not None:
            raise NotImplementedError("only CrossEntropyLoss(use_sigmoid=True, reduction='mean') is on the path")
        self.loss_weight = loss_weight

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None):
        assert weight is None and avg_factor is None
        onehot = F.one_hot(label.long(), cls_score.shape[-1]).to(cls_score.dtype)
        return self.loss_weight * F.binary_cross_entropy_with_logits(cls_score, onehot, reduction="mean")
