"""Drop-in for the hot-path members of the reference's ``utils/losses.py`` (mask_DiceLoss :8-77, DiceLoss :79-134,
to_one_hot :173-189, get_probability :192-206) backed by the fused Dice/CE kernel."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


class mask_DiceLoss(nn.Module):
    """Per-(n,c) masked soft Dice on softmax probabilities, smooth 1e-5 (utils/losses.py:47-77)."""

    def __init__(self, nclass, class_weights=None, smooth=1e-5):
        super().__init__()
        if abs(smooth - 1e-5) > 0:
            raise NotImplementedError("the fused kernel hard-codes the reference's smooth=1e-5")
        if class_weights is not None:
            raise NotImplementedError("class_weights are stored but never applied by the reference; not supported")
        self.smooth = smooth
        self.class_weights = nn.Parameter(torch.ones((1, nclass), dtype=torch.float32), requires_grad=False)

    def forward(self, logits, target, mask=None):
        if logits.shape[1] < 2:
            raise NotImplementedError("single-channel (sigmoid) logits are never used by the entry points")
        t = ops.to_u8_labels(target.reshape((logits.shape[0],) + tuple(logits.shape[2:])))
        from .BCP_utils import _box_of
        box = _box_of(mask) if mask is not None else (0, 0, 0, 0, 0, 0)
        m8 = None
        if mask is not None and box is None:
            m8 = (mask.reshape(t.shape) != 0).to(torch.uint8)
        return ops.MixLoss.apply(logits, t, t, box, m8, 0, 1.0, 0.0)[1]


class DiceLoss(nn.Module):
    """Batch-global per-class Dice, smooth 1e-10 (utils/losses.py:79-134).  ``softmax=False`` (the call pattern of
    ACDC_BCP_train.py:170,175: probabilities in, mask [N,1,H,W]) runs the dice-on-probabilities kernel pair;
    ``softmax=True`` differentiates through the softmax inside the fused Dice/CE kernel."""

    def __init__(self, n_classes):
        super().__init__()
        self.n_classes = n_classes

    def forward(self, inputs, target, mask=None, weight=None, softmax=False):
        if weight is not None:
            raise NotImplementedError("per-class weights are never passed by the entry points")
        n = inputs.shape[0]
        assert inputs.shape[1] == self.n_classes, "predict & target shape do not match"
        t = ops.to_u8_labels(target.reshape((n,) + tuple(inputs.shape[2:])))
        m8 = None if mask is None else (mask.reshape(t.shape) != 0).to(torch.uint8)
        if softmax:
            box = None if mask is not None else (0, 0, 0, 0)
            return ops.MixLoss.apply(inputs, t, t, box, m8, 1, 1.0, 0.0)[1]
        return ops.DiceProb.apply(inputs, t, m8)


def to_one_hot(tensor, nClasses):
    """API shim (utils/losses.py:173-189); not used by the fused step."""
    size = list(tensor.size())
    assert size[1] == 1
    size[1] = nClasses
    return torch.zeros(*size, device=tensor.device).scatter_(1, tensor, 1)


def get_probability(logits):
    """API shim (utils/losses.py:192-206); not used by the fused step."""
    if logits.size(1) > 1:
        return F.softmax(logits, dim=1), logits.size(1)
    pred = torch.sigmoid(logits)
    return torch.cat([1 - pred, pred], 1), 2
