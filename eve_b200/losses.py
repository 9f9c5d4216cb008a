"""Validity-masked sequence losses of the EVE hot path.

Same call contract as the reference's ``src/losses`` objects --
``loss(predictions, gt_key, reference_dict) -> scalar`` (base_loss_with_validity.py:32-73) --
but evaluated for all clips at once instead of a Python loop over the batch (and, for the
BCE, over time: cross_entropy.py:29-35).
"""
import math

import torch
import torch.nn.functional as F


def _masked_clip_mean(per_frame, validity):
    """Per clip sum(v * l) / n_valid (only when n_valid > 1), then the mean over clips."""
    v = validity.to(per_frame.dtype)
    assert v.shape == per_frame.shape
    n = v.sum(dim=1)
    acc = (v * per_frame).sum(dim=1)
    acc = torch.where(n > 1, acc / torch.clamp(n, min=1.0), acc)
    return acc.sum() / float(per_frame.shape[0])


class _Loss(object):
    def per_frame(self, a, b):
        raise NotImplementedError

    def __call__(self, predictions, gt_key, reference_dict):
        validity_key = gt_key + '_validity'
        assert validity_key in reference_dict
        return _masked_clip_mean(self.per_frame(predictions, reference_dict[gt_key]),
                                 reference_dict[validity_key])


def _feature_dims(a):
    return tuple(range(2, a.ndim))


class AngularLoss(_Loss):
    """angular.py:29-38, degrees."""

    def per_frame(self, a, b):
        from .models.common import pitchyaw_to_vector
        va, vb = pitchyaw_to_vector(a), pitchyaw_to_vector(b)
        sim = F.cosine_similarity(va, vb, dim=-1, eps=1e-8)
        sim = torch.clamp(sim, -1.0 + 1e-8, 1.0 - 1e-8)
        return torch.acos(sim) * (180.0 / math.pi)


class MSELoss(_Loss):
    def per_frame(self, a, b):
        return ((a - b) ** 2).mean(dim=_feature_dims(a)) if a.ndim > 2 else (a - b) ** 2


class L1Loss(_Loss):
    def per_frame(self, a, b):
        return (a - b).abs().mean(dim=_feature_dims(a)) if a.ndim > 2 else (a - b).abs()


class EuclideanLoss(_Loss):
    def per_frame(self, a, b):
        return torch.sqrt(((a - b) ** 2).sum(dim=_feature_dims(a)))


class CrossEntropyLoss(_Loss):
    """cross_entropy.py:29-35: F.binary_cross_entropy per frame.  The torch op is kept (not
    a log/clamp restatement) because its backward stays finite when the sigmoid saturates
    to exactly 0 or 1, which a clamp-of-log formulation does not (0 * inf)."""

    def per_frame(self, a, b):
        return F.binary_cross_entropy(a, b, reduction='none').mean(dim=_feature_dims(a))


cross_entropy_loss = CrossEntropyLoss()
euclidean_loss = EuclideanLoss()
angular_loss = AngularLoss()
mse_loss = MSELoss()
l1_loss = L1Loss()
