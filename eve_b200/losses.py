"""Validity-masked sequence losses of the EVE hot path.

Same call contract as the reference's ``src/losses`` objects --
``loss(predictions, gt_key, reference_dict) -> scalar`` (base_loss_with_validity.py:32-73).
On the product path every term is evaluated by the fused kernels of csrc/losses.cu
(``eve_masked_losses_*``: one launch for a whole table of terms, ``eve_heatmap_frame_losses_*``
for the per-frame BCE / MSE of heatmaps); ``evaluate_terms`` is what ``EVE.forward`` calls with all
of its ~30 terms at once.  ``torch_formula`` is the same arithmetic in torch (all clips at once
instead of the reference's Python loop over the batch -- and over time for the BCE,
cross_entropy.py:29-35): it is what oracle/check_host_logic.py pins to the unmodified reference
classes and what the GPU tests compare the kernels with; it is not on the product path.
"""
import math

import torch
import torch.nn.functional as F

from . import lib as L
from . import ops


def _masked_clip_mean(per_frame, validity):
    """Per clip sum(v * l) / n_valid (only when n_valid > 1), then the mean over clips."""
    v = validity.to(per_frame.dtype)
    assert v.shape == per_frame.shape
    n = v.sum(dim=1)
    acc = (v * per_frame).sum(dim=1)
    acc = torch.where(n > 1, acc / torch.clamp(n, min=1.0), acc)
    return acc.sum() / float(per_frame.shape[0])


def _feature_dims(a):
    return tuple(range(2, a.ndim))


def evaluate_terms(terms, return_vector=False):
    """``terms``: list of (loss object, predictions [B,T,...], ground truth, validity [B,T],
    optional second validity).  Returns one scalar tensor per term (views of one [n] vector that
    a single kernel launch produced; gradients flow to the predictions through one more launch).
    ``return_vector``: also return that [n] vector (None when the table needed several launches)."""
    if not terms:
        return ([], None) if return_vector else []
    preds, index, spec = [], {}, []

    def slot(t):
        key = id(t)
        if key not in index:
            index[key] = len(preds)
            preds.append(t)
        return index[key]

    frame_cache = {}
    for loss, pred, gt, valid, valid2 in terms:
        L.require_cuda(pred, type(loss).__name__)
        if pred.ndim > 3:
            # heatmap-shaped term: per-frame BCE and MSE from ONE read of the pair, then the
            # masked clip mean of the per-frame values
            assert loss.op in ('bce', 'mse'), loss.op
            key = (id(pred), id(gt))
            if key not in frame_cache:
                frame_cache[key] = ops.HeatmapFrameLossFn.apply(pred, gt.detach())
            bce, mse = frame_cache[key]
            spec.append(('identity', slot(bce if loss.op == 'bce' else mse), None, valid, valid2))
        else:
            spec.append((loss.op, slot(pred), gt.detach(), valid, valid2))
    out, vecs = [], []
    for at in range(0, len(spec), L.LOSS_MAX_TERMS):
        chunk = spec[at:at + L.LOSS_MAX_TERMS]
        used = sorted({pi for _, pi, _, _, _ in chunk})
        remap = {pi: i for i, pi in enumerate(used)}
        chunk = [(op, remap[pi], gt, v, v2) for op, pi, gt, v, v2 in chunk]
        vec = ops.MaskedLossesFn.apply(chunk, *[preds[pi] for pi in used])
        vecs.append(vec)
        out.extend(vec.unbind(0))
    if return_vector:
        return out, (vecs[0] if len(vecs) == 1 else None)
    return out


class _Loss(object):
    op = None

    def per_frame(self, a, b):
        raise NotImplementedError

    def torch_formula(self, predictions, gt_key, reference_dict):
        validity_key = gt_key + '_validity'
        assert validity_key in reference_dict
        return _masked_clip_mean(self.per_frame(predictions, reference_dict[gt_key]),
                                 reference_dict[validity_key])

    def __call__(self, predictions, gt_key, reference_dict):
        validity_key = gt_key + '_validity'
        assert validity_key in reference_dict
        return evaluate_terms([(self, predictions, reference_dict[gt_key],
                                reference_dict[validity_key], None)])[0]


class AngularLoss(_Loss):
    """angular.py:29-38, degrees."""
    op = 'angular'

    def per_frame(self, a, b):
        from .models.common import pitchyaw_to_vector
        va, vb = pitchyaw_to_vector(a), pitchyaw_to_vector(b)
        sim = F.cosine_similarity(va, vb, dim=-1, eps=1e-8)
        sim = torch.clamp(sim, -1.0 + 1e-8, 1.0 - 1e-8)
        return torch.acos(sim) * (180.0 / math.pi)


class MSELoss(_Loss):
    op = 'mse'

    def per_frame(self, a, b):
        return ((a - b) ** 2).mean(dim=_feature_dims(a)) if a.ndim > 2 else (a - b) ** 2


class L1Loss(_Loss):
    op = 'l1'

    def per_frame(self, a, b):
        return (a - b).abs().mean(dim=_feature_dims(a)) if a.ndim > 2 else (a - b).abs()


class EuclideanLoss(_Loss):
    op = 'euclidean'

    def per_frame(self, a, b):
        return torch.sqrt(((a - b) ** 2).sum(dim=_feature_dims(a)))


class CrossEntropyLoss(_Loss):
    """cross_entropy.py:29-35: F.binary_cross_entropy per frame (logs clamped at -100; the
    backward divides by max((1 - a) a, 1e-12), so it stays finite when the sigmoid saturates)."""
    op = 'bce'

    def per_frame(self, a, b):
        return F.binary_cross_entropy(a, b, reduction='none').mean(dim=_feature_dims(a))


cross_entropy_loss = CrossEntropyLoss()
euclidean_loss = EuclideanLoss()
angular_loss = AngularLoss()
mse_loss = MSELoss()
l1_loss = L1Loss()
