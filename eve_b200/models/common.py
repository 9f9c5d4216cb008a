"""Gaze geometry, heatmaps, soft-argmax and gaze-history maps of the EVE hot path.

Same names, argument meaning and results as the reference's ``src/models/common.py``; the
per-frame pieces EVE.forward calls -- ``to_screen_coordinates`` (:149-179),
``calculate_combined_gaze_direction`` (:129-146), ``apply_offset_augmentation`` (:182-218),
``batch_make_heatmaps`` (:226-243), ``soft_argmax`` (:294-323) and the gaze-history maps
(:249-287, as an O(T) recurrence) -- run as hand-written CUDA kernels through the C ABI
(eve_b200/ops.py).  The small angle / rotation helpers stay torch arithmetic; the ``_*_torch``
functions are the formulas the kernels are checked against (tests only).
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .. import lib as L
from .. import ops
from ..config import get_config

config = get_config()


# ------------------------------------------------------------------------- angles --
def pitchyaw_to_vector(a):
    """common.py:32-40."""
    if a.shape[-1] == 2:
        sin, cos = torch.sin(a), torch.cos(a)
        return torch.stack([cos[..., 0] * sin[..., 1], sin[..., 0], cos[..., 0] * cos[..., 1]],
                           dim=-1)
    if a.shape[-1] == 3:
        return F.normalize(a, dim=-1)
    raise ValueError('Do not know how to convert tensor of size %s' % (a.shape,))


def vector_to_pitchyaw(a):
    """common.py:43-54."""
    if a.shape[-1] == 2:
        return a
    if a.shape[-1] == 3 or (a.ndim >= 2 and a.shape[-2:] == (3, 1)):
        if a.shape[-1] != 3:
            a = a.squeeze(-1)
        n = a / (torch.norm(a, dim=-1, keepdim=True) + 1e-7)
        return torch.stack([torch.asin(n[..., 1]), torch.atan2(n[..., 0], n[..., 2])], dim=-1)
    raise ValueError('Do not know how to convert tensor of size %s' % (a.shape,))


def pitchyaw_to_rotation(a):
    """common.py:57-76: R = Ry(yaw) . Rx(pitch)."""
    if a.shape[-1] == 3:
        a = vector_to_pitchyaw(a)
    cos, sin = torch.cos(a), torch.sin(a)
    one, zero = torch.ones_like(cos[..., 0]), torch.zeros_like(cos[..., 0])
    rx = torch.stack([one, zero, zero, zero, cos[..., 0], sin[..., 0], zero, -sin[..., 0],
                      cos[..., 0]], dim=-1).reshape(*a.shape[:-1], 3, 3)
    ry = torch.stack([cos[..., 1], zero, sin[..., 1], zero, one, zero, -sin[..., 1], zero,
                      cos[..., 1]], dim=-1).reshape(*a.shape[:-1], 3, 3)
    return torch.matmul(ry, rx)


def rotation_to_vector(a):
    """common.py:79-86."""
    assert a.shape[-2:] == (3, 3)
    return a[..., :, 2:3]


def _as_vec3(vec):
    if vec.shape[-1] == 2:
        vec = pitchyaw_to_vector(vec)
    elif vec.shape[-2:] == (3, 1):
        vec = vec.squeeze(-1)
    return vec


def apply_transformation(T, vec):
    """common.py:89-94: homogeneous 4x4 transform of 3-vectors."""
    vec = _as_vec3(vec)
    return torch.matmul(T[..., :3, :3], vec.unsqueeze(-1)).squeeze(-1) + T[..., :3, 3]


def apply_rotation(T, vec):
    """common.py:97-102: rotation part of a transform applied to 3-vectors."""
    vec = _as_vec3(vec)
    return torch.matmul(T[..., :3, :3], vec.unsqueeze(-1)).squeeze(-1)


def get_intersect_with_zero(o, g):
    """common.py:109-126: intersect the ray o + t g with the plane z = 0."""
    o, g = _as_vec3(o), _as_vec3(g)
    t = (0.0 - o[..., 2]) / (g[..., 2] + 1e-7)
    return o[..., :2] + t.unsqueeze(-1) * g[..., :2]


def _combined_gaze_torch(avg_origin, avg_PoG, head_rotation, camera_transformation):
    """common.py:129-146 in torch arithmetic: the formula the kernel (and its hand-derived VJP in
    csrc/gaze_math.cuh) is checked against on the CPU; not on the product path."""
    p3 = F.pad(avg_PoG, (0, 1))
    p3 = apply_transformation(camera_transformation, p3)
    d = torch.matmul(head_rotation, (p3 - avg_origin).unsqueeze(-1)).squeeze(-1)
    return vector_to_pitchyaw(-d)


def calculate_combined_gaze_direction(avg_origin, avg_PoG, head_rotation, camera_transformation):
    """common.py:129-146 as one CUDA kernel (one thread per frame; leading dimensions free)."""
    L.require_cuda(avg_PoG, 'calculate_combined_gaze_direction')
    lead = avg_PoG.shape[:-1]
    g = ops.CombinedGazeFn.apply(avg_origin.reshape(-1, 3), avg_PoG.reshape(-1, 2),
                                 head_rotation.reshape(-1, 3, 3),
                                 camera_transformation.reshape(-1, 4, 4))
    return g.reshape(*lead, 2)


def to_screen_coordinates(origin, direction, rotation, reference_dict):
    """common.py:149-179 as one CUDA kernel per call (all samples at once).

    Leading dimensions are free: [n, .] as in the reference or [B, T, .]."""
    lead = direction.shape[:-1]
    inv_cam = reference_dict['inv_camera_transformation']
    ppm = reference_dict['pixels_per_millimeter']
    mm, px = ops.PogFn.apply(origin.reshape(-1, 3), direction.reshape(-1, 2),
                             rotation.reshape(-1, 3, 3), inv_cam.reshape(-1, 4, 4),
                             ppm.reshape(-1, 2), config.actual_screen_size)
    return mm.reshape(*lead, 2), px.reshape(*lead, 2)


def apply_offset_augmentation(gaze_direction, head_rotation, kappa, inverse_kappa=False):
    """common.py:182-218 as one CUDA kernel (one thread per frame).  A kappa that is one draw per
    clip expanded over time (eve.py:466-477) is read through its [B, 2] base, not copied."""
    L.require_cuda(gaze_direction, 'apply_offset_augmentation')
    lead = gaze_direction.shape[:-1]
    fpk = 1
    if kappa.ndim == 3 and kappa.shape[1] > 1 and kappa.stride(1) == 0:
        fpk = kappa.shape[1]
        kappa = kappa[:, 0]
    out = ops.OffsetAugmentationFn.apply(gaze_direction.reshape(-1, 2),
                                         head_rotation.reshape(-1, 3, 3), kappa.reshape(-1, 2),
                                         fpk, bool(inverse_kappa))
    return out.reshape(*lead, 2)


def _offset_augmentation_torch(gaze_direction, head_rotation, kappa, inverse_kappa=False):
    """common.py:182-218 in torch arithmetic: checker of the kernel math (tests/test_host_math.py,
    oracle/check_host_logic.py); not on the product path."""
    d = -pitchyaw_to_vector(gaze_direction)
    d = -torch.matmul(head_rotation.transpose(-1, -2), d.unsqueeze(-1)).squeeze(-1)
    kv = pitchyaw_to_vector(kappa)
    if inverse_kappa:
        kv = torch.cat([-kv[..., :2], kv[..., 2:]], dim=-1)
    rot = pitchyaw_to_rotation(vector_to_pitchyaw(d))
    d = -torch.matmul(rot, kv.unsqueeze(-1)).squeeze(-1)
    d = -torch.matmul(head_rotation, d.unsqueeze(-1)).squeeze(-1)
    return vector_to_pitchyaw(d)


# ----------------------------------------------------------------------- heatmaps --
def batch_make_heatmaps(centres, sigma):
    """common.py:242-243: centres [..., 2] in screen pixels -> [..., 1, H, W]."""
    lead = centres.shape[:-1]
    w, h = config.gaze_heatmap_size
    out = ops.HeatmapFn.apply(centres.reshape(-1, 2), float(sigma), (w, h),
                              config.actual_screen_size)
    return out.reshape(*lead, 1, h, w)


def make_heatmap(centre, sigma):
    """common.py:226-239: one centre [2] -> [1, H, W]."""
    return batch_make_heatmaps(centre.reshape(1, 2), sigma)[0]


def soft_argmax(heatmaps):
    """common.py:294-323: [n, 1, H, W] -> [n, 2] screen pixels."""
    n, _, h, w = heatmaps.shape
    assert w == config.gaze_heatmap_size[0]
    assert h == config.gaze_heatmap_size[1]
    return ops.SoftArgmaxFn.apply(heatmaps, config.actual_screen_size)


def gaze_history_weights(history_timestamps, validity):
    """Per-(b, t', t) weights of common.py:249-273 for every prefix length t at once:
    validity[t'] * decay ** ((last non-zero timestamp up to t) - timestamp[t']) ms, zero for
    padded entries (timestamp == 0) and for t' > t.  Returns [B, T(prefix), T(history)]."""
    ts = history_timestamps
    B, T = ts.shape
    nz = ts != 0
    idx = torch.arange(T, device=ts.device)
    # last non-zero index within each prefix 0..t
    last = torch.cummax(torch.where(nz, idx.view(1, T), torch.full_like(ts, -1)), dim=1)[0]
    target = ts.gather(1, last.clamp(min=0))                       # [B, T]
    diff = (target.unsqueeze(2) - ts.unsqueeze(1)).to(torch.float32) * 1e-6   # [B, t, t']
    decay = torch.tensor(config.gaze_history_map_decay_per_ms, dtype=torch.float32,
                         device=ts.device)
    wgt = torch.pow(decay, diff)
    mask = (idx.view(1, 1, T) <= idx.view(1, T, 1)) & nz.unsqueeze(1) & (last >= 0).unsqueeze(2)
    return wgt * mask.to(wgt.dtype) * validity.to(wgt.dtype).unsqueeze(1)


def batch_make_gaze_history_maps(history_timestamps, heatmaps, validity):
    """common.py:276-287.  ``heatmaps``: list of [B, 1, H, W] (history so far, as in the
    reference) or a stacked [B, t, 1, H, W] tensor; returns the map after the last entry."""
    if isinstance(heatmaps, (list, tuple)):
        heatmaps = torch.stack(list(heatmaps), dim=1)
    t = heatmaps.shape[1]
    return all_gaze_history_maps(history_timestamps[:, :t], heatmaps, validity[:, :t])[:, -1]


def _batch_make_gaze_history_maps_torch(history_timestamps, heatmaps, validity):
    """common.py:276-287 from the explicit weights (torch arithmetic; checker only)."""
    if isinstance(heatmaps, (list, tuple)):
        heatmaps = torch.stack(list(heatmaps), dim=1)
    t = heatmaps.shape[1]
    wgt = gaze_history_weights(history_timestamps[:, :t], validity[:, :t])[:, t - 1]   # [B, t]
    return (wgt.view(wgt.shape[0], t, 1, 1, 1).detach() * heatmaps).sum(dim=1)


def make_gaze_history_map(history_timestamps, heatmaps, validities):
    """common.py:249-273 for one clip: timestamps [t], heatmaps list/[t, 1, H, W]."""
    if isinstance(heatmaps, (list, tuple)):
        heatmaps = torch.stack(list(heatmaps), dim=0)
    return batch_make_gaze_history_maps(history_timestamps.unsqueeze(0), heatmaps.unsqueeze(0),
                                        validities.unsqueeze(0))[0]


def all_gaze_history_maps(history_timestamps, heatmaps, validity):
    """Every prefix at once: heatmaps [B, T, 1, H, W] -> [B, T, 1, H, W] where slice t is what
    the reference computes after step t (the O(T^2) Python loop of eve.py:596-601).  On the GPU
    this is the O(T) recurrence kernel eve_gaze_history_fwd (one pass over the heatmaps)."""
    L.require_cuda(heatmaps, 'gaze history maps')
    return ops.GazeHistoryFn.apply(heatmaps, history_timestamps, validity,
                                   float(config.gaze_history_map_decay_per_ms))


def _all_gaze_history_maps_torch(history_timestamps, heatmaps, validity):
    """The same maps from the explicit [T x T] weights (torch arithmetic): checker of the
    recurrence kernel against the reference's loops; not on the product path."""
    B, T = heatmaps.shape[:2]
    wgt = gaze_history_weights(history_timestamps, validity)            # [B, T, T]
    flat = heatmaps.reshape(B, T, -1)
    return torch.bmm(wgt.detach().to(flat.dtype), flat).reshape(heatmaps.shape)


# ------------------------------------------------------------------- ConvRNN cells --
# Stand-alone modules with the reference's names, constructor arguments, parameter names and
# per-step forward contract (common.py:326-415).  RefineNet itself runs these recurrences inside
# eve_refinenet_{fwd,bwd} over whole sequences; the classes exist so that code importing them
# from ``models.common`` keeps working, and they execute the same library convolution.
class Flatten(nn.Module):
    def forward(self, x):
        return x.view(x.size()[0], -1)


class _ConvParam(nn.Module):
    """Holds ``weight`` / ``bias`` under the name the reference's nn.Conv2d child had."""

    def __init__(self, cin, cout, k=3):
        super(_ConvParam, self).__init__()
        fan_in = cin * k * k
        bound = 1.0 / math.sqrt(fan_in)
        self.weight = nn.Parameter((torch.rand(cout, cin, k, k) * 2 - 1) * bound)
        self.bias = nn.Parameter((torch.rand(cout) * 2 - 1) * bound)
        self.padding = k // 2

    def forward(self, x):
        return ops.conv2d(x, self.weight, self.bias, 1, self.padding)


class CRNNCell(nn.Module):
    """common.py:331-352: h' = tanh(conv3x3(cat[x, h]))."""

    def __init__(self, input_size, hidden_size):
        super(CRNNCell, self).__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.cell = _ConvParam(input_size + hidden_size, hidden_size)

    def forward(self, x, previous_states=None):
        h = previous_states
        if h is None:
            h = x.new_zeros([x.shape[0], self.hidden_size] + list(x.shape[2:]))
        return torch.tanh(self.cell(torch.cat([x, h], dim=1)))


class CLSTMCell(nn.Module):
    """common.py:355-385: gates chunked as (in, forget, out, cell)."""

    def __init__(self, input_size, hidden_size):
        super(CLSTMCell, self).__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.gates = _ConvParam(input_size + hidden_size, 4 * hidden_size)

    def forward(self, x, previous_states=None):
        if previous_states is None:
            shape = [x.shape[0], self.hidden_size] + list(x.shape[2:])
            h, c = x.new_zeros(shape), x.new_zeros(shape)
        else:
            h, c = previous_states
        gi, gf, go, gc = self.gates(torch.cat([x, h], dim=1)).chunk(4, 1)
        cell = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gc)
        return torch.sigmoid(go) * torch.tanh(cell), cell


class CGRUCell(nn.Module):
    """common.py:388-415 (note the [r*h, x] order of the second concatenation, :412)."""

    def __init__(self, input_size, hidden_size):
        super(CGRUCell, self).__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.gates_1 = _ConvParam(input_size + hidden_size, 2 * hidden_size)
        self.gate_2 = _ConvParam(input_size + hidden_size, hidden_size)

    def forward(self, x, previous_states=None):
        h = previous_states
        if h is None:
            h = x.new_zeros([x.shape[0], self.hidden_size] + list(x.shape[2:]))
        r, z = torch.sigmoid(self.gates_1(torch.cat([x, h], dim=1))).chunk(2, 1)
        n = torch.tanh(self.gate_2(torch.cat([r * h, x], dim=1)))
        return (1.0 - z) * n + z * h
