"""EyeNet: eye patch (+ head pose) -> gaze direction and pupil size.

Mirror of the reference's ``src/models/eye_net.py`` (EyeNet :37-150): same constructor, same
``forward(input_dict, output_dict, side, previous_output_dict=None)`` contract, same config
knobs (read at construct AND forward time) and the same state_dict keys.  The arithmetic runs
in libeve_b200.so: ``eve_eyenet_cnn_*`` (ResNet-18 / InstanceNorm) and ``eve_eyenet_tail_*``
(fc_common, RNN cells or static_fc, gaze and pupil heads).
"""
import math

import torch
from torch import nn

from .. import lib as L
from .. import ops, synth
from ..config import get_config
from . import _params

config = get_config()
half_pi = 0.5 * math.pi

_CNN_ORDER = None


def cnn_weight_names():
    """Pointer-table order of eve_eyenet_cnn_* (include/eve_b200.h)."""
    global _CNN_ORDER
    if _CNN_ORDER is None:
        names = ['cnn_layers.conv1.weight']
        for li in (1, 2, 3, 4):
            for bi in (0, 1):
                p = 'cnn_layers.layer%d.%d.' % (li, bi)
                names += [p + 'conv1.weight', p + 'conv2.weight']
                if li > 1 and bi == 0:
                    names.append(p + 'downsample.0.weight')
        names += ['cnn_layers.fc.weight', 'cnn_layers.fc.bias']
        _CNN_ORDER = names
    return _CNN_ORDER


def tail_weight_names(use_rnn, num_cells):
    names = ['fc_common.0.weight', 'fc_common.0.bias', 'fc_common.2.weight', 'fc_common.2.bias']
    if use_rnn:
        for i in range(num_cells):
            p = 'rnn_cells.%d.' % i
            names += [p + 'weight_ih', p + 'weight_hh', p + 'bias_ih', p + 'bias_hh']
    else:
        names += ['static_fc.0.weight', 'static_fc.0.bias']
    names += ['fc_to_gaze.0.weight', 'fc_to_gaze.0.bias', 'fc_to_gaze.2.weight',
              'fc_to_pupil.0.weight', 'fc_to_pupil.0.bias', 'fc_to_pupil.2.weight',
              'fc_to_pupil.2.bias']
    return names


def _init(name, shape):
    # Same distributions as the modules the reference instantiates (eye_net.py:48-96):
    # torchvision ResNet convs kaiming-normal fan_out, Linear / RNN cells uniform, and the
    # last gaze layer zeroed (:96).
    if name == 'fc_to_gaze.2.weight':
        return torch.zeros(shape)
    if len(shape) == 4:
        return _params.kaiming_normal_fan_out(shape)
    if name.startswith('rnn_cells.'):
        hidden = config.eye_net_rnn_num_features
        return _params.uniform(shape, 1.0 / math.sqrt(hidden))
    if len(shape) == 2:
        return _params.linear_default_weight(shape)
    # biases of nn.Linear: U(-1/sqrt(fan_in), 1/sqrt(fan_in)); fan_in from the paired weight
    fan_in = _BIAS_FAN_IN.get(name)
    return _params.uniform(shape, 1.0 / math.sqrt(fan_in)) if fan_in else torch.zeros(shape)


_BIAS_FAN_IN = {}


class EyeNet(nn.Module):
    def __init__(self):
        super(EyeNet, self).__init__()
        if config.eye_net_use_rnn and config.eye_net_rnn_type not in ('RNN', 'LSTM', 'GRU'):
            raise ValueError('Unknown RNN type for EyeNet: %s' % config.eye_net_rnn_type)
        shapes = synth.eye_net_param_shapes(config)
        for k, s in shapes.items():
            if k.endswith('.weight') and len(s) == 2:
                _BIAS_FAN_IN[k[:-len('weight')] + 'bias'] = s[1]
        _params.attach(self, shapes, _init)
        self.num_features = shapes['cnn_layers.fc.weight'][0]
        self._use_rnn = bool(config.eye_net_use_rnn)
        self._rnn_type = config.eye_net_rnn_type if self._use_rnn else None
        self._rnn_cells = config.eye_net_rnn_num_cells if self._use_rnn else 0
        self._use_head_pose = bool(config.eye_net_use_head_pose_input)

    # -- pointer tables --------------------------------------------------------------
    def _cnn_weights(self):
        return [_params.lookup(self, n) for n in cnn_weight_names()]

    def _tail_weights(self):
        return [_params.lookup(self, n)
                for n in tail_weight_names(self._use_rnn, self._rnn_cells)]

    # -- whole-sequence entry points (used by EVE.forward) -----------------------------
    def cnn_features(self, patches):
        """patches [N,3,128,128] -> [N, num_features] (eye_net.py:106)."""
        return ops.EyeNetCnnFn.apply(patches, self.num_features, *self._cnn_weights())

    def tail_sequence(self, features, head_pose, h0=None, c0=None):
        """features [S,T,nf], head_pose [S,T,2] -> g [S,T,2], pupil [S,T], hT, cT
        (eye_net.py:109-140 for every time step; S = sequences)."""
        cfg = (self._use_head_pose, L.EYE_RNN_TYPES[self._rnn_type], max(self._rnn_cells, 1))
        hp = head_pose if self._use_head_pose else None
        return ops.EyeNetTailFn.apply(features, hp, h0, c0, cfg, *self._tail_weights())

    # -- the reference's per-time-step interface ---------------------------------------
    def forward(self, input_dict, output_dict, side, previous_output_dict=None):
        key = side + '_eye_patch'
        input_image = output_dict[key] if key in output_dict else input_dict[key]
        B = input_image.shape[0]
        feats = self.cnn_features(input_image).reshape(B, 1, self.num_features)
        hp = input_dict[side + '_h'].reshape(B, 1, 2) if config.eye_net_use_head_pose_input else None
        h0 = c0 = None
        if self._use_rnn and previous_output_dict is not None:
            hs, cs = [], []
            for i in range(self._rnn_cells):
                st = previous_output_dict[side + '_eye_rnn_states_%d' % i]
                if isinstance(st, tuple):
                    hs.append(st[0])
                    cs.append(st[1])
                else:
                    hs.append(st)
            h0 = torch.stack(hs, 0)
            c0 = torch.stack(cs, 0) if cs else None
        g, pupil, hT, cT = self.tail_sequence(feats, hp, h0, c0)
        if self._use_rnn:
            for i in range(self._rnn_cells):
                states = (hT[i], cT[i]) if cT is not None else hT[i]
                output_dict[side + '_eye_rnn_states_%d' % i] = states
        output_dict[side + '_g_initial'] = g[:, 0]
        output_dict[side + '_pupil_size'] = pupil[:, 0].reshape(-1)
        if config.eye_net_frozen:
            output_dict[side + '_g_initial'] = output_dict[side + '_g_initial'].detach()
