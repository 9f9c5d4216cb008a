"""GazeRefineNet: screen frame + initial gaze heatmap -> refined heatmap.

Mirror of the reference's ``src/models/refine_net.py`` (RefineNet :179-255 with BasicBlock,
WrapEncoderDecoder, Bottleneck): same constructor / ``forward(input_dict, output_dict,
previous_output_dict=None)`` contract, config knobs and state_dict keys; the arithmetic runs in
libeve_b200.so (``eve_refinenet_*``).
"""
import torch
import torch.nn.functional as F
from torch import nn

from .. import lib as L
from .. import ops, synth
from ..config import get_config
from . import _params

config = get_config()


def _init(name, shape):
    # refine_net.py:227-235: convs kaiming-normal fan_out with zero bias, norms (1, 0),
    # last 1x1 conv zeroed.
    if name == 'final.2.weight':
        return torch.zeros(shape)
    if len(shape) == 4:
        return _params.kaiming_normal_fan_out(shape)
    if name.endswith('.weight'):      # InstanceNorm gain
        return torch.ones(shape)
    return torch.zeros(shape)


class RefineNet(nn.Module):
    def __init__(self):
        super(RefineNet, self).__init__()
        if config.refine_net_use_rnn and \
                config.refine_net_rnn_type not in ('CRNN', 'CLSTM', 'CGRU'):
            raise ValueError('Unknown RNN type for RefineNet: %s' % config.refine_net_rnn_type)
        if list(config.screen_size) != [128, 72]:
            raise ValueError('RefineNet is built for screen_size [128, 72] (refine_net.py:188-212)')
        _params.attach(self, synth.refine_net_param_shapes(config), _init)
        self._in_c = 4 if config.load_screen_content else 1
        self._use_skip = bool(config.refine_net_use_skip_connections)
        self._rnn_type = config.refine_net_rnn_type if config.refine_net_use_rnn else None
        self._rnn_cells = config.refine_net_rnn_num_cells if config.refine_net_use_rnn else 0
        self._nf = config.refine_net_num_features
        self._cfg = (self._in_c, self._use_skip, L.REFINE_RNN_TYPES[self._rnn_type],
                     max(self._rnn_cells, 1), self._nf)
        self._names = None

    def _weights(self):
        if self._names is None:
            p = L.RefineNetParams(1, 1, *[int(v) for v in self._cfg])
            self._names = ops.refinenet_weight_names(p)
        return [_params.lookup(self, n) for n in self._names]

    def sequence(self, screen, heatmap, h0=None, c0=None):
        """screen [B,T,3,72,128] (or None), heatmap [B,T,1,72,128] -> heatmap_final
        [B,T,1,72,128], hT, cT -- refine_net.py:237-255 for every time step."""
        scr = screen if self._in_c == 4 else None
        return ops.RefineNetFn.apply(scr, heatmap, h0, c0, self._cfg, *self._weights())

    def forward(self, input_dict, output_dict, previous_output_dict=None):
        scr_w, scr_h = config.screen_size
        hm = output_dict['heatmap_initial']
        if tuple(hm.shape[-2:]) != (scr_h, scr_w):
            hm = F.interpolate(hm, (scr_h, scr_w), mode='bilinear', align_corners=False)
        screen = input_dict['screen_frame'].unsqueeze(1) if config.load_screen_content else None
        h0 = c0 = None
        if self._rnn_cells and previous_output_dict is not None:
            hs, cs = [], []
            for i in range(self._rnn_cells):
                st = previous_output_dict['refinenet_rnn_states_%d' % i]
                if isinstance(st, tuple):
                    hs.append(st[0])
                    cs.append(st[1])
                else:
                    hs.append(st)
            h0 = torch.stack(hs, 0)
            c0 = torch.stack(cs, 0) if cs else None
        out, hT, cT = self.sequence(screen, hm.unsqueeze(1), h0, c0)
        for i in range(self._rnn_cells):
            states = (hT[i], cT[i]) if cT is not None else hT[i]
            output_dict['refinenet_rnn_states_%d' % i] = states
        output_dict['heatmap_final'] = out[:, 0]
