"""CPU oracle for the EVE hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch (fp32 or fp64, CPU) functional restatement of the arithmetic the
reference performs in ``src/models/{eye_net,refine_net,common,eve}.py`` and
``src/losses/*.py``.  It exists only to check the CUDA path: nothing under
``eve_b200/`` imports it; only ``tests/``, ``__graft_entry__.smoke()`` and the baseline
legs of ``bench.py`` do (``cpu_baseline`` / ``--impl reference`` on the host cores, and
``also.stock_torch_b200``: the same plain-PyTorch arithmetic timed on the GPU through
cuDNN/cuBLAS as the "existing library path" baseline -- measured beside the product, never
part of it).

Parity status: **pinned against the reference itself.**  The reference ships no tests
or golden vectors, so ``oracle/gen_golden.py`` imports the
unmodified reference modules in the build container, runs them on seeded inputs /
weights (eve_b200/synth.py) and stores their outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` holds this file to those vectors.

Third-party arithmetic restated here because it is not in the reference tree:
torchvision 0.6.1 ``ResNet(BasicBlock,[2,2,2,2])`` (call site eye_net.py:48-50) and
torch 1.5.1 ``InstanceNorm2d / GRUCell / LSTMCell / RNNCell / AdaptiveMaxPool2d /
Upsample(bilinear)`` semantics.

Unlike the reference this restatement is written over whole ``B x T`` blocks wherever
the reference loops over ``b`` or ``t`` with no data dependence, and is
driven by a ``state_dict`` with the reference's parameter names instead of nn.Modules.
"""
import math

import torch
import torch.nn.functional as F

HALF_PI = 0.5 * math.pi
SELU_ALPHA = 1.6732632423543772
SELU_SCALE = 1.0507009873554805


# ----------------------------------------------------------------------------- norms --
def instance_norm(x, gamma=None, beta=None, eps=1e-5):
    """Per-(n, c) normalisation over H x W with the biased variance (torch
    InstanceNorm2d, track_running_stats=False; eye_net.py:50, refine_net.py:46)."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(2, 3), keepdim=True)
    y = (x - mean) * torch.rsqrt(var + eps)
    if gamma is not None:
        y = y * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
    return y


def selu(x):
    return SELU_SCALE * torch.where(x > 0, x, SELU_ALPHA * torch.expm1(x))


# ---------------------------------------------------------------------------- EyeNet --
def resnet18_in_features(sd, pre, x):
    """torchvision ResNet-18 with InstanceNorm2d(affine=False) up to and including
    ``fc`` (torchvision resnet.py: stem, 4 stages x 2 BasicBlocks, avgpool, fc)."""
    x = F.conv2d(x, sd[pre + 'conv1.weight'], stride=2, padding=3)
    x = F.relu(instance_norm(x))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li in (1, 2, 3, 4):
        for bi in (0, 1):
            p = '%slayer%d.%d.' % (pre, li, bi)
            stride = 2 if (li > 1 and bi == 0) else 1
            y = F.conv2d(x, sd[p + 'conv1.weight'], stride=stride, padding=1)
            y = F.relu(instance_norm(y))
            y = instance_norm(F.conv2d(y, sd[p + 'conv2.weight'], padding=1))
            if (p + 'downsample.0.weight') in sd:
                x = instance_norm(F.conv2d(x, sd[p + 'downsample.0.weight'], stride=stride))
            x = F.relu(y + x)
    x = x.mean(dim=(2, 3))
    return F.linear(x, sd[pre + 'fc.weight'], sd[pre + 'fc.bias'])


def _rnn_cell(kind, sd, p, x, state):
    """torch.nn.{RNN,LSTM,GRU}Cell arithmetic (eye_net.py:58-71)."""
    w_ih, w_hh = sd[p + 'weight_ih'], sd[p + 'weight_hh']
    b_ih, b_hh = sd[p + 'bias_ih'], sd[p + 'bias_hh']
    nh = w_hh.shape[1]
    if kind == 'LSTM':
        h, c = state if state is not None else (x.new_zeros(x.shape[0], nh),) * 2
    else:
        h = state if state is not None else x.new_zeros(x.shape[0], nh)
    gi = F.linear(x, w_ih, b_ih)
    gh = F.linear(h, w_hh, b_hh)
    if kind == 'RNN':
        return torch.tanh(gi + gh)
    if kind == 'GRU':
        i_r, i_z, i_n = gi.chunk(3, 1)
        h_r, h_z, h_n = gh.chunk(3, 1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        return (1.0 - z) * n + z * h
    if kind == 'LSTM':
        i, f, g, o = (gi + gh).chunk(4, 1)
        c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        return torch.sigmoid(o) * torch.tanh(c2), c2
    raise ValueError('Unknown RNN type for EyeNet: %s' % kind)


def eye_net_step(sd, cfg, patch, head_pose, prev_states=None, pre='eye_net.'):
    """One ``EyeNet.forward`` call (eye_net.py:98-150) for one eye at one time step.

    Returns (g_initial [B,2], pupil_size [B], list of rnn states)."""
    f = resnet18_in_features(sd, pre + 'cnn_layers.', patch)
    return eye_net_tail_step(sd, cfg, f, head_pose, prev_states, pre)


def eye_net_tail_step(sd, cfg, f, head_pose, prev_states=None, pre='eye_net.'):
    if cfg.eye_net_use_head_pose_input:
        f = torch.cat([f, head_pose], dim=1)
    f = F.linear(f, sd[pre + 'fc_common.0.weight'], sd[pre + 'fc_common.0.bias'])
    f = F.linear(selu(f), sd[pre + 'fc_common.2.weight'], sd[pre + 'fc_common.2.bias'])
    states = []
    if cfg.eye_net_use_rnn:
        for i in range(cfg.eye_net_rnn_num_cells):
            prev = None if prev_states is None else prev_states[i]
            st = _rnn_cell(cfg.eye_net_rnn_type, sd, '%srnn_cells.%d.' % (pre, i), f, prev)
            states.append(st)
            f = st[0] if isinstance(st, tuple) else st
    else:
        f = selu(F.linear(f, sd[pre + 'static_fc.0.weight'], sd[pre + 'static_fc.0.bias']))
    g = F.linear(selu(F.linear(f, sd[pre + 'fc_to_gaze.0.weight'], sd[pre + 'fc_to_gaze.0.bias'])),
                 sd[pre + 'fc_to_gaze.2.weight'])
    g = HALF_PI * torch.tanh(g)
    p = F.linear(selu(F.linear(f, sd[pre + 'fc_to_pupil.0.weight'], sd[pre + 'fc_to_pupil.0.bias'])),
                 sd[pre + 'fc_to_pupil.2.weight'], sd[pre + 'fc_to_pupil.2.bias'])
    return g, F.relu(p).reshape(-1), states


def eye_net_sequence(sd, cfg, patches, head_pose, pre='eye_net.'):
    """EyeNet over a whole clip: patches [B,T,3,128,128], head_pose [B,T,2].

    The CNN sees all B*T patches at once (exact, norms are per sample); only the
    recurrent cell walks over t (eve.py:91-111)."""
    B, T = patches.shape[:2]
    f = resnet18_in_features(sd, pre + 'cnn_layers.', patches.reshape(B * T, *patches.shape[2:]))
    f = f.reshape(B, T, -1)
    gs, ps, states = [], [], None
    for t in range(T):
        g, p, states = eye_net_tail_step(sd, cfg, f[:, t], head_pose[:, t], states or None, pre)
        gs.append(g)
        ps.append(p)
    return torch.stack(gs, 1), torch.stack(ps, 1)


# ------------------------------------------------------------------------- RefineNet --
def _leaky(x):
    return F.leaky_relu(x, 0.01)


def refine_block(sd, p, x, act):
    """Pre-activation residual block (refine_net.py:35-67)."""
    y = act(instance_norm(x, sd[p + 'layers.0.weight'], sd[p + 'layers.0.bias']))
    y = F.conv2d(y, sd[p + 'layers.2.weight'], sd[p + 'layers.2.bias'], padding=1)
    y = act(instance_norm(y, sd[p + 'layers.3.weight'], sd[p + 'layers.3.bias']))
    y = F.conv2d(y, sd[p + 'layers.5.weight'], sd[p + 'layers.5.bias'], padding=1)
    if (p + 'skip_layer.2.weight') in sd:
        s = act(instance_norm(x, sd[p + 'skip_layer.0.weight'], sd[p + 'skip_layer.0.bias']))
        x = F.conv2d(s, sd[p + 'skip_layer.2.weight'], sd[p + 'skip_layer.2.bias'])
    return y + x


def conv_rnn_cell(kind, sd, p, x, state):
    """CRNN / CLSTM / CGRU cells (common.py:331-415), zero initial state."""
    if kind == 'CLSTM':
        h, c = state if state is not None else (torch.zeros_like(x), torch.zeros_like(x))
        gates = F.conv2d(torch.cat([x, h], 1), sd[p + 'gates.weight'], sd[p + 'gates.bias'],
                         padding=1)
        gi, gf, go, gc = gates.chunk(4, 1)            # in, forget, out, cell (common.py:376)
        c2 = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gc)
        return torch.sigmoid(go) * torch.tanh(c2), c2
    h = state if state is not None else torch.zeros_like(x)
    if kind == 'CRNN':
        return torch.tanh(F.conv2d(torch.cat([x, h], 1), sd[p + 'cell.weight'],
                                   sd[p + 'cell.bias'], padding=1))
    if kind == 'CGRU':
        g1 = torch.sigmoid(F.conv2d(torch.cat([x, h], 1), sd[p + 'gates_1.weight'],
                                    sd[p + 'gates_1.bias'], padding=1))
        r, z = g1.chunk(2, 1)
        n = torch.tanh(F.conv2d(torch.cat([r * h, x], 1), sd[p + 'gate_2.weight'],
                                sd[p + 'gate_2.bias'], padding=1))   # [r*h, x] order (:412)
        return (1.0 - z) * n + z * h
    raise ValueError('Unknown RNN type for RefineNet: %s' % kind)


_LEVEL_HW = ((72, 128), (36, 64), (18, 32), (9, 16), (5, 8))


def refine_encoder(sd, cfg, screen, heatmap, pre='refine_net.'):
    """initial + the five encoder stages for a batch of frames; returns the bottleneck
    input [N,64,5,8] and the per-level skip tensors (refine_net.py:115-121,246-251)."""
    if cfg.load_screen_content:
        x = torch.cat([screen, heatmap], 1)
    else:
        x = heatmap
    x = F.conv2d(x, sd[pre + 'initial.0.weight'], sd[pre + 'initial.0.bias'], padding=1)
    x = F.relu(instance_norm(x, sd[pre + 'initial.1.weight'], sd[pre + 'initial.1.bias']))
    x = F.conv2d(x, sd[pre + 'initial.3.weight'], sd[pre + 'initial.3.bias'], padding=1)
    skips = []
    p = pre + 'network.'
    for lvl in range(5):
        j = 0
        while (p + 'encoder_blocks.%d.layers.2.weight' % j) in sd:
            x = refine_block(sd, p + 'encoder_blocks.%d.' % j, x, F.relu)
            j += 1
        skips.append(x)
        if lvl < 4:
            x = F.adaptive_max_pool2d(x, _LEVEL_HW[lvl + 1])
        p += 'between_module.'
    return x, skips


def refine_bottleneck_step(sd, cfg, x, prev_states, pre='refine_net.'):
    """Bottleneck.forward (refine_net.py:153-176) incl. the CLSTM quirk: with tuple
    states the features handed on are NOT replaced by the cell output (:168-174)."""
    p = pre + 'network.' + 'between_module.' * 5
    states = []
    if cfg.refine_net_use_rnn:
        for i in range(cfg.refine_net_rnn_num_cells):
            prev = None if prev_states is None else prev_states[i]
            st = conv_rnn_cell(cfg.refine_net_rnn_type, sd, '%srnn_cells.%d.' % (p, i), x, prev)
            states.append(st)
            if not isinstance(st, tuple):
                x = st
    return x, states


def refine_decoder(sd, cfg, x, skips, pre='refine_net.'):
    """Decoder stages + final head (refine_net.py:122-129,220-224)."""
    for lvl in range(4, -1, -1):
        p = pre + 'network.' + 'between_module.' * lvl
        if lvl < 4:
            x = F.interpolate(x, size=_LEVEL_HW[lvl], mode='bilinear', align_corners=False)
        if cfg.refine_net_use_skip_connections:
            x = torch.cat([x, skips[lvl]], 1)
        x = refine_block(sd, p + 'decoder_blocks.0.', x, _leaky)
    x = _leaky(F.conv2d(x, sd[pre + 'final.0.weight'], sd[pre + 'final.0.bias'], padding=1))
    return torch.sigmoid(F.conv2d(x, sd[pre + 'final.2.weight'], sd[pre + 'final.2.bias']))


def refine_net_step(sd, cfg, screen, heatmap, prev_states=None, pre='refine_net.'):
    """One ``RefineNet.forward`` (refine_net.py:237-255). heatmap [B,1,72,128]."""
    hm = heatmap
    if tuple(hm.shape[2:]) != (cfg.screen_size[1], cfg.screen_size[0]):
        hm = F.interpolate(hm, (cfg.screen_size[1], cfg.screen_size[0]), mode='bilinear',
                           align_corners=False)
    x, skips = refine_encoder(sd, cfg, screen, hm, pre)
    x, states = refine_bottleneck_step(sd, cfg, x, prev_states, pre)
    return refine_decoder(sd, cfg, x, skips, pre), states


def refine_net_sequence(sd, cfg, screen, heatmap, pre='refine_net.'):
    """RefineNet over a clip: screen [B,T,3,72,128], heatmap [B,T,1,72,128]."""
    B, T = heatmap.shape[:2]
    scr = None if screen is None else screen.reshape(B * T, *screen.shape[2:])
    x, skips = refine_encoder(sd, cfg, scr, heatmap.reshape(B * T, *heatmap.shape[2:]), pre)
    x = x.reshape(B, T, *x.shape[1:])
    outs, states = [], None
    for t in range(T):
        y, states = refine_bottleneck_step(sd, cfg, x[:, t], states or None, pre)
        outs.append(y)
    y = torch.stack(outs, 1).reshape(B * T, *x.shape[2:])
    hm = refine_decoder(sd, cfg, y, skips, pre)
    return hm.reshape(B, T, *hm.shape[1:])


# -------------------------------------------------------------------------- geometry --
def pitchyaw_to_vector(a):
    """common.py:32-40 (2-column input)."""
    s, c = torch.sin(a), torch.cos(a)
    return torch.stack([c[..., 0] * s[..., 1], s[..., 0], c[..., 0] * c[..., 1]], dim=-1)


def vector_to_pitchyaw(v):
    """common.py:43-54."""
    n = v / (torch.norm(v, dim=-1, keepdim=True) + 1e-7)
    return torch.stack([torch.asin(n[..., 1]), torch.atan2(n[..., 0], n[..., 2])], dim=-1)


def pitchyaw_to_rotation(a):
    """common.py:57-76: R = Ry(yaw) . Rx(pitch)."""
    c, s = torch.cos(a), torch.sin(a)
    one, zero = torch.ones_like(c[..., 0]), torch.zeros_like(c[..., 0])
    rx = torch.stack([one, zero, zero, zero, c[..., 0], s[..., 0], zero, -s[..., 0], c[..., 0]],
                     -1).reshape(*a.shape[:-1], 3, 3)
    ry = torch.stack([c[..., 1], zero, s[..., 1], zero, one, zero, -s[..., 1], zero, c[..., 1]],
                     -1).reshape(*a.shape[:-1], 3, 3)
    return ry @ rx


def _mv(M, v):
    return (M @ v.unsqueeze(-1)).squeeze(-1)


def to_screen_coordinates(origin, g, R, inv_cam, ppm, screen_wh=(1920.0, 1080.0)):
    """common.py:149-179 on [...,] batches.  Returns PoG_mm, PoG_px."""
    d = -pitchyaw_to_vector(g)
    d = _mv(R.transpose(-1, -2), d)
    d = _mv(inv_cam[..., :3, :3], d)
    o = _mv(inv_cam[..., :3, :3], origin) + inv_cam[..., :3, 3]
    # ray / plane z=0 with the reference's a=(1,0,0), n=(0,0,1) (common.py:109-126)
    t = (0.0 - o[..., 2]) / (d[..., 2] + 1e-7)
    mm = o[..., :2] + t.unsqueeze(-1) * d[..., :2]
    px = torch.stack([torch.clamp(mm[..., 0] * ppm[..., 0], 0.0, screen_wh[0]),
                      torch.clamp(mm[..., 1] * ppm[..., 1], 0.0, screen_wh[1])], -1)
    return mm, px


def combined_gaze_direction(origin, pog_mm, R, cam):
    """common.py:129-146."""
    p3 = F.pad(pog_mm, (0, 1))
    p3 = _mv(cam[..., :3, :3], p3) + cam[..., :3, 3]
    d = -_mv(R, p3 - origin)
    return vector_to_pitchyaw(d)


def offset_augmentation(g, head_R, kappa):
    """common.py:182-218 with inverse_kappa=False."""
    d = -pitchyaw_to_vector(g)
    d = -_mv(head_R.transpose(-1, -2), d)
    kv = pitchyaw_to_vector(kappa)
    d = -_mv(pitchyaw_to_rotation(vector_to_pitchyaw(d)), kv)
    d = -_mv(head_R, d)
    return vector_to_pitchyaw(d)


def make_heatmaps(centres_px, sigma, size_wh=(128, 72), screen_wh=(1920.0, 1080.0)):
    """common.py:226-243 for [..., 2] pixel centres -> [..., 1, H, W]."""
    w, h = size_wh
    xs = torch.arange(w, dtype=centres_px.dtype, device=centres_px.device).view(1, w)
    ys = torch.arange(h, dtype=centres_px.dtype, device=centres_px.device).view(h, 1)
    cx = ((w / screen_wh[0]) * centres_px[..., 0])[..., None, None]
    cy = ((h / screen_wh[1]) * centres_px[..., 1])[..., None, None]
    alpha = -0.5 / (sigma ** 2)
    hm = torch.exp(alpha * ((xs - cx) ** 2 + (ys - cy) ** 2))
    return (1e-8 + hm).unsqueeze(-3)


def soft_argmax(heatmaps, size_wh=(128, 72), screen_wh=(1920.0, 1080.0)):
    """common.py:294-323: [N,1,H,W] -> [N,2] pixels."""
    w, h = size_wh
    xs = torch.linspace(0, 1.0, w, dtype=torch.float64).to(heatmaps.dtype).to(heatmaps.device)
    ys = torch.linspace(0, 1.0, h, dtype=torch.float64).to(heatmaps.dtype).to(heatmaps.device)
    p = F.softmax(1e2 * heatmaps.reshape(-1, h * w), dim=-1).reshape(-1, h, w)
    lx = (p * xs.view(1, 1, w)).sum(dim=(1, 2))
    ly = (p * ys.view(1, h, 1)).sum(dim=(1, 2))
    return torch.stack([torch.clamp(screen_wh[0] * lx, 0.0, screen_wh[0]),
                        torch.clamp(screen_wh[1] * ly, 0.0, screen_wh[1])], -1)


def gaze_history_maps(timestamps, heatmaps, validity, decay):
    """common.py:249-287 for every prefix length at once.

    timestamps [B,T] int64, heatmaps [B,T,1,H,W], validity [B,T] ->
    [B,T,1,H,W] where slice t is what the reference computes after step t."""
    B, T = timestamps.shape
    out = []
    for t in range(T):
        ts = timestamps[:, :t + 1]
        nz = ts != 0
        # last non-zero timestamp of the prefix
        idx = (nz.long() * torch.arange(1, t + 2, device=ts.device).view(1, -1)).argmax(dim=1)
        target = ts.gather(1, idx.view(-1, 1))
        diff = ((target - ts) * 1e-6).to(heatmaps.dtype)
        wgt = torch.pow(torch.tensor(decay, dtype=heatmaps.dtype, device=heatmaps.device), diff)
        wgt = wgt * nz.to(heatmaps.dtype) * validity[:, :t + 1].to(heatmaps.dtype)
        out.append((wgt.view(B, t + 1, 1, 1, 1) * heatmaps[:, :t + 1]).sum(1))
    return torch.stack(out, 1)


# ---------------------------------------------------------------------------- losses --
def masked_clip_mean(per_frame, validity):
    """base_loss_with_validity.py:32-73: per clip sum(v*l)/n_valid (only if n_valid > 1),
    then the mean over clips."""
    v = validity.to(per_frame.dtype)
    n = v.sum(dim=1)
    acc = (v * per_frame).sum(dim=1)
    acc = torch.where(n > 1, acc / torch.clamp(n, min=1.0), acc)
    return acc.sum() / float(per_frame.shape[0])


def angular_error(a, b):
    """angular.py:33-38, degrees."""
    va, vb = pitchyaw_to_vector(a), pitchyaw_to_vector(b)
    sim = F.cosine_similarity(va, vb, dim=-1, eps=1e-8)
    sim = torch.clamp(sim, -1 + 1e-8, 1 - 1e-8)
    return torch.acos(sim) * (180.0 / math.pi)


def _feature_dims(a):
    return tuple(range(2, a.ndim))


def mse_per_frame(a, b):
    return ((a - b) ** 2).mean(dim=_feature_dims(a)) if a.ndim > 2 else (a - b) ** 2


def l1_per_frame(a, b):
    return (a - b).abs().mean(dim=_feature_dims(a)) if a.ndim > 2 else (a - b).abs()


def euclidean_per_frame(a, b):
    return torch.sqrt(((a - b) ** 2).sum(dim=_feature_dims(a)))


def bce_per_frame(a, b):
    """cross_entropy.py:29-35: F.binary_cross_entropy per frame (the reference's own call;
    its backward stays finite when a saturates to exactly 0 or 1).

    fp64 evaluations only (the yardstick the parity tests measure fp32 noise against): the
    reference's heatmaps are 1e-8 + exp(.) (common.py:243), which rounds to <= 1 in fp32 but is
    1 + 1e-8 in fp64 at a pixel centre, where F.binary_cross_entropy raises; clamp there."""
    if a.dtype == torch.float64:
        a = a.clamp(max=1.0)
        b = b.clamp(max=1.0)
    return F.binary_cross_entropy(a, b, reduction='none').mean(dim=_feature_dims(a))


# -------------------------------------------------------------------- input pipeline --
def preprocess_frames(frames):
    """datasources/eve_sequences.py:196-203: N x H x W x C uint8 -> N x C x H x W float32 in [-1, 1]
    (numpy, in-place float32 ops exactly as the reference writes them)."""
    import numpy as np
    frames = np.transpose(frames, [0, 3, 1, 2])
    frames = frames.astype(np.float32)
    frames *= 2.0 / 255.0
    frames -= 1.0
    return frames


def preprocess_screen_frames(frames):
    """datasources/eve_sequences.py:205-211."""
    import numpy as np
    frames = np.transpose(frames, [0, 3, 1, 2])
    frames = frames.astype(np.float32)
    frames *= 1.0 / 255.0
    return frames


def split_eye_patches(frames, ew):
    """eve_sequences.py:283-285 on preprocessed "eyes" frames: (left, right)."""
    return frames[:, :, :, ew:], frames[:, :, :, :ew]


# ------------------------------------------------------------------------------- EVE --
def derive_labels(inp, cfg, training, kappas=None):
    """EVE.calculate_additional_labels (eve.py:441-543), vectorised over B and T.

    ``kappas``: dict side -> [B,2] (the reference draws them from np.random, :468-469)."""
    d = dict(inp)
    for side in ('left', 'right'):
        if side + '_PoG_tobii' in d:
            d[side + '_PoG_cm_tobii'] = d[side + '_PoG_tobii'] * (0.1 * d['millimeters_per_pixel'])
            d[side + '_PoG_cm_tobii_validity'] = d[side + '_PoG_tobii_validity']
    if training and cfg.refine_net_do_offset_augmentation:
        T = d['left_eye_patch'].shape[1]
        for side in ('left', 'right'):
            d[side + '_kappa_fake'] = kappas[side].unsqueeze(1).expand(-1, T, -1)
    if 'left_o' in d:
        d['o'] = torch.stack([d['left_o'], d['right_o']], -1).mean(-1)
        d['o_validity'] = d['left_o_validity']
    if 'left_PoG_tobii' in d:
        d['PoG_px_tobii'] = torch.stack([d['left_PoG_tobii'], d['right_PoG_tobii']], -1).mean(-1)
        d['PoG_cm_tobii'] = torch.stack([d['left_PoG_cm_tobii'], d['right_PoG_cm_tobii']],
                                        -1).mean(-1)
        v = d['left_PoG_tobii_validity'].bool() & d['right_PoG_tobii_validity'].bool()
        d['PoG_px_tobii_validity'] = v
        d['PoG_cm_tobii_validity'] = v
        if cfg.refine_net_enabled:
            vf = v.to(d['PoG_px_tobii'].dtype)[..., None, None, None]
            for name, sigma in (('initial', cfg.gaze_heatmap_sigma_initial),
                                ('history', cfg.gaze_heatmap_sigma_history),
                                ('final', cfg.gaze_heatmap_sigma_final)):
                d['heatmap_' + name] = make_heatmaps(d['PoG_px_tobii'], sigma) * vf
                d['heatmap_' + name + '_validity'] = v
    if 'PoG_cm_tobii' in d:
        d['g'] = combined_gaze_direction(d['o'], 10.0 * d['PoG_cm_tobii'], d['left_R'],
                                         d['camera_transformation'])
        d['g_validity'] = d['PoG_cm_tobii_validity']
    return d


def _pog_bundle(d, out, in_suffix, out_suffix, cfg, sigma):
    """EVE.from_g_to_PoG_history (eve.py:545-601) without the history lists."""
    for side in ('left', 'right'):
        mm, px = to_screen_coordinates(d[side + '_o'], out[side + '_g_' + in_suffix],
                                       d[side + '_R'], d['inv_camera_transformation'],
                                       d['pixels_per_millimeter'],
                                       tuple(float(v) for v in cfg.actual_screen_size))
        out[side + '_PoG_cm_' + out_suffix] = 0.1 * mm
        out[side + '_PoG_px_' + out_suffix] = px
    for unit in ('px', 'cm'):
        out['PoG_%s_%s' % (unit, out_suffix)] = torch.stack(
            [out['left_PoG_%s_%s' % (unit, out_suffix)],
             out['right_PoG_%s_%s' % (unit, out_suffix)]], -1).mean(-1)
    out['PoG_mm_' + out_suffix] = 10.0 * out['PoG_cm_' + out_suffix]
    out['g_' + out_suffix] = combined_gaze_direction(d['o'], out['PoG_mm_' + out_suffix],
                                                     d['left_R'], d['camera_transformation'])
    if cfg.refine_net_enabled:
        out['heatmap_' + out_suffix] = make_heatmaps(out['PoG_px_' + out_suffix], sigma)


def eve_forward(sd, cfg, inputs, training, kappas=None, with_history=False):
    """EVE.forward (eve.py:69-284) over whole clips.

    Returns (outputs, intermediates): ``outputs`` holds every loss_/metric_ scalar,
    full_loss and the pupil sizes; ``intermediates`` the B x T x ... tensors the
    reference stacks at eve.py:175-182."""
    d = derive_labels(inputs, cfg, training, kappas)
    mid = {}
    for side in ('left', 'right'):
        g, p = eye_net_sequence(sd, cfg, d[side + '_eye_patch'], d[side + '_h'])
        if cfg.eye_net_frozen:
            g = g.detach()
        mid[side + '_g_initial'] = g
        mid[side + '_pupil_size'] = p
    has_geometry = 'inv_camera_transformation' in d
    augment = training and cfg.refine_net_do_offset_augmentation
    if augment:
        if has_geometry:
            _pog_bundle(d, mid, 'initial', 'initial_unaugmented', cfg,
                        cfg.gaze_heatmap_sigma_initial)
        for side in ('left', 'right'):
            mid[side + '_g_initial_unaugmented'] = mid[side + '_g_initial']
            mid[side + '_g_initial'] = offset_augmentation(
                mid[side + '_g_initial'], d['head_R'], d[side + '_kappa_fake'])
        if has_geometry:
            _pog_bundle(d, mid, 'initial', 'initial_augmented', cfg,
                        cfg.gaze_heatmap_sigma_initial)
    if has_geometry:
        _pog_bundle(d, mid, 'initial', 'initial', cfg, cfg.gaze_heatmap_sigma_initial)
        if with_history and cfg.refine_net_enabled and 'PoG_px_tobii' in d:
            hist = make_heatmaps(mid['PoG_px_initial'], cfg.gaze_heatmap_sigma_history)
            mid['history_initial'] = gaze_history_maps(
                d['timestamps'], hist, d['PoG_px_tobii_validity'],
                cfg.gaze_history_map_decay_per_ms)
    if cfg.refine_net_enabled:
        hm = refine_net_sequence(sd, cfg, d.get('screen_frame'), mid['heatmap_initial'])
        mid['heatmap_final'] = hm
        B, T = hm.shape[:2]
        mid['PoG_px_final'] = soft_argmax(hm.reshape(B * T, *hm.shape[2:])).reshape(B, T, 2)
        mid['PoG_cm_final'] = mid['PoG_px_final'] * (0.1 * d['millimeters_per_pixel'])
        mid['g_final'] = combined_gaze_direction(d['o'], 10.0 * mid['PoG_cm_final'],
                                                 d['left_R'], d['camera_transformation'])
        if with_history and 'PoG_px_tobii' in d:
            mid['refined_gaze_history'] = gaze_history_maps(
                d['timestamps'], hm, d['PoG_px_tobii_validity'],
                cfg.gaze_history_map_decay_per_ms)[:, -1]
    out = {'left_pupil_size': mid['left_pupil_size'], 'right_pupil_size': mid['right_pupil_size']}
    eve_losses(d, mid, out, cfg, training)
    return out, mid


def eve_losses(d, mid, out, cfg, training):
    """EVE.calculate_losses_and_metrics + the weighted sum (eve.py:286-439, 231-265)."""
    augment = training and cfg.refine_net_do_offset_augmentation

    def apply(fn, pred_key, gt_key, ref=None):
        ref = d if ref is None else ref
        return masked_clip_mean(fn(mid[pred_key], ref[gt_key]), ref[gt_key + '_validity'])

    for side in ('left', 'right'):
        src = side + ('_g_initial_unaugmented' if augment else '_g_initial')
        if src in mid and side + '_g_tobii' in d:
            out['loss_ang_%s_g_initial' % side] = apply(angular_error, src, side + '_g_tobii')
        src = side + ('_PoG_cm_initial_unaugmented' if augment else '_PoG_cm_initial')
        if src in mid and side + '_PoG_cm_tobii' in d:
            out['loss_mse_%s_PoG_cm_initial' % side] = apply(mse_per_frame, src,
                                                             side + '_PoG_cm_tobii')
            out['metric_euc_%s_PoG_cm_initial' % side] = apply(euclidean_per_frame, src,
                                                               side + '_PoG_cm_tobii')
        if side + '_PoG_px_initial' in mid and side + '_PoG_tobii' in d:
            out['metric_euc_%s_PoG_px_initial' % side] = apply(
                euclidean_per_frame, side + '_PoG_px_initial', side + '_PoG_tobii')
        if side + '_pupil_size' in mid and side + '_p' in d:
            out['loss_l1_%s_pupil_size' % side] = apply(l1_per_frame, side + '_pupil_size',
                                                        side + '_p')
    if 'left_PoG_tobii' in d and 'right_PoG_tobii' in d and 'left_PoG_cm_initial' in mid:
        ref = {'right_PoG_cm_initial': mid['right_PoG_cm_initial'],
               'right_PoG_cm_initial_validity':
                   d['left_PoG_tobii_validity'] & d['right_PoG_tobii_validity']}
        out['loss_mse_lr_consistency'] = apply(mse_per_frame, 'left_PoG_cm_initial',
                                               'right_PoG_cm_initial', ref)
        out['metric_euc_lr_consistency'] = apply(euclidean_per_frame, 'left_PoG_cm_initial',
                                                 'right_PoG_cm_initial', ref)
    src = 'heatmap_initial_unaugmented' if augment else 'heatmap_initial'
    if src in mid and 'heatmap_initial' in d:
        out['loss_ce_heatmap_initial'] = apply(bce_per_frame, src, 'heatmap_initial')
    if 'heatmap_final' in mid and 'heatmap_final' in d:
        out['loss_ce_heatmap_final'] = apply(bce_per_frame, 'heatmap_final', 'heatmap_final')
        out['loss_mse_heatmap_final'] = apply(mse_per_frame, 'heatmap_final', 'heatmap_final')
    stages = ['initial', 'final']
    if cfg.refine_net_do_offset_augmentation:
        stages.insert(0, 'initial_unaugmented')
    for stage in stages:
        for unit in ('px', 'cm'):
            k = 'PoG_%s_%s' % (unit, stage)
            if k in mid and 'PoG_%s_tobii' % unit in d:
                if stage != 'initial_unaugmented':
                    out['loss_mse_' + k] = apply(mse_per_frame, k, 'PoG_%s_tobii' % unit)
                out['metric_euc_' + k] = apply(euclidean_per_frame, k, 'PoG_%s_tobii' % unit)
        if 'g_' + stage in mid and 'g' in d:
            out['metric_ang_g_' + stage] = apply(angular_error, 'g_' + stage, 'g')

    total = torch.zeros((), dtype=mid['left_g_initial'].dtype, device=mid['left_g_initial'].device)
    if 'loss_ang_left_g_initial' in out:
        total = total + cfg.loss_coeff_g_ang_initial * (
            out['loss_ang_left_g_initial'] + out['loss_ang_right_g_initial'])
    if 'loss_mse_left_PoG_cm_initial' in out and cfg.loss_coeff_PoG_cm_initial > 0.0:
        total = total + cfg.loss_coeff_PoG_cm_initial * (
            out['loss_mse_left_PoG_cm_initial'] + out['loss_mse_right_PoG_cm_initial'])
    if 'loss_l1_left_pupil_size' in out:
        total = total + cfg.loss_coeff_pupil_size * (
            out['loss_l1_left_pupil_size'] + out['loss_l1_right_pupil_size'])
    if 'loss_mse_PoG_cm_final' in out:
        total = total + cfg.loss_coeff_PoG_cm_final * out['loss_mse_PoG_cm_final']
    if 'loss_ce_heatmap_initial' in out:
        total = total + cfg.loss_coeff_heatmap_ce_initial * out['loss_ce_heatmap_initial']
    if 'loss_ce_heatmap_final' in out:
        total = total + cfg.loss_coeff_heatmap_ce_final * out['loss_ce_heatmap_final']
    if 'loss_mse_heatmap_final' in out:
        total = total + cfg.loss_coeff_heatmap_mse_final * out['loss_mse_heatmap_final']
    out['full_loss'] = total
    return out
