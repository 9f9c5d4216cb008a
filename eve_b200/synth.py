"""Seeded synthetic clips and weights for the EVE hot path.

The key set and tensor layout follow what the reference's dataset class hands to
``EVE.forward`` (reference: src/datasources/eve_sequences.py:215-335, DATASET.md:65-92):
every entry is ``B x T x ...``, frames are NCHW fp32, eye patches lie in [-1, 1]
(eve_sequences.py:196-203), screen frames in [0, 1] (eve_sequences.py:205-211).

Everything is drawn from ``numpy.random.RandomState`` so a seed reproduces the same
bytes on every machine and torch version (the golden fixtures under tests/golden rely
on that).  Nothing in here touches a GPU; callers move tensors where they want them.
"""
import numpy as np
import torch

ACTUAL_SCREEN = (1920.0, 1080.0)
SCREEN_MM = (553.0, 311.0)


def _small_rotations(rs, n, scale):
    """n random rotation matrices a few degrees away from identity (Rodrigues)."""
    rvec = rs.normal(scale=scale, size=(n, 3))
    theta = np.linalg.norm(rvec, axis=1, keepdims=True) + 1e-12
    k = rvec / theta
    K = np.zeros((n, 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    s = np.sin(theta)[:, :, None]
    c = np.cos(theta)[:, :, None]
    return np.eye(3)[None] + s * K + (1.0 - c) * (K @ K)


def make_clip_batch(B, T, seed=0, with_screen=True, with_labels=True, pad_last=0):
    """Return an ordered dict[str, torch.Tensor] of one synthetic batch of clips.

    ``pad_last`` > 0 zero-pads (zero data, zero validity, zero timestamps) the last
    ``pad_last`` frames of the *last* clip, the way the dataset pads short clips
    (eve_sequences.py:288-297).
    """
    rs = np.random.RandomState(seed)
    f32 = np.float32
    d = {}
    # The first entry must be B x T x ... : EVE.forward reads T from it (eve.py:90).
    for side in ('left', 'right'):
        d[side + '_eye_patch'] = rs.uniform(-1, 1, size=(B, T, 3, 128, 128)).astype(f32)
    if with_screen:
        d['screen_frame'] = rs.uniform(0, 1, size=(B, T, 3, 72, 128)).astype(f32)

    head_R = _small_rotations(rs, B * T, 0.05).reshape(B, T, 3, 3)
    d['head_R'] = head_R.astype(f32)
    eye_R = _small_rotations(rs, B * T, 0.05).reshape(B, T, 3, 3).astype(f32)
    d['left_R'] = eye_R
    d['right_R'] = eye_R.copy()          # by definition left_R == right_R (eve.py:164)

    # A fixed rigid camera->screen transform per clip and its exact inverse.
    cam = np.zeros((B, T, 4, 4))
    inv = np.zeros((B, T, 4, 4))
    for b in range(B):
        R = _small_rotations(rs, 1, 0.08)[0] @ np.diag([-1.0, 1.0, -1.0])
        t = np.array([SCREEN_MM[0] / 2, -15.0, 10.0]) + rs.normal(scale=4.0, size=3)
        M = np.eye(4)
        M[:3, :3], M[:3, 3] = R, t
        Minv = np.eye(4)
        Minv[:3, :3], Minv[:3, 3] = R.T, -R.T @ t
        inv[b, :] = M          # camera coords -> screen coords
        cam[b, :] = Minv       # screen coords -> camera coords
    d['camera_transformation'] = cam.astype(f32)
    d['inv_camera_transformation'] = inv.astype(f32)

    mmpp = np.array([SCREEN_MM[0] / ACTUAL_SCREEN[0], SCREEN_MM[1] / ACTUAL_SCREEN[1]])
    d['millimeters_per_pixel'] = np.broadcast_to(mmpp, (B, T, 2)).astype(f32).copy()
    d['pixels_per_millimeter'] = np.broadcast_to(1.0 / mmpp, (B, T, 2)).astype(f32).copy()

    for side, sx in (('left', -30.0), ('right', 30.0)):
        o = np.array([sx, 0.0, 600.0]) + rs.normal(scale=5.0, size=(B, T, 3))
        d[side + '_o'] = o.astype(f32)
        d[side + '_o_validity'] = np.ones((B, T), dtype=bool)
        d[side + '_h'] = rs.uniform(-0.2, 0.2, size=(B, T, 2)).astype(f32)
        if with_labels:
            d[side + '_g_tobii'] = rs.uniform(-0.2, 0.2, size=(B, T, 2)).astype(f32)
            d[side + '_g_tobii_validity'] = np.ones((B, T), dtype=bool)
            pog = np.stack([rs.uniform(0, ACTUAL_SCREEN[0], size=(B, T)),
                            rs.uniform(0, ACTUAL_SCREEN[1], size=(B, T))], axis=-1)
            d[side + '_PoG_tobii'] = pog.astype(f32)
            d[side + '_PoG_tobii_validity'] = rs.uniform(size=(B, T)) < 0.9
            d[side + '_p'] = rs.uniform(2.0, 5.0, size=(B, T)).astype(f32)
            d[side + '_p_validity'] = np.ones((B, T), dtype=bool)

    t0 = rs.randint(1, 1 << 30, size=(B, 1)).astype(np.int64) * 1000
    step = (100_000_000 + rs.randint(-2_000_000, 2_000_000, size=(B, T))).astype(np.int64)
    d['timestamps'] = t0 + np.cumsum(step, axis=1)

    if pad_last > 0:
        for k, v in d.items():
            v[B - 1, T - pad_last:] = 0
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in d.items()}


# --------------------------------------------------------------------------------------
# Weights.  Shapes/names are the reference's state_dict (eye_net.py:38-96,
# refine_net.py:180-235, torchvision resnet.py BasicBlock/ResNet); values come from a
# numpy RandomState so that fixtures are reproducible without torch's RNG.
# --------------------------------------------------------------------------------------

def eye_net_param_shapes(cfg):
    nf = cfg.eye_net_rnn_num_features if cfg.eye_net_use_rnn else cfg.eye_net_static_num_features
    s = {'cnn_layers.conv1.weight': (64, 3, 7, 7)}
    cin = 64
    for li, cout in enumerate((64, 128, 256, 512), start=1):
        for bi in range(2):
            p = 'cnn_layers.layer%d.%d.' % (li, bi)
            s[p + 'conv1.weight'] = (cout, cin if bi == 0 else cout, 3, 3)
            s[p + 'conv2.weight'] = (cout, cout, 3, 3)
            if bi == 0 and li > 1:
                s[p + 'downsample.0.weight'] = (cout, cin, 1, 1)
        cin = cout
    s['cnn_layers.fc.weight'] = (nf, 512)
    s['cnn_layers.fc.bias'] = (nf,)
    hp = 2 if cfg.eye_net_use_head_pose_input else 0
    s['fc_common.0.weight'] = (nf, nf + hp)
    s['fc_common.0.bias'] = (nf,)
    s['fc_common.2.weight'] = (nf, nf)
    s['fc_common.2.bias'] = (nf,)
    if cfg.eye_net_use_rnn:
        mult = {'RNN': 1, 'LSTM': 4, 'GRU': 3}[cfg.eye_net_rnn_type]
        for i in range(cfg.eye_net_rnn_num_cells):
            p = 'rnn_cells.%d.' % i
            s[p + 'weight_ih'] = (mult * nf, nf)
            s[p + 'weight_hh'] = (mult * nf, nf)
            s[p + 'bias_ih'] = (mult * nf,)
            s[p + 'bias_hh'] = (mult * nf,)
    else:
        s['static_fc.0.weight'] = (nf, nf)
        s['static_fc.0.bias'] = (nf,)
    s['fc_to_gaze.0.weight'] = (nf, nf)
    s['fc_to_gaze.0.bias'] = (nf,)
    s['fc_to_gaze.2.weight'] = (2, nf)
    s['fc_to_pupil.0.weight'] = (nf, nf)
    s['fc_to_pupil.0.bias'] = (nf,)
    s['fc_to_pupil.2.weight'] = (1, nf)
    s['fc_to_pupil.2.bias'] = (1,)
    return s


def _block_shapes(s, p, ic, oc):
    s[p + 'layers.0.weight'] = (ic,)
    s[p + 'layers.0.bias'] = (ic,)
    s[p + 'layers.2.weight'] = (oc, ic, 3, 3)
    s[p + 'layers.2.bias'] = (oc,)
    s[p + 'layers.3.weight'] = (oc,)
    s[p + 'layers.3.bias'] = (oc,)
    s[p + 'layers.5.weight'] = (oc, oc, 3, 3)
    s[p + 'layers.5.bias'] = (oc,)
    if ic != oc:
        s[p + 'skip_layer.0.weight'] = (ic,)
        s[p + 'skip_layer.0.bias'] = (ic,)
        s[p + 'skip_layer.2.weight'] = (oc, ic, 1, 1)
        s[p + 'skip_layer.2.bias'] = (oc,)


# (channels at this level, channels handed to the level below, number of encoder blocks)
REFINE_LEVELS = ((16, 32, 1), (32, 64, 2), (64, 128, 2), (128, 256, 2), (256, None, 2))


def refine_net_param_shapes(cfg):
    nf = cfg.refine_net_num_features
    in_c = 4 if cfg.load_screen_content else 1
    skip = bool(cfg.refine_net_use_skip_connections)
    s = {'initial.0.weight': (16, in_c, 3, 3), 'initial.0.bias': (16,),
         'initial.1.weight': (16,), 'initial.1.bias': (16,),
         'initial.3.weight': (16, 16, 3, 3), 'initial.3.bias': (16,)}
    prefix = 'network.'
    for lvl, (c, c_below, n_enc) in enumerate(REFINE_LEVELS):
        b_ic = c_below if c_below is not None else nf
        b_oc = b_ic      # every wrapped module maps b_ic -> b_ic channels
        _block_shapes(s, prefix + 'encoder_blocks.0.', c, b_ic)
        for j in range(1, n_enc):
            _block_shapes(s, prefix + 'encoder_blocks.%d.' % j, b_ic, b_ic)
        _block_shapes(s, prefix + 'decoder_blocks.0.', b_oc + (b_ic if skip else 0), c)
        prefix += 'between_module.'
    if cfg.refine_net_use_rnn:
        for i in range(cfg.refine_net_rnn_num_cells):
            p = prefix + 'rnn_cells.%d.' % i
            if cfg.refine_net_rnn_type == 'CRNN':
                s[p + 'cell.weight'] = (nf, 2 * nf, 3, 3)
                s[p + 'cell.bias'] = (nf,)
            elif cfg.refine_net_rnn_type == 'CLSTM':
                s[p + 'gates.weight'] = (4 * nf, 2 * nf, 3, 3)
                s[p + 'gates.bias'] = (4 * nf,)
            elif cfg.refine_net_rnn_type == 'CGRU':
                s[p + 'gates_1.weight'] = (2 * nf, 2 * nf, 3, 3)
                s[p + 'gates_1.bias'] = (2 * nf,)
                s[p + 'gate_2.weight'] = (nf, 2 * nf, 3, 3)
                s[p + 'gate_2.bias'] = (nf,)
    s['final.0.weight'] = (16, 16, 3, 3)
    s['final.0.bias'] = (16,)
    s['final.2.weight'] = (1, 16, 1, 1)
    s['final.2.bias'] = (1,)
    return s


def make_state_dict(shapes, seed, prefix=''):
    """Random fp32 weights for parity work (NOT the reference initialiser).

    Conv/linear weights ~ N(0, 2/fan_in) keeps activations O(1) through the stacks,
    norm gains ~ 1 +- 0.1, biases ~ N(0, 0.05).  The two layers the reference
    zero-initialises (eye_net.py:96, refine_net.py:235) get small non-zero values,
    otherwise g == 0 and heatmap == 0.5 and any comparison is vacuous.
    """
    rs = np.random.RandomState(seed)
    out = {}
    for name, shp in shapes.items():
        if name.endswith('fc_to_gaze.2.weight'):
            w = rs.normal(scale=0.05, size=shp)
        elif name.endswith('final.2.weight'):
            w = rs.normal(scale=0.25, size=shp)
        elif len(shp) >= 2:
            fan_in = int(np.prod(shp[1:]))
            w = rs.normal(scale=np.sqrt(2.0 / fan_in), size=shp)
            if 'rnn_cells' in name or 'fc_to' in name or 'fc_common' in name \
                    or 'static_fc' in name or name.endswith('fc.weight'):
                w = rs.normal(scale=np.sqrt(1.0 / fan_in), size=shp)
        elif name.endswith('.weight'):          # norm gain
            w = 1.0 + rs.normal(scale=0.1, size=shp)
        else:
            w = rs.normal(scale=0.05, size=shp)
        out[prefix + name] = torch.from_numpy(w.astype(np.float32))
    return out
