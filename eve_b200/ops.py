"""torch.autograd.Function wrappers around the C ABI (eve_b200/lib.py).

Each Function owns nothing but pointers: PyTorch allocates the tensors, the current CUDA
stream carries the work, `loss.backward()` (reference: src/core/training.py:489) reaches the
`*_bwd` entry points through these classes.  Under `torch.no_grad()` the activation tape goes
to a reusable scratch buffer instead of a fresh allocation.
"""
import ctypes as C

import torch

from . import lib as L


def _on_tensor_device(fn):
    """Run an autograd.Function static method with the CUDA device of its first CUDA tensor
    argument current (the library launches on the current device / its current stream)."""
    import functools

    @functools.wraps(fn)
    def wrapped(ctx, *args):
        for a in args:
            if torch.is_tensor(a) and a.is_cuda:
                with torch.cuda.device(a.device):
                    return fn(ctx, *args)
        return fn(ctx, *args)
    return wrapped


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _bytes(nbytes, device, keep, tag):
    if keep:
        return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
    return L.workspace(nbytes, device, tag)


# --------------------------------------------------------------------------- conv2d --
class Conv2dFn(torch.autograd.Function):
    """nn.Conv2d on NCHW tensors through eve_conv2d_{fwd,dgrad,wgrad} (used by the stand-alone
    ConvRNN cell modules of models/common.py; the networks call the fused entries instead)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, weight, bias, stride, padding):
        L.require_cuda(x, 'conv2d')
        lib = L.load()
        x, weight, bias = _f32c(x), _f32c(weight), _f32c(bias)
        n, cin, h, w = x.shape
        cout, _, k, _ = weight.shape
        p = L.ConvParams(n, h, w, cin, cout, k, int(stride), int(padding))
        oh = (h + 2 * padding - k) // stride + 1
        ow = (w + 2 * padding - k) // stride + 1
        xh = x.permute(0, 2, 3, 1).contiguous()
        yh = torch.empty((n, oh, ow, cout), dtype=torch.float32, device=x.device)
        nbytes = lib.eve_conv2d_workspace_bytes(C.byref(p))
        if nbytes == 0:
            raise ValueError(L.last_error())
        ws = L.workspace(nbytes, x.device)
        L.check(lib.eve_conv2d_fwd(C.byref(p), L.ptr(xh), L.ptr(weight), L.ptr(bias), L.ptr(yh),
                                   L.ptr(ws), ws.numel(), L.stream_ptr()), 'eve_conv2d_fwd')
        ctx.p = p
        ctx.has_bias = bias is not None
        ctx.save_for_backward(xh, weight)
        return yh.permute(0, 3, 1, 2)

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dy):
        lib = L.load()
        xh, weight = ctx.saved_tensors
        p = ctx.p
        dyh = _f32c(dy.permute(0, 2, 3, 1))
        ws = L.workspace(lib.eve_conv2d_workspace_bytes(C.byref(p)), xh.device)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dxh = torch.empty_like(xh)
            L.check(lib.eve_conv2d_dgrad(C.byref(p), L.ptr(dyh), L.ptr(weight), L.ptr(dxh), L.ptr(ws),
                                         ws.numel(), L.stream_ptr()), 'eve_conv2d_dgrad')
            dx = dxh.permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw = torch.empty_like(weight)
            db = torch.empty(weight.shape[0], device=xh.device) if ctx.has_bias else None
            L.check(lib.eve_conv2d_wgrad(C.byref(p), L.ptr(xh), L.ptr(dyh), L.ptr(dw), L.ptr(db),
                                         L.ptr(ws), ws.numel(), L.stream_ptr()), 'eve_conv2d_wgrad')
        return dx, dw, db, None, None


def conv2d(x, weight, bias=None, stride=1, padding=0):
    return Conv2dFn.apply(x, weight, bias, stride, padding)


# ------------------------------------------------------------------------ EyeNet CNN --
class EyeNetCnnFn(torch.autograd.Function):
    """ResNet-18/InstanceNorm features of eye patches (eye_net.py:106).

    x [N,3,H,W] -> feat [N,nf]; weights in the order of include/eve_b200.h."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, nf, *weights):
        L.require_cuda(x, 'EyeNet CNN')
        lib = L.load()
        x = _f32c(x)
        weights = [_f32c(w) for w in weights]
        assert len(weights) == 22
        n, _, h, w_ = x.shape
        p = L.EyeNetCnnParams(n, int(nf), h, w_)
        keep = any(ctx.needs_input_grad)
        saved = _bytes(lib.eve_eyenet_cnn_saved_bytes(C.byref(p)), x.device, keep, 'eye_saved')
        ws = L.workspace(lib.eve_eyenet_cnn_workspace_bytes(C.byref(p)), x.device)
        feat = torch.empty((n, int(nf)), dtype=torch.float32, device=x.device)
        L.check(lib.eve_eyenet_cnn_fwd(C.byref(p), L.ptr(x), L.ptr_table(weights), L.ptr(feat),
                                       L.ptr(saved), saved.numel(), L.ptr(ws), ws.numel(),
                                       L.stream_ptr()), 'eve_eyenet_cnn_fwd')
        if keep:
            ctx.p = p
            ctx.saved = saved
            ctx.save_for_backward(*weights)
        return feat

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dfeat):
        lib = L.load()
        weights = ctx.saved_tensors
        p = ctx.p
        dfeat = _f32c(dfeat)
        grads = [torch.empty_like(w) if ctx.needs_input_grad[2 + i] else None
                 for i, w in enumerate(weights)]
        ws = L.workspace(lib.eve_eyenet_cnn_workspace_bytes(C.byref(p)), dfeat.device)
        L.check(lib.eve_eyenet_cnn_bwd(C.byref(p), L.ptr(dfeat), L.ptr_table(weights),
                                       L.ptr_table(grads), 0, L.ptr(ctx.saved), ctx.saved.numel(),
                                       L.ptr(ws), ws.numel(), L.stream_ptr()),
                'eve_eyenet_cnn_bwd')
        ctx.saved = None
        return (None, None) + tuple(grads)


# ----------------------------------------------------------------------- EyeNet tail --
class EyeNetTailFn(torch.autograd.Function):
    """fc_common -> RNN cells / static_fc -> gaze + pupil heads over whole sequences
    (eye_net.py:109-140).  cfg = (use_head_pose, rnn_type_code, rnn_cells).

    feat [S,T,nf], head_pose [S,T,2] or None, h0/c0 [cells,S,nf] or None ->
    g [S,T,2], pupil [S,T], hT, cT ([cells,S,nf] or None)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, feat, head_pose, h0, c0, cfg, *weights):
        L.require_cuda(feat, 'EyeNet tail')
        lib = L.load()
        use_hp, rnn_type, cells = cfg
        feat, head_pose, h0, c0 = _f32c(feat), _f32c(head_pose), _f32c(h0), _f32c(c0)
        weights = [_f32c(w) for w in weights]
        S, T, nf = feat.shape
        p = L.EyeNetTailParams(S, T, nf, int(bool(use_hp)), rnn_type, cells)
        nw = lib.eve_eyenet_tail_num_weights(C.byref(p))
        if nw < 0:
            raise ValueError(L.last_error())
        assert nw == len(weights), (nw, len(weights))
        keep = any(ctx.needs_input_grad)
        saved = _bytes(lib.eve_eyenet_tail_saved_bytes(C.byref(p)), feat.device, keep, 'tail_saved')
        ws = L.workspace(lib.eve_eyenet_tail_workspace_bytes(C.byref(p)), feat.device)
        dev = feat.device
        g = torch.empty((S, T, 2), dtype=torch.float32, device=dev)
        pupil = torch.empty((S, T), dtype=torch.float32, device=dev)
        hT = cT = None
        if rnn_type != 0:
            hT = torch.empty((cells, S, nf), dtype=torch.float32, device=dev)
            if rnn_type == L.EYE_RNN_TYPES['LSTM']:
                cT = torch.empty((cells, S, nf), dtype=torch.float32, device=dev)
        L.check(lib.eve_eyenet_tail_fwd(C.byref(p), L.ptr(feat), L.ptr(head_pose), L.ptr(h0),
                                        L.ptr(c0), L.ptr_table(weights), L.ptr(g), L.ptr(pupil),
                                        L.ptr(hT), L.ptr(cT), L.ptr(saved), saved.numel(),
                                        L.ptr(ws), ws.numel(), L.stream_ptr()),
                'eve_eyenet_tail_fwd')
        if keep:
            ctx.p = p
            ctx.saved = saved
            ctx.has_h0 = h0 is not None
            ctx.has_c0 = c0 is not None
            ctx.save_for_backward(*weights)
        ctx.set_materialize_grads(False)
        return g, pupil, hT, cT

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dg, dpupil, dhT, dcT):
        lib = L.load()
        weights = ctx.saved_tensors
        p = ctx.p
        dev = weights[0].device
        S, T, nf = p.batch, p.steps, p.nf
        dg = _f32c(dg) if dg is not None else torch.zeros((S, T, 2), device=dev)
        dpupil = _f32c(dpupil) if dpupil is not None else torch.zeros((S, T), device=dev)
        dhT, dcT = _f32c(dhT), _f32c(dcT)
        dfeat = torch.empty((S, T, nf), dtype=torch.float32, device=dev)
        dh0 = torch.empty((p.rnn_cells, S, nf), device=dev) \
            if (ctx.has_h0 and ctx.needs_input_grad[2]) else None
        dc0 = torch.empty((p.rnn_cells, S, nf), device=dev) \
            if (ctx.has_c0 and ctx.needs_input_grad[3]) else None
        grads = [torch.empty_like(w) if ctx.needs_input_grad[5 + i] else None
                 for i, w in enumerate(weights)]
        ws = L.workspace(lib.eve_eyenet_tail_workspace_bytes(C.byref(p)), dev)
        L.check(lib.eve_eyenet_tail_bwd(C.byref(p), L.ptr(dg), L.ptr(dpupil), L.ptr(dhT),
                                        L.ptr(dcT), L.ptr_table(weights), L.ptr(dfeat), L.ptr(dh0),
                                        L.ptr(dc0), L.ptr_table(grads), 0, L.ptr(ctx.saved),
                                        ctx.saved.numel(), L.ptr(ws), ws.numel(), L.stream_ptr()),
                'eve_eyenet_tail_bwd')
        ctx.saved = None
        return (dfeat, None, dh0, dc0, None) + tuple(grads)


# ------------------------------------------------------------------------- RefineNet --
def refinenet_weight_names(p):
    lib = L.load()
    n = lib.eve_refinenet_num_weights(C.byref(p))
    if n < 0:
        raise ValueError(L.last_error())
    return [lib.eve_refinenet_weight_name(C.byref(p), i).decode() for i in range(n)]


class RefineNetFn(torch.autograd.Function):
    """RefineNet.forward over whole sequences (refine_net.py:237-255).
    cfg = (in_channels, use_skip, rnn_type_code, rnn_cells, nf).

    screen [B,T,3,72,128] or None, heatmap [B,T,1,72,128], h0/c0 [cells,B,nf,5,8] or None ->
    out [B,T,1,72,128], hT, cT."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, screen, heatmap, h0, c0, cfg, *weights):
        L.require_cuda(heatmap, 'RefineNet')
        lib = L.load()
        in_c, use_skip, rnn_type, cells, nf = cfg
        screen, heatmap, h0, c0 = _f32c(screen), _f32c(heatmap), _f32c(h0), _f32c(c0)
        weights = [_f32c(w) for w in weights]
        B, T = heatmap.shape[:2]
        if tuple(heatmap.shape[2:]) != (1, 72, 128):
            raise ValueError('RefineNet needs 72x128 heatmaps (screen_size [128, 72]), got %s'
                             % (tuple(heatmap.shape),))
        p = L.RefineNetParams(B, T, in_c, int(bool(use_skip)), rnn_type, cells, nf)
        nw = lib.eve_refinenet_num_weights(C.byref(p))
        if nw < 0:
            raise ValueError(L.last_error())
        assert nw == len(weights), (nw, len(weights))
        keep = any(ctx.needs_input_grad)
        dev = heatmap.device
        saved = _bytes(lib.eve_refinenet_saved_bytes(C.byref(p)), dev, keep, 'refine_saved')
        ws = L.workspace(lib.eve_refinenet_workspace_bytes(C.byref(p)), dev)
        out = torch.empty((B, T, 1, 72, 128), dtype=torch.float32, device=dev)
        hT = cT = None
        if rnn_type != 0:
            hT = torch.empty((cells, B, nf, 5, 8), dtype=torch.float32, device=dev)
            if rnn_type == L.REFINE_RNN_TYPES['CLSTM']:
                cT = torch.empty((cells, B, nf, 5, 8), dtype=torch.float32, device=dev)
        L.check(lib.eve_refinenet_fwd(C.byref(p), L.ptr(screen), L.ptr(heatmap), L.ptr(h0),
                                      L.ptr(c0), L.ptr_table(weights), L.ptr(out), L.ptr(hT),
                                      L.ptr(cT), L.ptr(saved), saved.numel(), L.ptr(ws), ws.numel(),
                                      L.stream_ptr()), 'eve_refinenet_fwd')
        if keep:
            ctx.p = p
            ctx.saved = saved
            ctx.has_h0 = h0 is not None
            ctx.save_for_backward(*weights)
        ctx.set_materialize_grads(False)
        return out, hT, cT

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dout, dhT, dcT):
        lib = L.load()
        weights = ctx.saved_tensors
        p = ctx.p
        dev = weights[0].device
        B, T = p.batch, p.steps
        dout = _f32c(dout) if dout is not None else torch.zeros((B, T, 1, 72, 128), device=dev)
        if p.rnn_type == L.REFINE_RNN_TYPES['CLSTM']:
            dhT = dcT = None   # the CLSTM state never reaches the heatmap (refine_net.py:168-174)
        dhT, dcT = _f32c(dhT), _f32c(dcT)
        dheat = torch.empty((B, T, 1, 72, 128), device=dev) if ctx.needs_input_grad[1] else None
        dh0 = torch.empty((p.rnn_cells, B, p.nf, 5, 8), device=dev) \
            if (ctx.has_h0 and ctx.needs_input_grad[2]) else None
        grads = [torch.empty_like(w) if ctx.needs_input_grad[5 + i] else None
                 for i, w in enumerate(weights)]
        ws = L.workspace(lib.eve_refinenet_workspace_bytes(C.byref(p)), dev)
        L.check(lib.eve_refinenet_bwd(C.byref(p), L.ptr(dout), L.ptr(dhT), L.ptr(dcT),
                                      L.ptr_table(weights), L.ptr(dheat), L.ptr(dh0), None,
                                      L.ptr_table(grads), 0, L.ptr(ctx.saved), ctx.saved.numel(),
                                      L.ptr(ws), ws.numel(), L.stream_ptr()), 'eve_refinenet_bwd')
        ctx.saved = None
        return (None, dheat, dh0, None, None) + tuple(grads)


# --------------------------------------------------------------------- heatmap / PoG --
class HeatmapFn(torch.autograd.Function):
    """batch_make_heatmaps (common.py:226-243): centres [n,2] px -> [n,1,H,W]."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, centres, sigma, size_wh, screen_wh):
        L.require_cuda(centres, 'make_heatmap')
        lib = L.load()
        centres = _f32c(centres)
        n = centres.shape[0]
        p = L.HeatmapParams(n, int(size_wh[0]), int(size_wh[1]), float(screen_wh[0]),
                            float(screen_wh[1]), float(sigma))
        out = torch.empty((n, 1, int(size_wh[1]), int(size_wh[0])), dtype=torch.float32,
                          device=centres.device)
        L.check(lib.eve_heatmap_fwd(C.byref(p), L.ptr(centres), L.ptr(out), L.stream_ptr()),
                'eve_heatmap_fwd')
        ctx.p = p
        ctx.save_for_backward(centres)
        return out

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dout):
        lib = L.load()
        centres, = ctx.saved_tensors
        dout = _f32c(dout)
        dc = torch.empty_like(centres)
        L.check(lib.eve_heatmap_bwd(C.byref(ctx.p), L.ptr(centres), L.ptr(dout), L.ptr(dc),
                                    L.stream_ptr()), 'eve_heatmap_bwd')
        return dc, None, None, None


class SoftArgmaxFn(torch.autograd.Function):
    """soft_argmax (common.py:294-323): heatmaps [n,1,H,W] -> PoG [n,2] px."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, heatmaps, screen_wh):
        L.require_cuda(heatmaps, 'soft_argmax')
        lib = L.load()
        heatmaps = _f32c(heatmaps)
        n, _, h, w = heatmaps.shape
        p = L.HeatmapParams(n, w, h, float(screen_wh[0]), float(screen_wh[1]), 1.0)
        out = torch.empty((n, 2), dtype=torch.float32, device=heatmaps.device)
        L.check(lib.eve_soft_argmax_fwd(C.byref(p), L.ptr(heatmaps), L.ptr(out), L.stream_ptr()),
                'eve_soft_argmax_fwd')
        ctx.p = p
        ctx.save_for_backward(heatmaps)
        return out

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dpog):
        lib = L.load()
        heatmaps, = ctx.saved_tensors
        dpog = _f32c(dpog)
        dh = torch.empty_like(heatmaps)
        L.check(lib.eve_soft_argmax_bwd(C.byref(ctx.p), L.ptr(heatmaps), L.ptr(dpog), L.ptr(dh),
                                        L.stream_ptr()), 'eve_soft_argmax_bwd')
        return dh, None


class PogFn(torch.autograd.Function):
    """to_screen_coordinates (common.py:149-179): gradient flows to the gaze direction only
    (origin, rotation and calibration are data)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, origin, g, rot, inv_cam, ppm, screen_wh):
        L.require_cuda(g, 'to_screen_coordinates')
        lib = L.load()
        origin, g, rot, inv_cam, ppm = (_f32c(t) for t in (origin, g, rot, inv_cam, ppm))
        n = g.shape[0]
        mm = torch.empty((n, 2), dtype=torch.float32, device=g.device)
        px = torch.empty((n, 2), dtype=torch.float32, device=g.device)
        L.check(lib.eve_pog_fwd(n, L.ptr(origin), L.ptr(g), L.ptr(rot), L.ptr(inv_cam), L.ptr(ppm),
                                float(screen_wh[0]), float(screen_wh[1]), L.ptr(mm), L.ptr(px),
                                L.stream_ptr()), 'eve_pog_fwd')
        ctx.screen = (float(screen_wh[0]), float(screen_wh[1]))
        ctx.save_for_backward(origin, g, rot, inv_cam, ppm)
        ctx.set_materialize_grads(False)
        return mm, px

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dmm, dpx):
        lib = L.load()
        origin, g, rot, inv_cam, ppm = ctx.saved_tensors
        if dmm is None and dpx is None:
            return None, None, None, None, None, None
        dmm, dpx = _f32c(dmm), _f32c(dpx)
        dg = torch.empty_like(g)
        L.check(lib.eve_pog_bwd(g.shape[0], L.ptr(origin), L.ptr(g), L.ptr(rot), L.ptr(inv_cam),
                                L.ptr(ppm), ctx.screen[0], ctx.screen[1], L.ptr(dmm), L.ptr(dpx),
                                L.ptr(dg), L.stream_ptr()), 'eve_pog_bwd')
        return None, dg, None, None, None, None


class CombinedGazeFn(torch.autograd.Function):
    """calculate_combined_gaze_direction (common.py:129-146): one kernel, one thread per frame.
    Gradient flows to the averaged PoG only (origin, rotation and calibration are data)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, origin, pog_mm, head_rot, cam):
        L.require_cuda(pog_mm, 'calculate_combined_gaze_direction')
        lib = L.load()
        origin, pog_mm, head_rot, cam = (_f32c(t) for t in (origin, pog_mm, head_rot, cam))
        n = pog_mm.shape[0]
        g = torch.empty((n, 2), dtype=torch.float32, device=pog_mm.device)
        L.check(lib.eve_combined_gaze_fwd(n, L.ptr(origin), L.ptr(pog_mm), L.ptr(head_rot),
                                          L.ptr(cam), L.ptr(g), L.stream_ptr()),
                'eve_combined_gaze_fwd')
        ctx.save_for_backward(origin, pog_mm, head_rot, cam)
        return g

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dg):
        lib = L.load()
        origin, pog_mm, head_rot, cam = ctx.saved_tensors
        dg = _f32c(dg)
        dp = torch.empty_like(pog_mm)
        L.check(lib.eve_combined_gaze_bwd(pog_mm.shape[0], L.ptr(origin), L.ptr(pog_mm),
                                          L.ptr(head_rot), L.ptr(cam), L.ptr(dg), L.ptr(dp),
                                          L.stream_ptr()), 'eve_combined_gaze_bwd')
        return None, dp, None, None


class OffsetAugmentationFn(torch.autograd.Function):
    """apply_offset_augmentation (common.py:182-218).  ``kappa`` is [n / frames_per_kappa, 2]: one
    kappa per clip repeated over time (eve.py:466-477) needs no expanded copy."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, g, head_rot, kappa, frames_per_kappa, inverse):
        L.require_cuda(g, 'apply_offset_augmentation')
        lib = L.load()
        g, head_rot, kappa = _f32c(g), _f32c(head_rot), _f32c(kappa)
        n = g.shape[0]
        assert kappa.shape[0] * int(frames_per_kappa) == n
        out = torch.empty_like(g)
        L.check(lib.eve_offset_augmentation_fwd(n, int(frames_per_kappa), L.ptr(g), L.ptr(head_rot),
                                                L.ptr(kappa), int(bool(inverse)), L.ptr(out),
                                                L.stream_ptr()), 'eve_offset_augmentation_fwd')
        ctx.args = (int(frames_per_kappa), int(bool(inverse)))
        ctx.save_for_backward(g, head_rot, kappa)
        return out

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dout):
        lib = L.load()
        g, head_rot, kappa = ctx.saved_tensors
        dout = _f32c(dout)
        dg = torch.empty_like(g)
        L.check(lib.eve_offset_augmentation_bwd(g.shape[0], ctx.args[0], L.ptr(g), L.ptr(head_rot),
                                                L.ptr(kappa), ctx.args[1], L.ptr(dout), L.ptr(dg),
                                                L.stream_ptr()), 'eve_offset_augmentation_bwd')
        return dg, None, None, None, None


def _u8(t):
    """Validity flags as a contiguous byte tensor (torch.bool is one byte per element)."""
    if t.dtype == torch.bool:
        return t.contiguous().view(torch.uint8)
    return (t != 0).contiguous().view(torch.uint8)


def frame_labels(d):
    """The per-frame block of EVE.calculate_additional_labels (eve.py:449-456, 498-517, 534-543)
    as one kernel; returns the dict entries it derives (no gradients: labels)."""
    lib = L.load()
    lp = _f32c(d['left_PoG_tobii'])
    L.require_cuda(lp, 'calculate_additional_labels')
    lead = lp.shape[:-1]
    n = lp.numel() // 2
    dev = lp.device
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    out = {'left_PoG_cm_tobii': f(*lead, 2), 'right_PoG_cm_tobii': f(*lead, 2), 'o': f(*lead, 3),
           'PoG_px_tobii': f(*lead, 2), 'PoG_cm_tobii': f(*lead, 2), 'g': f(*lead, 2)}
    valid = torch.empty(lead, dtype=torch.bool, device=dev)
    keep = [_f32c(d[k]) for k in ('right_PoG_tobii', 'millimeters_per_pixel', 'left_o', 'right_o',
                                  'left_R', 'camera_transformation')]
    lv, rv = _u8(d['left_PoG_tobii_validity']), _u8(d['right_PoG_tobii_validity'])
    a = L.LabelArgs(n, L.ptr(lp), L.ptr(keep[0]), L.ptr(lv), L.ptr(rv), L.ptr(keep[1]),
                    L.ptr(keep[2]), L.ptr(keep[3]), L.ptr(keep[4]), L.ptr(keep[5]),
                    L.ptr(out['left_PoG_cm_tobii']), L.ptr(out['right_PoG_cm_tobii']),
                    L.ptr(out['o']), L.ptr(out['PoG_px_tobii']), L.ptr(out['PoG_cm_tobii']),
                    valid.data_ptr(), L.ptr(out['g']))
    with torch.cuda.device(dev):
        L.check(lib.eve_labels_fwd(C.byref(a), L.stream_ptr()), 'eve_labels_fwd')
    out['PoG_px_tobii_validity'] = valid
    return out


def heatmap_labels(centres, valid, sigmas, size_wh, screen_wh):
    """eve.py:519-531: the validity-scaled label heatmaps for up to three sigmas, one launch.
    centres [..., 2] px, valid [...] bool -> list of [..., 1, H, W]."""
    lib = L.load()
    centres = _f32c(centres)
    L.require_cuda(centres, 'label heatmaps')
    lead = centres.shape[:-1]
    n = centres.numel() // 2
    w, h = int(size_wh[0]), int(size_wh[1])
    outs = [torch.empty((*lead, 1, h, w), dtype=torch.float32, device=centres.device)
            for _ in sigmas]
    v = _u8(valid)
    p = L.HeatmapParams(n, w, h, float(screen_wh[0]), float(screen_wh[1]), 1.0)
    sg = (C.c_float * len(sigmas))(*[float(s_) for s_ in sigmas])
    with torch.cuda.device(centres.device):
        L.check(lib.eve_heatmap_labels_fwd(C.byref(p), L.ptr(centres), L.ptr(v), len(sigmas), sg,
                                           L.ptr_table(outs), L.stream_ptr()),
                'eve_heatmap_labels_fwd')
    return outs


class GazeHistoryFn(torch.autograd.Function):
    """Gaze-history maps of every prefix (common.py:249-287 evaluated after each step as in
    eve.py:596-601) by the O(T) recurrence of eve_gaze_history_fwd.
    heatmaps [B,T,1,H,W], timestamps [B,T] int64, validity [B,T] -> [B,T,1,H,W]."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, heatmaps, timestamps, validity, decay):
        L.require_cuda(heatmaps, 'gaze history maps')
        lib = L.load()
        heatmaps = _f32c(heatmaps)
        B, T = heatmaps.shape[:2]
        hw = heatmaps[0, 0].numel()
        ts = timestamps.to(torch.int64).contiguous()
        v = _u8(validity)
        out = torch.empty_like(heatmaps)
        ws = L.workspace(lib.eve_gaze_history_scratch_bytes(B, T), heatmaps.device, 'hist')
        L.check(lib.eve_gaze_history_fwd(B, T, hw, L.ptr(ts), L.ptr(v), float(decay),
                                         L.ptr(heatmaps), L.ptr(out), L.ptr(ws), ws.numel(),
                                         L.stream_ptr()), 'eve_gaze_history_fwd')
        ctx.args = (B, T, hw, float(decay))
        ctx.save_for_backward(ts, v)
        return out

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dout):
        lib = L.load()
        ts, v = ctx.saved_tensors
        B, T, hw, decay = ctx.args
        dout = _f32c(dout)
        dh = torch.empty_like(dout)
        ws = L.workspace(lib.eve_gaze_history_scratch_bytes(B, T), dout.device, 'hist')
        L.check(lib.eve_gaze_history_bwd(B, T, hw, L.ptr(ts), L.ptr(v), decay, L.ptr(dout),
                                         L.ptr(dh), L.ptr(ws), ws.numel(), L.stream_ptr()),
                'eve_gaze_history_bwd')
        return dh, None, None, None


# --------------------------------------------------------------------------- losses --
class HeatmapFrameLossFn(torch.autograd.Function):
    """Per-frame BCE and MSE of heatmap pairs (cross_entropy.py:29-35, mse.py) from one read:
    pred / gt [B,T,1,H,W] -> (bce [B,T], mse [B,T])."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, pred, gt):
        L.require_cuda(pred, 'heatmap losses')
        lib = L.load()
        pred, gt = _f32c(pred), _f32c(gt)
        lead = pred.shape[:2]
        n = lead[0] * lead[1]
        hw = pred.numel() // max(n, 1)
        bce = torch.empty(lead, dtype=torch.float32, device=pred.device)
        mse = torch.empty(lead, dtype=torch.float32, device=pred.device)
        L.check(lib.eve_heatmap_frame_losses_fwd(n, hw, L.ptr(pred), L.ptr(gt), L.ptr(bce),
                                                 L.ptr(mse), L.stream_ptr()),
                'eve_heatmap_frame_losses_fwd')
        ctx.args = (n, hw)
        ctx.save_for_backward(pred, gt)
        ctx.set_materialize_grads(False)
        return bce, mse

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dbce, dmse):
        if dbce is None and dmse is None:
            return None, None
        lib = L.load()
        pred, gt = ctx.saved_tensors
        n, hw = ctx.args
        dbce, dmse = _f32c(dbce), _f32c(dmse)
        dp = torch.empty_like(pred)
        L.check(lib.eve_heatmap_frame_losses_bwd(n, hw, L.ptr(pred), L.ptr(gt), L.ptr(dbce),
                                                 L.ptr(dmse), L.ptr(dp), L.stream_ptr()),
                'eve_heatmap_frame_losses_bwd')
        return dp, None


class MaskedLossesFn(torch.autograd.Function):
    """A table of validity-masked sequence losses (base_loss_with_validity.py:32-73) in one
    launch.  ``spec`` is a list of (op, pred_index, gt, valid, valid2); ``preds`` the distinct
    prediction tensors [B,T,dim] / [B,T].  Returns a [len(spec)] vector of scalars."""

    @staticmethod
    def _table(spec, preds, dpreds):
        terms = (L.LossTerm * len(spec))()
        for i, (op, pi, gt, valid, valid2) in enumerate(spec):
            pr = preds[pi]
            dim = 1 if pr.ndim == 2 else pr.shape[-1]
            terms[i] = L.LossTerm(L.LOSS_OPS[op], dim, L.ptr(pr), L.ptr(gt), L.ptr(valid),
                                  L.ptr(valid2), None if dpreds is None or dpreds[pi] is None
                                  else dpreds[pi].data_ptr())
        return terms

    @staticmethod
    @_on_tensor_device
    def forward(ctx, spec, *preds):
        lib = L.load()
        assert 0 < len(spec) <= L.LOSS_MAX_TERMS
        preds = [_f32c(p) for p in preds]
        L.require_cuda(preds[0], 'losses')
        B, T = preds[0].shape[:2]
        spec = [(op, pi, None if gt is None else _f32c(gt), _u8(v), None if v2 is None else _u8(v2))
                for op, pi, gt, v, v2 in spec]
        for op, pi, gt, v, v2 in spec:
            assert preds[pi].shape[:2] == (B, T) and v.shape == (B, T)
            assert gt is None or gt.shape == preds[pi].shape, (op, gt.shape, preds[pi].shape)
        out = torch.empty(len(spec), dtype=torch.float32, device=preds[0].device)
        terms = MaskedLossesFn._table(spec, preds, None)
        L.check(lib.eve_masked_losses_fwd(len(spec), terms, B, T, L.ptr(out), L.stream_ptr()),
                'eve_masked_losses_fwd')
        ctx.spec, ctx.bt = spec, (B, T)
        ctx.save_for_backward(*preds)
        return out

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dout):
        lib = L.load()
        preds = ctx.saved_tensors
        B, T = ctx.bt
        dout = _f32c(dout)
        need = [ctx.needs_input_grad[1 + i] for i in range(len(preds))]
        # one zeroed flat buffer for every gradient, sliced per prediction tensor
        sizes = [p.numel() if nd else 0 for p, nd in zip(preds, need)]
        flat = torch.zeros(max(sum(sizes), 1), dtype=torch.float32, device=dout.device)
        dpreds, off = [], 0
        for p, nd, sz in zip(preds, need, sizes):
            dpreds.append(flat[off:off + sz].view(p.shape) if nd else None)
            off += sz
        terms = MaskedLossesFn._table(ctx.spec, preds, dpreds)
        L.check(lib.eve_masked_losses_bwd(len(ctx.spec), terms, B, T, L.ptr(dout), L.stream_ptr()),
                'eve_masked_losses_bwd')
        return (None,) + tuple(dpreds)


# ------------------------------------------------------------------- fused optimiser --
def adam_clip_step(params, grads, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8,
                   weight_decay=0.0, max_norm=0.0, grad_scale=1.0, step_dev=None, lr_dev=None):
    """clip_grad_norm_ + Adam on flat fp32 buffers (training.py:492-502).  Returns the
    pre-clip gradient norm as a 1-element device tensor.  ``step_dev`` (int32 device tensor)
    makes the kernel keep the step count itself, so the call can live in a CUDA graph;
    ``lr_dev`` (1-element fp32 device tensor) overrides ``lr`` the same way, so that a
    learning-rate schedule keeps working across replays."""
    lib = L.load()
    L.require_cuda(params, 'adam_clip_step')
    p = L.AdamParams(params.numel(), lr, betas[0], betas[1], eps, weight_decay, max_norm,
                     grad_scale, int(step), None if step_dev is None else step_dev.data_ptr(),
                     None if lr_dev is None else lr_dev.data_ptr())
    ws = L.workspace(lib.eve_adam_clip_workspace_bytes(C.byref(p)), params.device, 'adam')
    norm = torch.empty(1, dtype=torch.float32, device=params.device)
    L.check(lib.eve_adam_clip_step(C.byref(p), L.ptr(params), L.ptr(grads), L.ptr(exp_avg),
                                   L.ptr(exp_avg_sq), L.ptr(norm), L.ptr(ws), ws.numel(),
                                   L.stream_ptr()), 'eve_adam_clip_step')
    return norm
