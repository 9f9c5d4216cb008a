"""Thin test-side callers of the C ABI (single ops), NHWC <-> NCHW helpers."""
import ctypes as C

import torch

from eve_b200 import lib as L


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def _ws(p):
    lib = L.load()
    n = lib.eve_conv2d_workspace_bytes(C.byref(p))
    assert n > 0, L.last_error()
    return torch.empty(n, dtype=torch.uint8, device='cuda')


def conv_fwd(x_nchw, w, bias, stride, pad):
    lib = L.load()
    n, cin, h, wd = x_nchw.shape
    cout, _, k, _ = w.shape
    p = L.ConvParams(n, h, wd, cin, cout, k, stride, pad)
    oh = (h + 2 * pad - k) // stride + 1
    ow = (wd + 2 * pad - k) // stride + 1
    y = torch.empty((n, oh, ow, cout), device='cuda')
    ws = _ws(p)
    x, w = nhwc(x_nchw), w.contiguous()          # keep alive: L.ptr() takes raw addresses
    L.check(lib.eve_conv2d_fwd(C.byref(p), L.ptr(x), L.ptr(w),
                               L.ptr(bias), L.ptr(y), L.ptr(ws), ws.numel(), L.stream_ptr()),
            'conv_fwd')
    return nchw(y)


def conv_dgrad(dy_nchw, w, in_hw, stride, pad):
    lib = L.load()
    n, cout, oh, ow = dy_nchw.shape
    _, cin, k, _ = w.shape
    h, wd = in_hw
    p = L.ConvParams(n, h, wd, cin, cout, k, stride, pad)
    dx = torch.empty((n, h, wd, cin), device='cuda')
    ws = _ws(p)
    dy, w = nhwc(dy_nchw), w.contiguous()
    L.check(lib.eve_conv2d_dgrad(C.byref(p), L.ptr(dy), L.ptr(w),
                                 L.ptr(dx), L.ptr(ws), ws.numel(), L.stream_ptr()), 'conv_dgrad')
    return nchw(dx)


def conv_wgrad(x_nchw, dy_nchw, k, stride, pad, with_bias=True):
    lib = L.load()
    n, cin, h, wd = x_nchw.shape
    cout = dy_nchw.shape[1]
    p = L.ConvParams(n, h, wd, cin, cout, k, stride, pad)
    dw = torch.empty((cout, cin, k, k), device='cuda')
    db = torch.empty((cout,), device='cuda') if with_bias else None
    ws = _ws(p)
    x, dy = nhwc(x_nchw), nhwc(dy_nchw)
    L.check(lib.eve_conv2d_wgrad(C.byref(p), L.ptr(x), L.ptr(dy), L.ptr(dw),
                                 L.ptr(db), L.ptr(ws), ws.numel(), L.stream_ptr()), 'conv_wgrad')
    return dw, db


def instnorm_fwd(x_nchw, gamma, beta, act):
    lib = L.load()
    n, c, h, w = x_nchw.shape
    x = nhwc(x_nchw)
    y = torch.empty_like(x)
    mean = torch.empty((n, c), device='cuda')
    rstd = torch.empty((n, c), device='cuda')
    L.check(lib.eve_instnorm_act_fwd(L.ptr(x), n, h * w, c, L.ptr(gamma), L.ptr(beta), act,
                                     L.ptr(y), L.ptr(mean), L.ptr(rstd), L.stream_ptr()),
            'instnorm_fwd')
    return nchw(y), mean, rstd


def instnorm_bwd(dy_nchw, y_nchw, x_nchw, mean, rstd, gamma, act):
    lib = L.load()
    n, c, h, w = x_nchw.shape
    dx = torch.empty((n, h, w, c), device='cuda')
    dgamma = torch.empty(c, device='cuda') if gamma is not None else None
    dbeta = torch.empty(c, device='cuda') if gamma is not None else None
    ws = torch.empty(2 * n * c * 4 + 256, dtype=torch.uint8, device='cuda')
    dy, y, x = nhwc(dy_nchw), nhwc(y_nchw), nhwc(x_nchw)
    L.check(lib.eve_instnorm_act_bwd(L.ptr(dy), L.ptr(y), L.ptr(x),
                                     n, h * w, c, L.ptr(mean), L.ptr(rstd), L.ptr(gamma), act,
                                     L.ptr(dx), L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws), ws.numel(),
                                     L.stream_ptr()), 'instnorm_bwd')
    return nchw(dx), dgamma, dbeta


def adaptive_maxpool(x_nchw, oh, ow):
    lib = L.load()
    n, c, h, w = x_nchw.shape
    y = torch.empty((n, oh, ow, c), device='cuda')
    idx = torch.empty((n, oh, ow, c), dtype=torch.int32, device='cuda')
    x = nhwc(x_nchw)
    L.check(lib.eve_adaptive_maxpool_fwd(L.ptr(x), n, h, w, c, oh, ow, L.ptr(y),
                                         L.ptr(idx), L.stream_ptr()), 'amp_fwd')
    return nchw(y), idx.permute(0, 3, 1, 2).contiguous()


def adaptive_maxpool_bwd(dy_nchw, idx_nchw, h, w):
    lib = L.load()
    n, c, oh, ow = dy_nchw.shape
    dx = torch.empty((n, h, w, c), device='cuda')
    idx = idx_nchw.permute(0, 2, 3, 1).contiguous()
    dy = nhwc(dy_nchw)
    L.check(lib.eve_adaptive_maxpool_bwd(L.ptr(dy), L.ptr(idx), n, h, w, c, oh, ow,
                                         L.ptr(dx), L.stream_ptr()), 'amp_bwd')
    return nchw(dx)


def upsample(x_nchw, oh, ow):
    lib = L.load()
    n, c, h, w = x_nchw.shape
    y = torch.empty((n, oh, ow, c), device='cuda')
    x = nhwc(x_nchw)
    L.check(lib.eve_upsample_bilinear_fwd(L.ptr(x), n, h, w, c, oh, ow, L.ptr(y),
                                          L.stream_ptr()), 'up_fwd')
    return nchw(y)


def upsample_bwd(dy_nchw, h, w):
    lib = L.load()
    n, c, oh, ow = dy_nchw.shape
    dx = torch.empty((n, h, w, c), device='cuda')
    dy = nhwc(dy_nchw)
    L.check(lib.eve_upsample_bilinear_bwd(L.ptr(dy), n, h, w, c, oh, ow, L.ptr(dx),
                                          L.stream_ptr()), 'up_bwd')
    return nchw(dx)


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
