"""A/B of the padded-strip kernel against the per-tap box kernel at the bench geometries:
accuracy of both against an fp64 convolution and kernel time from the library profiler
(CUDA events around the convolution kernel only).  Usage: python tools/probe_strip.py [reps]"""
import ctypes as C
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from eve_b200 import lib as L        # noqa: E402
from tests import gpu_util as G      # noqa: E402

lib = L.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
import os
GEOMS = [(480, 32, 32, 64, 64), (240, 36, 64, 64, 64), (240, 9, 16, 256, 256), (240, 18, 32, 128, 128),
         (480, 16, 16, 128, 128), (480, 8, 8, 256, 256), (480, 4, 4, 512, 512), (240, 36, 64, 128, 32),
         (240, 36, 64, 32, 64), (240, 36, 64, 32, 32), (240, 9, 16, 512, 128), (240, 18, 32, 256, 64),
         (240, 18, 32, 64, 128), (240, 9, 16, 128, 256), (240, 5, 8, 256, 256), (240, 18, 32, 64, 64),
         (240, 9, 16, 128, 128), (240, 5, 8, 128, 256), (3, 7, 10, 64, 64), (5, 9, 16, 32, 32)]
if os.environ.get('PROBE_STRIP_GEOMS'):
    GEOMS = [tuple(int(v) for v in t.split('x')) for t in os.environ['PROBE_STRIP_GEOMS'].split(',')]


def timed(x, w, b, opt):
    L.set_option('tc_strip', opt)
    y = G.conv_fwd(x, w, b, 1, 1)          # warm-up + result
    torch.cuda.synchronize()
    lib.eve_profile_reset()
    lib.eve_profile_enable(1)
    for _ in range(reps):
        G.conv_fwd(x, w, b, 1, 1)
    torch.cuda.synchronize()
    lib.eve_profile_enable(0)
    v = [C.c_double(), C.c_double(), C.c_double(), C.c_longlong()]
    L.check(lib.eve_profile_read(0, C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), C.byref(v[3])), 'read')
    lib.eve_profile_reset()
    return y, v[0].value / max(v[3].value, 1), v[1].value / max(v[0].value, 1e-9) * 1e-9


print('%-26s %9s %8s %9s %8s %8s %9s %9s  plan' % ('geometry', 'box us', 'TF/s', 'strip us', 'TF/s', 'speedup',
                                                    'err box', 'err strip'))
for n, h, wd, ci, co in GEOMS:
    g = torch.Generator(device='cuda').manual_seed(n + ci + co)
    x = torch.randn(n, ci, h, wd, generator=g, device='cuda')
    w = torch.randn(co, ci, 3, 3, generator=g, device='cuda') / (ci * 9) ** 0.5
    b = torch.randn(co, generator=g, device='cuda')
    want = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    y0, t0, f0 = timed(x, w, b, 0)
    y2, t2, f2 = timed(x, w, b, 2)
    den = float(want.abs().max())
    e0 = float((y0.double() - want).abs().max()) / den
    e2 = float((y2.double() - want).abs().max()) / den
    p = L.ConvParams(n, h, wd, ci, co, 3, 1, 1)
    buf = C.create_string_buffer(256)
    lib.eve_conv2d_describe(C.byref(p), buf, 256)
    print('%4dx%3dx%-3d %4d->%-4d %9.1f %8.1f %9.1f %8.1f %8.2f %9.1e %9.1e  %s' % (
        n, h, wd, ci, co, t0 * 1e3, f0, t2 * 1e3, f2, t0 / t2, e0, e2, buf.value.decode()[:60]))
L.set_option('tc_strip', 1)
