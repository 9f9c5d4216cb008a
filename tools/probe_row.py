"""GPU probe: halo-row conv kernel variants (descriptor shift with / without base offset,
pre-shifted copies), mixed-format wgrad, and their kernel times at the bench geometry.
Prints a table; asserts nothing (tests/test_gpu_conv_tc.py is the gate)."""
import ctypes as C
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from eve_b200 import lib as L
from tests import gpu_util as G

lib = L.load()
lib.eve_set_conv_mode(1)


def prof_ms(kind, fn, iters=3):
    fn(); torch.cuda.synchronize()
    lib.eve_profile_reset(); lib.eve_profile_enable(1)
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    lib.eve_profile_enable(0)
    v = [C.c_double(), C.c_double(), C.c_double(), C.c_longlong()]
    lib.eve_profile_read(kind, C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), C.byref(v[3]))
    lib.eve_profile_reset()
    return v[0].value / iters


VARIANTS = [('generic', dict(tc_row_kernel=0)), ('row', dict(tc_row_kernel=1))]
CASES = [(16, 16), (16, 32), (32, 32), (64, 16), (32, 16), (16, 64)]


def apply(opts, strips=0):
    for k, v in opts.items():
        L.set_option(k, v)
    L.set_option('tc_row_strips', strips)


print('== halo-row kernel accuracy (N=3, 72x128, 3x3): rel err fwd / dgrad vs fp64')
for cin, cout in CASES:
    g = torch.Generator().manual_seed(cin * 100 + cout)
    x = torch.randn(3, cin, 72, 128, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    b = torch.randn(cout, generator=g)
    xd = x.double().requires_grad_(True)
    y = F.conv2d(xd, wt.double(), b.double(), padding=1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy.double())
    for name, opts in VARIANTS:
        for strips in (0, 3):
            if name == 'generic' and strips:
                continue
            apply(opts, strips)
            try:
                got = G.conv_fwd(x.cuda(), wt.cuda(), b.cuda(), 1, 1)
                dx = G.conv_dgrad(dy.cuda(), wt.cuda(), (72, 128), 1, 1)
                torch.cuda.synchronize()
                print('  %2d->%2d %-12s strips=%d  fwd %.2e  dgrad %.2e' %
                      (cin, cout, name, strips, G.rel(got, y), G.rel(dx, xd.grad)), flush=True)
            except Exception as e:  # noqa: BLE001
                print('  %2d->%2d %-12s strips=%d  ERROR %s' % (cin, cout, name, strips, e), flush=True)

print('== halo-row wgrad accuracy (N=3, 72x128, 3x3): rel err dw / db vs fp64')
for cin, cout in [(16, 16), (16, 32), (32, 32), (64, 16), (32, 16), (64, 32)]:
    g = torch.Generator().manual_seed(cin * 100 + cout + 7)
    x = torch.randn(3, cin, 72, 128, generator=g)
    wd = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).double().requires_grad_(True)
    y = F.conv2d(x.double(), wd, None, padding=1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy.double())
    for row in (0, 1):
        for strips in ((0, 3, 72) if row else (0,)):
            L.set_option('tc_row_wgrad', row)
            L.set_option('tc_row_strips', strips)
            try:
                dw, db = G.conv_wgrad(x.cuda(), dy.cuda(), 3, 1, 1)
                torch.cuda.synchronize()
                print('  %2d->%2d row=%d strips=%2d  dw %.2e' % (cin, cout, row, strips, G.rel(dw, wd.grad)),
                      flush=True)
            except Exception as e:  # noqa: BLE001
                print('  %2d->%2d row=%d strips=%d ERROR %s' % (cin, cout, row, strips, e), flush=True)
L.set_option('tc_row_strips', 0)

print('== wgrad kernel time at the bench geometry (N=240, 72x128), ms (kernel + split reduce)')
for cin, cout in [(16, 16), (16, 32), (32, 32), (64, 16)]:
    x = torch.randn(240, 72, 128, cin, device='cuda')
    dy = torch.randn(240, 72, 128, cout, device='cuda')
    dw = torch.empty(cout, cin, 3, 3, device='cuda')
    p = L.ConvParams(240, 72, 128, cin, cout, 3, 1, 1)
    ws = torch.empty(lib.eve_conv2d_workspace_bytes(C.byref(p)), dtype=torch.uint8, device='cuda')

    def runw():
        L.check(lib.eve_conv2d_wgrad(C.byref(p), L.ptr(x), L.ptr(dy), L.ptr(dw), None, L.ptr(ws),
                                     ws.numel(), L.stream_ptr()), 'wgrad')
    for row in (0, 1):
        L.set_option('tc_row_wgrad', row)
        try:
            print('  %2d->%2d row=%d  %.3f ms' % (cin, cout, row, prof_ms(2, runw)), flush=True)
        except Exception as e:  # noqa: BLE001
            print('  %2d->%2d row=%d ERROR %s' % (cin, cout, row, e), flush=True)
    del x, dy, ws

print('== kernel time at the bench geometry (N=240, 72x128), conv kernel only, ms')
for cin, cout in CASES[:4]:
    x = torch.randn(240, 72, 128, cin, device='cuda')
    wt = torch.randn(cout, cin, 3, 3, device='cuda') / (cin * 9) ** 0.5
    y = torch.empty(240, 72, 128, cout, device='cuda')
    p = L.ConvParams(240, 72, 128, cin, cout, 3, 1, 1)
    ws = torch.empty(lib.eve_conv2d_workspace_bytes(C.byref(p)), dtype=torch.uint8, device='cuda')

    def run():
        L.check(lib.eve_conv2d_fwd(C.byref(p), L.ptr(x), L.ptr(wt), None, L.ptr(y), L.ptr(ws),
                                   ws.numel(), L.stream_ptr()), 'fwd')
    for name, opts in VARIANTS:
        apply(opts, 0)
        for cap in (24,):
            L.set_option('tc_stage_cap', cap)
            try:
                print('  %2d->%2d %-12s stage_cap=%2d  %.3f ms' % (cin, cout, name, cap, prof_ms(0, run)),
                      flush=True)
            except Exception as e:  # noqa: BLE001
                print('  %2d->%2d %-12s ERROR %s' % (cin, cout, name, e), flush=True)
    del x, y, ws

print('== generic split-K wgrad vs number of CTA waves, ms (kernel + reduce)')
for (n, cin, h, w, cout, k, st) in [(480, 64, 32, 32, 64, 3, 1), (480, 128, 16, 16, 128, 3, 1),
                                    (480, 256, 8, 8, 256, 3, 1), (480, 512, 4, 4, 512, 3, 1),
                                    (240, 64, 36, 64, 64, 3, 1), (240, 128, 18, 32, 128, 3, 1),
                                    (480, 64, 32, 32, 128, 3, 2), (240, 16, 72, 128, 32, 1, 1)]:
    x = torch.randn(n, h, w, cin, device='cuda')
    oh, ow = h // st, w // st
    dy = torch.randn(n, oh, ow, cout, device='cuda')
    dw = torch.empty(cout, cin, k, k, device='cuda')
    p = L.ConvParams(n, h, w, cin, cout, k, st, k // 2)
    ws = torch.empty(lib.eve_conv2d_workspace_bytes(C.byref(p)), dtype=torch.uint8, device='cuda')

    def runw():
        L.check(lib.eve_conv2d_wgrad(C.byref(p), L.ptr(x), L.ptr(dy), L.ptr(dw), None, L.ptr(ws),
                                     ws.numel(), L.stream_ptr()), 'wgrad')
    line = '  n%d %d->%d %dx%d k%d s%d :' % (n, cin, cout, h, w, k, st)
    for waves in (1, 2, 3, 4, 6):
        L.set_option('tc_wgrad_waves', waves)
        line += '  w%d %.3f' % (waves, prof_ms(2, runw))
    print(line, flush=True)
    del x, dy, ws
L.set_option('tc_wgrad_waves', 3)
