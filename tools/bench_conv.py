import sys, torch, ctypes as C
sys.path.insert(0, '.')
from eve_b200 import lib as L
lib = L.load()
lib.eve_set_conv_mode(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
shapes = [  # n, cin, h, w, cout, k, stride
    (480, 64, 32, 32, 64, 3, 1), (480, 128, 16, 16, 128, 3, 1), (480, 256, 8, 8, 256, 3, 1),
    (480, 512, 4, 4, 512, 3, 1), (240, 64, 36, 64, 64, 3, 1), (240, 128, 18, 32, 128, 3, 1),
    (240, 256, 9, 16, 256, 3, 1), (240, 16, 72, 128, 32, 3, 1), (240, 64, 72, 128, 16, 3, 1),
    (240, 512, 9, 16, 128, 3, 1), (480, 64, 32, 32, 128, 3, 2)]
only = int(sys.argv[2]) if len(sys.argv) > 2 else -1
for si, (n, cin, h, w, cout, k, st) in enumerate(shapes):
    if only >= 0 and si != only: continue
    pad = k // 2
    oh, ow = (h + 2 * pad - k) // st + 1, (w + 2 * pad - k) // st + 1
    p = L.ConvParams(n, h, w, cin, cout, k, st, pad)
    x = torch.randn(n, h, w, cin, device='cuda'); wt = torch.randn(cout, cin, k, k, device='cuda') * 0.05
    y = torch.empty(n, oh, ow, cout, device='cuda'); dy = torch.randn_like(y); dx = torch.empty_like(x)
    dw = torch.empty_like(wt)
    ws = torch.empty(lib.eve_conv2d_workspace_bytes(C.byref(p)), dtype=torch.uint8, device='cuda')
    fl = 2.0 * n * oh * ow * cout * cin * k * k
    def t(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    s = L.stream_ptr()
    tf = t(lambda: L.check(lib.eve_conv2d_fwd(C.byref(p), L.ptr(x), L.ptr(wt), None, L.ptr(y), L.ptr(ws), ws.numel(), s), 'f'))
    td = t(lambda: L.check(lib.eve_conv2d_dgrad(C.byref(p), L.ptr(dy), L.ptr(wt), L.ptr(dx), L.ptr(ws), ws.numel(), s), 'd'))
    tw = t(lambda: L.check(lib.eve_conv2d_wgrad(C.byref(p), L.ptr(x), L.ptr(dy), L.ptr(dw), None, L.ptr(ws), ws.numel(), s), 'w'))
    print('%-36s GF %7.1f | fwd %.3f ms %6.1f TF | dgrad %.3f ms %6.1f TF | wgrad %.3f ms %6.1f TF (incl. split/prep kernels)' % (
        str((n, cin, h, w, cout, k, st)), fl / 1e9, tf, fl / tf / 1e9, td, fl / td / 1e9, tw, fl / tw / 1e9))
