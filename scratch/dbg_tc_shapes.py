import sys, torch
sys.path.insert(0, '.')
import torch.nn.functional as F
from eve_b200 import lib as L
from tests import gpu_util as G
lib = L.load()
lib.eve_set_conv_mode(1)
def run(n, cin, h, w, cout, k, scale=1.0, seed=0):
    pad = k // 2
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    xd = x.double().requires_grad_(True); wd = wt.double().requires_grad_(True)
    y = F.conv2d(xd, wd, None, padding=pad)
    dy = torch.randn(y.shape, generator=g) * scale
    y.backward(dy.double())
    got = G.conv_fwd(x.cuda(), wt.cuda(), None, 1, pad)
    dx = G.conv_dgrad(dy.cuda(), wt.cuda(), (h, w), 1, pad)
    dw, _ = G.conv_wgrad(x.cuda(), dy.cuda(), k, 1, pad, with_bias=False)
    print((n, cin, h, w, cout, k, scale), 'fwd %.1e dgrad %.1e wgrad %.1e' % (G.rel(got, y), G.rel(dx, xd.grad), G.rel(dw, wd.grad)))
for n in (1, 2, 3, 5, 8):
    run(n, 512, 4, 4, 512, 3)
for n in (1, 3):
    run(n, 256, 8, 8, 256, 3)
    run(n, 128, 16, 16, 128, 3)
    run(n, 64, 32, 32, 64, 3)
run(3, 512, 4, 4, 512, 3, scale=1e-4)
run(3, 512, 4, 4, 512, 3, scale=1e4)
