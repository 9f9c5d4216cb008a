"""Single-kernel parity: each C-ABI building block against fp64 CPU arithmetic of the same
op (the torch semantics the reference relies on).  fp32 results are
held to 2e-5 relative (summation-order noise); index outputs are bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from tests import gpu_util as G          # noqa: E402

TOL = 2e-5

# (n, cin, h, w, cout, k, stride, pad) -- every distinct conv geometry family on the path
CONV_CASES = [
    (2, 3, 128, 128, 64, 7, 2, 3),    # EyeNet stem
    (3, 64, 32, 32, 64, 3, 1, 1),     # layer1
    (2, 64, 32, 32, 128, 3, 2, 1),    # layer2.0.conv1 (stride 2)
    (2, 64, 32, 32, 128, 1, 2, 0),    # layer2.0.downsample
    (2, 256, 8, 8, 512, 3, 2, 1),     # layer4.0.conv1
    (5, 512, 4, 4, 512, 3, 1, 1),     # layer4 tail, M = 16 per sample
    (2, 4, 72, 128, 16, 3, 1, 1),     # RefineNet initial.0
    (1, 1, 72, 128, 16, 3, 1, 1),     # ... without screen content
    (2, 16, 72, 128, 32, 1, 1, 0),    # encoder skip 1x1
    (2, 16, 72, 128, 1, 1, 1, 0),     # final.2 (Cout = 1)
    (2, 128, 5, 8, 128, 3, 1, 1),     # ConvGRU gates_1
    (3, 130, 1, 1, 128, 1, 1, 0),     # a Linear (fc_common.0)
    (2, 12, 9, 16, 20, 3, 1, 1),      # ragged channel counts / odd spatial size
]


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv2d_forward_dgrad_wgrad(case):
    n, cin, h, w, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    xd = x.double().requires_grad_(True)
    wd = wt.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    y = F.conv2d(xd, wd, bd, stride=stride, padding=pad)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy.double())

    got = G.conv_fwd(x.cuda(), wt.cuda(), b.cuda(), stride, pad)
    assert got.shape == y.shape
    assert G.rel(got, y) < TOL
    got_nb = G.conv_fwd(x.cuda(), wt.cuda(), None, stride, pad)
    assert G.rel(got_nb, y - bd.view(1, -1, 1, 1)) < TOL
    dx = G.conv_dgrad(dy.cuda(), wt.cuda(), (h, w), stride, pad)
    assert G.rel(dx, xd.grad) < TOL
    dw, db = G.conv_wgrad(x.cuda(), dy.cuda(), k, stride, pad)
    assert G.rel(dw, wd.grad) < TOL
    assert G.rel(db, bd.grad) < TOL


@pytest.mark.parametrize('mode', [0, 1, 2])
@pytest.mark.parametrize('n,h,w', [(3, 128, 128), (2, 64, 96)])
def test_stem_forward_window_modes(mode, n, h, w):
    """EyeNet stem (7x7 stride 2 pad 3, 3 -> 64 channels; eye_net.py:48 via torchvision conv1) with
    the im2col matrix (0), one 32-value window per output column (1) and windows overlapping inside
    the zero-padded image (2): all three against fp64, image borders included."""
    from eve_b200 import lib as L
    L.load()
    saved = L.get_option('stem_windows')
    try:
        L.set_option('stem_windows', mode)
        g = torch.Generator().manual_seed(11 + mode)
        x = torch.randn(n, 3, h, w, generator=g)
        wt = torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5
        b = torch.randn(64, generator=g)
        want = F.conv2d(x.double(), wt.double(), b.double(), stride=2, padding=3)
        got = G.conv_fwd(x.cuda(), wt.cuda(), b.cuda(), 2, 3)
        assert got.shape == want.shape
        assert G.rel(got, want) < TOL
        # the border rows / columns on their own (zero padding of the windows)
        for sl in ((slice(None), slice(None), slice(0, 2)), (slice(None), slice(None), slice(-2, None)),
                   (slice(None), slice(None), slice(None), slice(0, 2)),
                   (slice(None), slice(None), slice(None), slice(-2, None))):
            assert G.rel(got[sl], want[sl]) < TOL
    finally:
        L.set_option('stem_windows', saved)


@pytest.mark.parametrize('ties', [False, True])
@pytest.mark.parametrize('n,c,h,w', [(3, 64, 64, 64), (2, 8, 6, 10)])
def test_stem_norm_relu_maxpool_backward(ties, n, c, h, w):
    """Backward of maxpool3x3s2p1(relu(InstanceNorm(x))) (torchvision ResNet bn1 / relu / maxpool
    behind eye_net.py:48-50) in the gather form (eve_in_relu_maxpool_bwd: the norm's sums taken over
    the pool windows, no dense un-pooled gradient) against fp64 autograd; `ties` draws x from three
    values, so that most windows hold their maximum several times (the first one takes the gradient,
    as in ATen).  Both output forms: fp32 and bf16 hi + lo planes."""
    from eve_b200 import lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(5 + ties)
    x = (torch.randint(0, 3, (n, c, h, w), generator=g).float() if ties
         else torch.randn(n, c, h, w, generator=g) * 2.0 + 1.0)
    oh, ow = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    dy = torch.randn(n, c, oh, ow, generator=g)
    xh, dyh = G.nhwc(x.cuda()), G.nhwc(dy.cuda())
    y = torch.empty((n, oh, ow, c), device='cuda')
    idx = torch.empty((n, oh, ow, c), dtype=torch.int32, device='cuda')
    mean, rstd = torch.empty((n, c), device='cuda'), torch.empty((n, c), device='cuda')
    L.check(lib.eve_in_relu_maxpool_fwd(L.ptr(xh), n, h, w, c, L.ptr(mean), L.ptr(rstd), L.ptr(y),
                                        L.ptr(idx), L.stream_ptr()), 'in_relu_maxpool_fwd')
    dx = torch.empty_like(xh)
    hi = torch.empty(xh.shape, dtype=torch.bfloat16, device='cuda')
    lo = torch.empty_like(hi)
    scratch = torch.empty(2 * n * c, device='cuda')
    L.check(lib.eve_in_relu_maxpool_bwd(L.ptr(dyh), L.ptr(y), L.ptr(idx), L.ptr(xh), n, h, w, c,
                                        L.ptr(mean), L.ptr(rstd), L.ptr(dx), L.ptr(hi), L.ptr(lo),
                                        L.ptr(scratch), L.stream_ptr()), 'in_relu_maxpool_bwd')
    torch.cuda.synchronize()
    xd = x.double().requires_grad_(True)
    pooled = F.max_pool2d(F.relu(F.instance_norm(xd, eps=1e-5)), 3, 2, 1)
    pooled.backward(dy.double())
    assert G.rel(G.nchw(y), pooled.detach()) < 1e-5
    assert G.rel(G.nchw(dx), xd.grad) < 2e-5
    planes = hi.float() + lo.float()
    assert float((planes - dx).abs().max()) <= 2.0 ** -15 * float(dx.abs().max())


def test_conv2d_empty_batch():
    wt = torch.randn(8, 4, 3, 3).cuda()
    y = G.conv_fwd(torch.zeros(0, 4, 8, 8).cuda(), wt, None, 1, 1)
    assert y.shape == (0, 8, 8, 8)
    dw, db = G.conv_wgrad(torch.zeros(0, 4, 8, 8).cuda(), torch.zeros(0, 8, 8, 8).cuda(), 3, 1, 1)
    assert float(dw.abs().max()) == 0.0 and float(db.abs().max()) == 0.0


@pytest.mark.parametrize('shape,affine,act', [((3, 64, 64, 64), False, 1), ((2, 16, 72, 128), True, 1),
                                              ((2, 256, 5, 8), True, 2), ((4, 512, 4, 4), False, 0),
                                              ((2, 32, 9, 16), True, 2)])
def test_instance_norm_act(shape, affine, act):
    g = torch.Generator().manual_seed(7)
    n, c, h, w = shape
    x = torch.randn(shape, generator=g) * 3.0 + 5.0      # |mean| >> std stresses the variance
    gamma = (1.0 + 0.1 * torch.randn(c, generator=g)) if affine else None
    beta = 0.1 * torch.randn(c, generator=g) if affine else None
    xd = x.double().requires_grad_(True)
    gd = gamma.double().requires_grad_(True) if affine else None
    bd = beta.double().requires_grad_(True) if affine else None
    y = F.instance_norm(xd, weight=gd, bias=bd, eps=1e-5)
    y = {0: lambda t: t, 1: F.relu, 2: lambda t: F.leaky_relu(t, 0.01)}[act](y)
    dy = torch.randn(shape, generator=g)
    y.backward(dy.double())
    cg = gamma.cuda() if affine else None
    cb = beta.cuda() if affine else None
    got, mean, rstd = G.instnorm_fwd(x.cuda(), cg, cb, act)
    assert G.rel(got, y) < TOL
    assert G.rel(mean, x.double().mean(dim=(2, 3))) < 1e-6
    dx, dgamma, dbeta = G.instnorm_bwd(dy.cuda(), got, x.cuda(), mean, rstd, cg, act)
    assert G.rel(dx, xd.grad) < 5e-5
    if affine:
        assert G.rel(dgamma, gd.grad) < 5e-5
        assert G.rel(dbeta, bd.grad) < 5e-5


@pytest.mark.parametrize('h,w,oh,ow', [(72, 128, 36, 64), (18, 32, 9, 16), (9, 16, 5, 8), (7, 5, 3, 4)])
def test_adaptive_maxpool_values_indices_and_backward(h, w, oh, ow):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 8, h, w, generator=g)
    # ties: quantise so that equal maxima occur and the first-max rule is exercised
    x = torch.round(x * 2.0) / 2.0
    xd = x.clone().requires_grad_(True)
    y, idx = F.adaptive_max_pool2d(xd, (oh, ow), return_indices=True)
    got, gidx = G.adaptive_maxpool(x.cuda(), oh, ow)
    assert torch.equal(got.cpu(), y.detach())
    assert torch.equal(gidx.cpu().long(), idx)           # bit-exact index output
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    dx = G.adaptive_maxpool_bwd(dy.cuda(), gidx, h, w)
    assert G.rel(dx, xd.grad) < 1e-6


@pytest.mark.parametrize('h,w,oh,ow', [(5, 8, 9, 16), (9, 16, 18, 32), (36, 64, 72, 128), (3, 4, 7, 5),
                                          (1, 1, 2, 2), (1, 4, 2, 8), (2, 1, 4, 2)])
def test_bilinear_upsample(h, w, oh, ow):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 12, h, w, generator=g)
    xd = x.double().requires_grad_(True)
    y = F.interpolate(xd, size=(oh, ow), mode='bilinear', align_corners=False)
    got = G.upsample(x.cuda(), oh, ow)
    assert G.rel(got, y) < 1e-6
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy.double())
    dx = G.upsample_bwd(dy.cuda(), h, w)
    assert G.rel(dx, xd.grad) < 1e-5


def test_layout_round_trip():
    import ctypes as C
    from eve_b200 import lib as L
    lib = L.load()
    x = torch.randn(3, 5, 7, 9).cuda()
    y = torch.empty(3, 7, 9, 5, device='cuda')
    z = torch.empty_like(x)
    L.check(lib.eve_nchw_to_nhwc(L.ptr(x), 3, 5, 7, 9, L.ptr(y), L.stream_ptr()), 'a')
    L.check(lib.eve_nhwc_to_nchw(L.ptr(y), 3, 5, 7, 9, L.ptr(z), L.stream_ptr()), 'b')
    assert torch.equal(y, x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(z, x)
    # three-channel images take the plane-interleaving kernel
    x = torch.randn(5, 3, 12, 20).cuda()
    y = torch.empty(5, 12, 20, 3, device='cuda')
    L.check(lib.eve_nchw_to_nhwc(L.ptr(x), 5, 3, 12, 20, L.ptr(y), L.stream_ptr()), 'c')
    assert torch.equal(y, x.permute(0, 2, 3, 1).contiguous())


def test_adam_clip_matches_torch():
    from eve_b200 import ops
    g = torch.Generator().manual_seed(5)
    n = 100003
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone().double())
    opt = torch.optim.Adam([ref], lr=1e-2, weight_decay=5e-3)
    p = p0.clone().cuda()
    m = torch.zeros(n, device='cuda')
    v = torch.zeros(n, device='cuda')
    for step in range(1, 4):
        grad = torch.randn(n, generator=g) * (30.0 if step == 2 else 0.001)
        ref.grad = grad.double().clone()
        want_norm = torch.nn.utils.clip_grad_norm_([ref], 5.0)
        opt.step()
        # the flat buffer holds a 2-rank sum: grad_scale = 1/2 undoes it
        norm = ops.adam_clip_step(p, (2.0 * grad).cuda(), m, v, step, lr=1e-2,
                                  weight_decay=5e-3, max_norm=5.0, grad_scale=0.5)
        assert abs(float(norm) - float(want_norm)) < 1e-4 * float(want_norm)
        assert G.rel(p, ref.data) < 1e-5
