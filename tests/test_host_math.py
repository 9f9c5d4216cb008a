"""The device math of the geometry / loss kernels, checked on the CPU.

eve_b200/csrc/gaze_math.cuh holds the value and the hand-derived vector-Jacobian products of the
gaze geometry (common.py:32-218) and of the angular loss (losses/angular.py) as scalar-templated
host/device functions.  Here the header is compiled for the host in double precision (g++) and
every function and VJP is compared with torch autograd over the product's torch formulas, which
tests/test_host_logic.py pins to the unmodified reference functions."""
import ctypes as C
import math
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def hm():
    src = os.path.join(REPO, 'tests', 'host_math', 'gaze_math_host.cpp')
    out = os.path.join(tempfile.mkdtemp(prefix='eve_hm_'), 'libgm.so')
    subprocess.check_call(['g++', '-O1', '-shared', '-fPIC', '-o', out, src])
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _rot(n, g):
    from eve_b200.models.common import pitchyaw_to_rotation
    return pitchyaw_to_rotation(torch.rand(n, 2, generator=g, dtype=torch.float64) - 0.5)


def test_combined_gaze_and_its_vjp(hm):
    from eve_b200.models.common import _combined_gaze_torch as calculate_combined_gaze_direction
    g = torch.Generator().manual_seed(0)
    n = 64
    o = torch.randn(n, 3, generator=g, dtype=torch.float64) * 5 + torch.tensor([0.0, 0.0, 600.0])
    pog = (torch.rand(n, 2, generator=g, dtype=torch.float64) * 500).requires_grad_(True)
    R = _rot(n, g)
    cam = torch.eye(4, dtype=torch.float64).repeat(n, 1, 1)
    cam[:, :3, :3] = _rot(n, g)
    cam[:, :3, 3] = torch.randn(n, 3, generator=g, dtype=torch.float64) * 20
    want = calculate_combined_gaze_direction(o, pog, R, cam)
    gg = torch.randn(n, 2, generator=g, dtype=torch.float64)
    want.backward(gg)
    got, dpog = np.zeros((n, 2)), np.zeros((n, 2))
    args = [np.ascontiguousarray(t.detach().numpy()) for t in (o, pog, R, cam)]
    hm.hm_combined_gaze(n, *map(_p, args), _p(got))
    hm.hm_combined_gaze_vjp(n, *map(_p, args), _p(np.ascontiguousarray(gg.numpy())), _p(dpog))
    assert np.abs(got - want.detach().numpy()).max() < 1e-12
    assert np.abs(dpog - pog.grad.numpy()).max() < 1e-12 * max(1.0, np.abs(dpog).max())


@pytest.mark.parametrize('inverse', [False, True])
def test_offset_augmentation_and_its_vjp(hm, inverse):
    from eve_b200.models.common import _offset_augmentation_torch as apply_offset_augmentation
    g = torch.Generator().manual_seed(1)
    n = 64
    gz = ((torch.rand(n, 2, generator=g, dtype=torch.float64) - 0.5) * 0.8).requires_grad_(True)
    R = _rot(n, g)
    kappa = torch.randn(n, 2, generator=g, dtype=torch.float64) * math.radians(3.0)
    want = apply_offset_augmentation(gz, R, kappa, inverse_kappa=inverse)
    gout = torch.randn(n, 2, generator=g, dtype=torch.float64)
    want.backward(gout)
    got, dg = np.zeros((n, 2)), np.zeros((n, 2))
    args = [np.ascontiguousarray(t.detach().numpy()) for t in (gz, R, kappa)]
    hm.hm_offset_aug(n, *map(_p, args), int(inverse), _p(got))
    hm.hm_offset_aug_vjp(n, *map(_p, args), int(inverse), _p(np.ascontiguousarray(gout.numpy())), _p(dg))
    assert np.abs(got - want.detach().numpy()).max() < 1e-12
    assert np.abs(dg - gz.grad.numpy()).max() < 1e-11


def test_angular_error_and_its_vjp(hm):
    from eve_b200.losses import angular_loss
    g = torch.Generator().manual_seed(2)
    n = 64
    a = ((torch.rand(n, 2, generator=g, dtype=torch.float64) - 0.5)).requires_grad_(True)
    b = (torch.rand(n, 2, generator=g, dtype=torch.float64) - 0.5)
    want = angular_loss.per_frame(a, b)
    gl = torch.randn(n, generator=g, dtype=torch.float64)
    want.backward(gl)
    got, da = np.zeros(n), np.zeros((n, 2))
    an, bn = np.ascontiguousarray(a.detach().numpy()), np.ascontiguousarray(b.numpy())
    hm.hm_angular.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
    hm.hm_angular_vjp.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                  C.c_void_p, C.c_void_p]
    hm.hm_angular(n, _p(an), _p(bn), -1.0 + 1e-8, 1.0 - 1e-8, _p(got))
    hm.hm_angular_vjp(n, _p(an), _p(bn), -1.0 + 1e-8, 1.0 - 1e-8,
                      _p(np.ascontiguousarray(gl.numpy())), _p(da))
    assert np.abs(got - want.detach().numpy()).max() < 1e-9
    assert np.abs(da - a.grad.numpy()).max() < 1e-8 * max(1.0, np.abs(da).max())


@pytest.mark.parametrize('ties', [False, True])
def test_stem_backward_window_sum_identity(ties):
    """The algebra behind stem_pool_in_backward (eve_b200/csrc/norm.cu): behind
    maxpool3x3s2p1(relu(InstanceNorm(x))) (torchvision bn1 / relu / maxpool, eye_net.py:48-50) the
    two per-(n, c) sums of the norm's backward, sum g and sum g * xhat over the PIXELS with g the
    un-pooled, ReLU-masked gradient, equal sums over the POOL WINDOWS of the pooled tensors alone:
    sum dp [p > 0] and sum dp * p.  Checked in fp64 against torch autograd, with and without ties
    (the first maximum of a window takes the gradient, a zero maximum passes none)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(21 + ties)
    n, c, h, w = 2, 5, 12, 10
    x = (torch.randint(0, 3, (n, c, h, w), generator=g).double() if ties
         else torch.randn(n, c, h, w, generator=g, dtype=torch.float64))
    mean = x.mean(dim=(2, 3), keepdim=True)
    rstd = (x.var(dim=(2, 3), unbiased=False, keepdim=True) + 1e-5).rsqrt()
    xhat = ((x - mean) * rstd).requires_grad_(True)
    p = F.max_pool2d(F.relu(xhat), 3, 2, 1)
    dp = torch.randn(p.shape, generator=g, dtype=torch.float64)
    p.backward(dp)
    gpix = xhat.grad                                   # dL/dxhat = un-pooled, masked gradient
    sum_g = gpix.sum(dim=(2, 3))
    sum_gx = (gpix * xhat.detach()).sum(dim=(2, 3))
    pd = p.detach()
    win_g = (dp * (pd > 0)).sum(dim=(2, 3))
    win_gx = (dp * pd).sum(dim=(2, 3))
    assert float((sum_g - win_g).abs().max()) < 1e-12 * max(1.0, float(sum_g.abs().max()))
    assert float((sum_gx - win_gx).abs().max()) < 1e-12 * max(1.0, float(sum_gx.abs().max()))
    # and the norm's input gradient built from them is what autograd gives for the whole chain
    xd = x.clone().requires_grad_(True)
    F.max_pool2d(F.relu(F.instance_norm(xd, eps=1e-5)), 3, 2, 1).backward(dp)
    hw = h * w
    dx = rstd * (gpix - win_g[..., None, None] / hw - xhat.detach() * win_gx[..., None, None] / hw)
    assert float((dx - xd.grad).abs().max()) < 1e-10 * max(1.0, float(xd.grad.abs().max()))


@pytest.mark.parametrize('h,w', [(1, 1), (1, 4), (2, 1), (5, 7), (9, 16)])
def test_exact_2x_bilinear_weights(h, w):
    """The constant weights of upsample2x_{fwd,bwd}_kernel (eve_b200/csrc/pool.cu) for
    nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) (refine_net.py:101,124):
    output 2k reads inputs (k-1, k) with (0.25, 0.75) -- (k, .) with weight 1 at k = 0 --, output
    2k+1 reads (k, min(k+1, last)) with (0.75, 0.25); hence input k receives 0.25 / 0.75 / 0.75 /
    0.25 from outputs 2k-1 .. 2k+2, with 1 instead of 0.75 where the source is clamped.  Restated
    in numpy and compared with torch (forward) and its autograd (backward) in fp64."""
    import torch.nn.functional as F

    def axis_matrix(n):
        m = np.zeros((2 * n, n))
        for k in range(n):
            if k == 0:
                m[0, 0] = 1.0
            else:
                m[2 * k, k - 1], m[2 * k, k] = 0.25, 0.75
            m[2 * k + 1, k] += 0.75
            m[2 * k + 1, min(k + 1, n - 1)] += 0.25
        return m

    def gather_weights(n):
        wts = np.zeros((n, 4))            # outputs 2k-1 .. 2k+2 -> input k
        for k in range(n):
            wts[k] = [0.25 if k > 0 else 0.0, 1.0 if k == 0 else 0.75,
                      1.0 if k == n - 1 else 0.75, 0.25 if k < n - 1 else 0.0]
        return wts

    rng = np.random.RandomState(h * 31 + w)
    x = rng.randn(2, 3, h, w)
    my, mx = axis_matrix(h), axis_matrix(w)
    xt = torch.from_numpy(x).requires_grad_(True)
    y = F.interpolate(xt, scale_factor=2, mode='bilinear', align_corners=False)
    mine = np.einsum('ah,nchw,bw->ncab', my, x, mx)
    assert np.abs(mine - y.detach().numpy()).max() < 1e-13
    dy = rng.randn(*y.shape)
    y.backward(torch.from_numpy(dy))
    gy, gx = gather_weights(h), gather_weights(w)
    dx = np.zeros_like(x)
    for k in range(h):
        for j in range(w):
            for a in range(4):
                for b in range(4):
                    oy, ox = 2 * k - 1 + a, 2 * j - 1 + b
                    if gy[k, a] and gx[j, b]:
                        dx[:, :, k, j] += gy[k, a] * gx[j, b] * dy[:, :, oy, ox]
    assert np.abs(dx - xt.grad.numpy()).max() < 1e-12


@pytest.mark.parametrize('rows_per_strip', [1, 2, 3, 7])
def test_row_stacked_weight_gradient_decomposition(rows_per_strip):
    """The work decomposition of conv_tc_wgrad_row3_kernel (eve_b200/csrc/conv_tc.cu) restated in
    numpy: an image is cut into strips of rows; ring entry e of a strip holds the dy row h0-1+e
    (one zero pixel on each side, rows outside the image zero) and the x row of the same index --
    ZERO for the two halo entries, whose x rows belong to the neighbouring strips.  For every row h
    of the strip ONE product of the dy row h with the x rows (h-1, h, h+1) feeds the accumulators
    (D_0, D_1, D_2) of the three filter rows; the strip's dy halo rows meet their single x row
    separately (row h0-1 with x row h0 into D_2, row h1 with x row h1-1 into D_0).  M block b of a
    product is the dy row shifted by b pixels, i.e. filter column q = 2 - b.  Summed over strips and
    images this must be the weight gradient of a 3x3 stride-1 'same' convolution (fp64 autograd)."""
    import torch.nn.functional as F
    rng = np.random.RandomState(rows_per_strip)
    n, cin, cout, H, W = 2, 2, 3, 7, 6
    x = rng.randn(n, cin, H, W)
    dy = rng.randn(n, cout, H, W)
    xt = torch.from_numpy(x)
    wt = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(xt, wt, padding=1).backward(torch.from_numpy(dy))
    want = wt.grad.numpy()

    def product(dyrow, xrow):
        # dyrow [W + 2][cout] staged from pixel -1, xrow [W][cin] -> D[b][co][ci]
        return np.stack([dyrow[b:b + W].T @ xrow for b in range(3)])

    D = np.zeros((3, 3, cout, cin))                       # [filter row r][M block b][co][ci]
    for img in range(n):
        for h0 in range(0, H, rows_per_strip):
            h1 = min(H, h0 + rows_per_strip)
            ent_dy, ent_x = [], []
            for hr in range(h0 - 1, h1 + 1):
                row = np.zeros((W + 2, cout))
                if 0 <= hr < H:
                    row[1:W + 1] = dy[img, :, hr, :].T
                ent_dy.append(row)
                ent_x.append(x[img, :, hr, :].T if h0 <= hr < h1 else np.zeros((W, cin)))
            for h in range(h0, h1):
                e = h - h0 + 1                              # centre entry
                for r in range(3):                          # one N = 3 Cin instruction on the device
                    D[r] += product(ent_dy[e], ent_x[e - 1 + r])
                if h == h0:
                    D[2] += product(ent_dy[e - 1], ent_x[e])
                if h == h1 - 1:
                    D[0] += product(ent_dy[e + 1], ent_x[e])
    got = np.zeros_like(want)
    for r in range(3):
        for q in range(3):
            got[:, :, r, q] = D[r][2 - q]
    assert np.abs(got - want).max() < 1e-12 * max(1.0, np.abs(want).max())
