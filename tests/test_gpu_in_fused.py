"""One-pass InstanceNorm cluster kernels (csrc/in_fused.cu) against fp64 torch arithmetic.

Reference semantics: nn.InstanceNorm2d(eps=1e-5, biased variance) + ReLU / LeakyReLU(0.01) as
used at eye_net.py:48-50 (non-affine, residual added before the ReLU inside torchvision's
BasicBlock) and refine_net.py:45-62 (affine, two affine sets over the same statistics when a
block has a skip convolution).  Every output of the kernels is checked: fp32 result, operand
planes (hi + lo), statistics, dx (fp32 and bf16 planes), g_out, affine and bias gradients.
Shapes cover every (H*W, C) decomposition the networks hit: cluster sizes 1..8, channel
groups 8..256, the two-tensor staging of the EyeNet block end.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from eve_b200 import lib as L            # noqa: E402
from tests import gpu_util as G          # noqa: E402

ACT = {0: lambda t: t, 1: F.relu, 2: lambda t: F.leaky_relu(t, 0.01)}


def _planes(hi, lo, fmt):
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    return hi.view(dt).double() + lo.view(dt).double()


def _u16(shape):
    return torch.empty(shape, dtype=torch.int16, device='cuda')


# (n, c, h, w): RefineNet levels 0..4 and EyeNet layers 1..4 channel/size pairs (small n)
SHAPES = [(2, 16, 72, 128), (1, 64, 72, 128), (2, 32, 36, 64), (2, 128, 36, 64), (2, 64, 18, 32),
          (2, 256, 18, 32), (2, 128, 9, 16), (2, 512, 9, 16), (3, 256, 5, 8), (3, 64, 5, 8),
          (3, 64, 32, 32), (2, 128, 16, 16), (2, 256, 8, 8), (5, 512, 4, 4), (2, 20, 7, 5)]


@pytest.mark.parametrize('shape', SHAPES, ids=lambda s: 'x'.join(map(str, s)))
@pytest.mark.parametrize('act', [1, 2])
def test_forward_two_affine_sets_and_planes(shape, act):
    lib = L.load()
    n, c, h, w = shape
    g = torch.Generator().manual_seed(11 + c + h)
    x = torch.randn(shape, generator=g) * 2.0 + 3.0
    ga, gb = 1.0 + 0.2 * torch.randn(c, generator=g), 1.0 + 0.2 * torch.randn(c, generator=g)
    ba, bb = 0.3 * torch.randn(c, generator=g), 0.3 * torch.randn(c, generator=g)
    xd = x.double()
    want_a = ACT[act](F.instance_norm(xd, weight=ga.double(), bias=ba.double(), eps=1e-5))
    want_b = ACT[act](F.instance_norm(xd, weight=gb.double(), bias=bb.double(), eps=1e-5))
    xh = G.nhwc(x.cuda())
    y = torch.empty_like(xh)
    mean, rstd = torch.empty((n, c), device='cuda'), torch.empty((n, c), device='cuda')
    cga, cba, cgb, cbb = ga.cuda(), ba.cuda(), gb.cuda(), bb.cuda()   # keep the device copies alive
    for fmt, tol in ((0, 2e-6), (1, 2e-4)):
        hi_a, lo_a, hi_b, lo_b = (_u16(xh.shape) for _ in range(4))
        L.check(lib.eve_instnorm_fused_fwd(
            L.ptr(xh), None, 0, n, h * w, c, L.ptr(cga), L.ptr(cba), L.ptr(cgb),
            L.ptr(cbb), act, fmt, L.ptr(mean), L.ptr(rstd), None, None, L.ptr(y),
            L.ptr(hi_a), L.ptr(lo_a), L.ptr(hi_b), L.ptr(lo_b), L.stream_ptr()), 'fused_fwd')
        torch.cuda.synchronize()
        assert G.rel(G.nchw(y), want_a) < 2e-6
        assert G.rel(mean, xd.mean(dim=(2, 3))) < 1e-6
        assert G.rel(rstd, 1.0 / torch.sqrt(xd.var(dim=(2, 3), unbiased=False) + 1e-5)) < 2e-6
        assert G.rel(G.nchw(_planes(hi_a, lo_a, fmt)), want_a) < tol
        assert G.rel(G.nchw(_planes(hi_b, lo_b, fmt)), want_b) < tol


@pytest.mark.parametrize('shape', [(3, 64, 32, 32), (2, 128, 16, 16), (2, 256, 8, 8), (5, 512, 4, 4)],
                         ids=lambda s: 'x'.join(map(str, s)))
@pytest.mark.parametrize('mode', [1, 2])
def test_forward_block_end_with_residual(shape, mode):
    """torchvision BasicBlock tail: relu(IN(b) + identity) / relu(IN(b) + IN(downsample))."""
    lib = L.load()
    n, c, h, w = shape
    g = torch.Generator().manual_seed(5 + c)
    b = torch.randn(shape, generator=g) * 1.5 - 0.5
    r = torch.randn(shape, generator=g) * 0.7 + 0.2
    want = F.instance_norm(b.double(), eps=1e-5)
    want = F.relu(want + (r.double() if mode == 1 else F.instance_norm(r.double(), eps=1e-5)))
    bh, rh = G.nhwc(b.cuda()), G.nhwc(r.cuda())
    y = torch.empty_like(bh)
    st = [torch.empty((n, c), device='cuda') for _ in range(4)]
    hi, lo = _u16(bh.shape), _u16(bh.shape)
    L.check(lib.eve_instnorm_fused_fwd(
        L.ptr(bh), L.ptr(rh), mode, n, h * w, c, None, None, None, None, 1, 0, L.ptr(st[0]),
        L.ptr(st[1]), L.ptr(st[2]), L.ptr(st[3]), L.ptr(y), L.ptr(hi), L.ptr(lo), None, None,
        L.stream_ptr()), 'fused_fwd')
    torch.cuda.synchronize()
    assert G.rel(G.nchw(y), want) < 2e-6
    assert G.rel(G.nchw(_planes(hi, lo, 0)), want) < 2e-6
    if mode == 2:
        assert G.rel(st[2], r.double().mean(dim=(2, 3))) < 1e-6


@pytest.fixture()
def in_stream():
    """Sets the `in_stream` switch (0 staged in shared memory, 1 automatic, 2 streaming: second read
    of dy / x from L2) for one test and restores it."""
    L.load()
    saved = L.get_option('in_stream')
    yield lambda v: L.set_option('in_stream', v)
    L.set_option('in_stream', saved)


@pytest.mark.parametrize('shape', SHAPES, ids=lambda s: 'x'.join(map(str, s)))
@pytest.mark.parametrize('dual', [False, True])
@pytest.mark.parametrize('stream', [0, 1, 2])
def test_backward_all_outputs(shape, dual, stream, in_stream):
    """dx (+ addend) in fp32 and as bf16 planes, affine gradients of both sets, bias gradient;
    the staged and the streaming kernel both against fp64."""
    lib = L.load()
    in_stream(stream)
    n, c, h, w = shape
    act = 2
    g = torch.Generator().manual_seed(23 + c + h)
    x = torch.randn(shape, generator=g) * 2.0 + 1.0
    ga, gb = 1.0 + 0.2 * torch.randn(c, generator=g), 1.0 + 0.2 * torch.randn(c, generator=g)
    ba, bb = 0.3 * torch.randn(c, generator=g), 0.3 * torch.randn(c, generator=g)
    dy, dy2 = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    add = torch.randn(shape, generator=g)
    xd = x.double().requires_grad_(True)
    gad, bad = ga.double().requires_grad_(True), ba.double().requires_grad_(True)
    gbd, bbd = gb.double().requires_grad_(True), bb.double().requires_grad_(True)
    ya = ACT[act](F.instance_norm(xd, weight=gad, bias=bad, eps=1e-5))
    loss = (ya * dy.double()).sum()
    if dual:
        yb = ACT[act](F.instance_norm(xd, weight=gbd, bias=bbd, eps=1e-5))
        loss = loss + (yb * dy2.double()).sum()
    loss.backward()
    want_dx = xd.grad + add.double()
    xh = G.nhwc(x.cuda())
    mean = x.double().mean(dim=(2, 3)).float().cuda()
    rstd = (1.0 / torch.sqrt(x.double().var(dim=(2, 3), unbiased=False) + 1e-5)).float().cuda()
    dx = torch.empty_like(xh)
    hi, lo = _u16(xh.shape), _u16(xh.shape)
    dg, db, dg2, db2, dbias = (torch.empty(c, device='cuda') for _ in range(5))
    ws = torch.empty(lib.eve_instnorm_fused_workspace_bytes(n, h * w, c), dtype=torch.uint8,
                     device='cuda')
    cdy, cdy2, cadd = G.nhwc(dy.cuda()), G.nhwc(dy2.cuda()), G.nhwc(add.cuda())
    cga, cba, cgb, cbb = ga.cuda(), ba.cuda(), gb.cuda(), bb.cuda()   # keep the device copies alive
    L.check(lib.eve_instnorm_fused_bwd(
        L.ptr(cdy), L.ptr(cdy2) if dual else None, None, L.ptr(xh), n,
        h * w, c, L.ptr(mean), L.ptr(rstd), L.ptr(cga), L.ptr(cba),
        L.ptr(cgb) if dual else None, L.ptr(cbb) if dual else None, act,
        L.ptr(cadd), L.ptr(dx), L.ptr(hi), L.ptr(lo), None, L.ptr(dg), L.ptr(db),
        L.ptr(dg2) if dual else None, L.ptr(db2) if dual else None, L.ptr(dbias), L.ptr(ws),
        ws.numel(), L.stream_ptr()), 'fused_bwd')
    torch.cuda.synchronize()
    assert G.rel(G.nchw(dx), want_dx) < 2e-5
    assert G.rel(G.nchw(_planes(hi, lo, 1)), want_dx) < 5e-5
    assert G.rel(dg, gad.grad) < 2e-5 and G.rel(db, bad.grad) < 2e-5
    if dual:
        assert G.rel(dg2, gbd.grad) < 2e-5 and G.rel(db2, bbd.grad) < 2e-5
    assert G.rel(dbias, want_dx.sum(dim=(0, 2, 3))) < 2e-5


@pytest.mark.parametrize('shape', [(3, 64, 32, 32), (5, 512, 4, 4), (1, 64, 72, 128)],
                         ids=lambda s: 'x'.join(map(str, s)))
@pytest.mark.parametrize('stream', [0, 2])
def test_backward_block_end_mask_from_saved_output(shape, stream, in_stream):
    """out = relu(IN(b) + skip): act' from the saved output, g_out = dy * relu'(out) (the
    gradient of the skip branch), dx through the non-affine norm as bf16 planes only."""
    lib = L.load()
    in_stream(stream)
    n, c, h, w = shape
    g = torch.Generator().manual_seed(31 + c)
    b = torch.randn(shape, generator=g)
    skip = torch.randn(shape, generator=g)
    dy = torch.randn(shape, generator=g)
    bd = b.double().requires_grad_(True)
    sd = skip.double().requires_grad_(True)
    out = F.relu(F.instance_norm(bd, eps=1e-5) + sd)
    out.backward(dy.double())
    bh = G.nhwc(b.cuda())
    mean = b.double().mean(dim=(2, 3)).float().cuda()
    rstd = (1.0 / torch.sqrt(b.double().var(dim=(2, 3), unbiased=False) + 1e-5)).float().cuda()
    hi, lo = _u16(bh.shape), _u16(bh.shape)
    gout = torch.empty_like(bh)
    ws = torch.empty(lib.eve_instnorm_fused_workspace_bytes(n, h * w, c), dtype=torch.uint8,
                     device='cuda')
    cdy, cout = G.nhwc(dy.cuda()), G.nhwc(out.detach().float().cuda())
    L.check(lib.eve_instnorm_fused_bwd(
        L.ptr(cdy), None, L.ptr(cout), L.ptr(bh), n,
        h * w, c, L.ptr(mean), L.ptr(rstd), None, None, None, None, 1, None, None, L.ptr(hi),
        L.ptr(lo), L.ptr(gout), None, None, None, None, None, L.ptr(ws), ws.numel(),
        L.stream_ptr()), 'fused_bwd')
    torch.cuda.synchronize()
    assert G.rel(G.nchw(gout), sd.grad) < 1e-6
    assert G.rel(G.nchw(_planes(hi, lo, 1)), bd.grad) < 5e-5


def test_statistics_do_not_depend_on_the_batch():
    """A frame's statistics are bit-identical whatever batch it is normalised in (the cluster
    decomposition depends on (C, H*W) only): needed by chunked streaming inference."""
    lib = L.load()
    c, h, w = 32, 36, 64
    x = torch.randn(5, c, h, w, generator=torch.Generator().manual_seed(3)).cuda()
    outs = []
    for sl in (slice(0, 5), slice(2, 3)):
        xs = G.nhwc(x[sl])
        n = xs.shape[0]
        y = torch.empty_like(xs)
        mean, rstd = torch.empty((n, c), device='cuda'), torch.empty((n, c), device='cuda')
        L.check(lib.eve_instnorm_fused_fwd(L.ptr(xs), None, 0, n, h * w, c, None, None, None, None,
                                           1, 0, L.ptr(mean), L.ptr(rstd), None, None, L.ptr(y),
                                           None, None, None, None, L.stream_ptr()), 'fused_fwd')
        outs.append((y, mean, rstd))
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0][2:3], outs[1][0])
    assert torch.equal(outs[0][1][2:3], outs[1][1]) and torch.equal(outs[0][2][2:3], outs[1][2])


@pytest.mark.parametrize('dual', [False, True])
def test_streaming_backward_agrees_with_the_staged_kernel(dual, in_stream):
    """The streaming kernel performs the same arithmetic per element and the same fixed-order
    reductions as the staged one; only where the second read comes from differs (and how the
    compiler contracts multiply-adds: agreement to a few ulp, not bit-identity)."""
    lib = L.load()
    n, c, h, w = 3, 32, 72, 128
    g = torch.Generator().manual_seed(5)
    x = G.nhwc((torch.randn(n, c, h, w, generator=g) * 1.5 + 0.5).cuda())
    dy, dy2 = (G.nhwc(torch.randn(n, c, h, w, generator=g).cuda()) for _ in range(2))
    add = G.nhwc(torch.randn(n, c, h, w, generator=g).cuda())
    ga, ba, gb, bb = (torch.randn(c, generator=g).cuda() for _ in range(4))
    mean = x.mean(dim=(1, 2)).contiguous()
    rstd = (1.0 / torch.sqrt(x.var(dim=(1, 2), unbiased=False) + 1e-5)).contiguous()
    res = []
    for mode in (0, 2):
        in_stream(mode)
        dx = torch.empty_like(x)
        hi, lo = _u16(x.shape), _u16(x.shape)
        outs = [torch.empty(c, device='cuda') for _ in range(5)]
        ws = torch.empty(lib.eve_instnorm_fused_workspace_bytes(n, h * w, c), dtype=torch.uint8,
                         device='cuda')
        L.check(lib.eve_instnorm_fused_bwd(
            L.ptr(dy), L.ptr(dy2) if dual else None, None, L.ptr(x), n, h * w, c, L.ptr(mean),
            L.ptr(rstd), L.ptr(ga), L.ptr(ba), L.ptr(gb) if dual else None,
            L.ptr(bb) if dual else None, 2, L.ptr(add), L.ptr(dx), L.ptr(hi), L.ptr(lo), None,
            L.ptr(outs[0]), L.ptr(outs[1]), L.ptr(outs[2]) if dual else None,
            L.ptr(outs[3]) if dual else None, L.ptr(outs[4]), L.ptr(ws), ws.numel(),
            L.stream_ptr()), 'fused_bwd')
        torch.cuda.synchronize()
        res.append([dx, hi, lo, outs[0], outs[1], outs[4]] + (outs[2:4] if dual else []))
    for a, b in zip(*res):
        if a.dtype == torch.int16:      # bf16 planes: identical up to the rounding of a last-ulp difference
            a, b = a.view(torch.bfloat16).float(), b.view(torch.bfloat16).float()
            assert float((a - b).abs().max()) <= 2.0 ** -7 * float(b.abs().max())
        else:
            assert float((a - b).abs().max()) <= 2e-6 * float(b.abs().max())
