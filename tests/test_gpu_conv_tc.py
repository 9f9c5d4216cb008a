"""tcgen05 implicit-GEMM convolution (conv_tc.cu) against fp64 arithmetic.

mode 1 (split bf16, three MMAs per K step) must reproduce fp32-level results (<= 3e-5
relative, the bar the parity tests of the whole network rest on); mode 2 (single bf16 pass)
is reported against its expected ~1e-2 error; mode 0 is the fp32 CUDA-core kernel."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from eve_b200 import lib as L            # noqa: E402
from tests import gpu_util as G          # noqa: E402

# (n, cin, h, w, cout, k[, stride]) -- the geometries the tensor-core path takes
TC_CASES = [
    (2, 16, 72, 128, 32, 3),     # RefineNet encoder level 0 (16-channel chunks, SWIZZLE_32B)
    (2, 32, 72, 128, 32, 3),     # ... second conv (32-channel chunks, SWIZZLE_64B)
    (2, 16, 72, 128, 32, 1),     # ... 1x1 skip
    (2, 64, 72, 128, 16, 3),     # decoder level 0: Cout = 16
    (2, 16, 72, 128, 16, 3),     # initial.3 / final.0
    (2, 32, 36, 64, 64, 3),      # encoder level 1 first conv
    (3, 128, 36, 64, 32, 3),     # decoder level 1: Cout = 32
    (3, 64, 32, 32, 128, 3, 2),  # EyeNet layer2.0.conv1 (stride 2, TMA traversal stride)
    (3, 64, 32, 32, 128, 1, 2),  # layer2.0.downsample (1x1 stride 2)
    (5, 256, 8, 8, 512, 3, 2),   # layer4.0.conv1
    (3, 64, 32, 32, 64, 3),      # EyeNet layer1 (box 32x4x1)
    (2, 128, 16, 16, 128, 3),    # layer2 (box 16x8x1)
    (4, 256, 8, 8, 256, 3),      # layer3 (box 8x8x2, two images per tile)
    (9, 512, 4, 4, 512, 3),      # layer4 (box 4x4x8, ragged last tile)
    (2, 64, 36, 64, 64, 3),      # RefineNet level 1 (box 64x2x1)
    (2, 128, 18, 32, 128, 3),    # level 2 (18 rows: partial last tile)
    (2, 256, 9, 16, 256, 3),     # level 3
    (5, 128, 5, 8, 128, 3),      # ConvGRU gates (box 8x5x3 = 120 rows)
    (2, 64, 72, 128, 64, 1),     # 1x1
    (2, 512, 9, 16, 128, 3),     # decoder 512 -> 128
    (1, 64, 72, 128, 64, 3),     # full-resolution map, one image row per tile
]


@pytest.fixture()
def conv_mode():
    lib = L.load()
    prev = lib.eve_get_conv_mode()
    yield lib.eve_set_conv_mode
    lib.eve_set_conv_mode(prev)


@pytest.mark.parametrize('case', TC_CASES)
def test_tensor_core_conv_forward_dgrad_wgrad(case, conv_mode):
    n, cin, h, w, cout, k = case[:6]
    stride = case[6] if len(case) > 6 else 1
    pad = k // 2
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    xd = x.double().requires_grad_(True)
    wd = wt.double().requires_grad_(True)
    y = F.conv2d(xd, wd, b.double(), stride=stride, padding=pad)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy.double())
    errs = {}
    for mode in (1, 2, 0):
        conv_mode(mode)
        got = G.conv_fwd(x.cuda(), wt.cuda(), b.cuda(), stride, pad)
        dx = G.conv_dgrad(dy.cuda(), wt.cuda(), (h, w), stride, pad)
        dw, db = G.conv_wgrad(x.cuda(), dy.cuda(), k, stride, pad)
        torch.cuda.synchronize()
        errs[mode] = (G.rel(got, y), G.rel(dx, xd.grad), G.rel(dw, wd.grad))
        assert G.rel(db, dy.double().sum(dim=(0, 2, 3))) < 2e-5
    assert max(errs[0]) < 2e-5, errs
    assert max(errs[1]) < 3e-5, errs            # split-bf16: fp32-class accuracy
    assert 1e-4 < errs[2][0] < 3e-2, errs       # single bf16 pass really is bf16


def test_mode_switch_is_visible():
    lib = L.load()
    prev = lib.eve_get_conv_mode()
    lib.eve_set_conv_mode(2)
    assert lib.eve_get_conv_mode() == 2
    lib.eve_set_conv_mode(prev)


# (n, cin, h, w, cout, k[, stride]): geometries of the per-tap box kernel with 64 (stacked issue) or
# 128 output channels per tile; odd tile counts leave one CTA of the last pair without a tile
PAIR_CASES = [
    (3, 64, 32, 32, 64, 3),      # 24 tiles
    (5, 64, 32, 32, 64, 3),      # 40 tiles
    (3, 128, 16, 16, 128, 3),    # 6 tiles
    (5, 128, 16, 16, 128, 3),    # 10 tiles
    (3, 256, 8, 8, 256, 3),      # two output-channel tiles, two images per pixel tile, ragged
    (7, 512, 4, 4, 512, 3),      # four output-channel tiles, one ragged pixel tile
    (3, 64, 32, 32, 128, 3, 2),  # stride 2
    (3, 64, 32, 32, 128, 1, 2),  # 1x1 stride 2 (one K stage per tile)
    (300, 64, 32, 32, 64, 3),    # 2400 tiles: every pair walks many pair-rows (the bench regime)
]


@pytest.mark.parametrize('case', PAIR_CASES)
def test_weight_multicast_pairs_are_bit_identical(case, conv_mode):
    """tc_pair: clusters of two CTAs that TMA-multicast each half of a weight stage to both rings.
    Only the delivery of the B operand changes, so forward and data gradient must be bit-identical
    to the single-CTA launch (and the single-CTA launch is held to fp64 by the tests above)."""
    n, cin, h, w, cout, k = case[:6]
    stride = case[6] if len(case) > 6 else 1
    pad = k // 2
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(n, cin, h, w, generator=g).cuda()
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    b = torch.randn(cout, generator=g).cuda()
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    dy = torch.randn(n, cout, oh, ow, generator=g).cuda()
    conv_mode(1)
    saved = L.get_option('tc_pair')
    saved_strip = L.get_option('tc_strip')
    saved_dual = L.get_option('tc_dual')
    try:
        L.set_option('tc_strip', 0)          # keep these geometries on the box kernel
        L.set_option('tc_dual', 0)           # (the pair launch has one issuing warp: compare like with like)
        res = []
        for mode in (0, 2):
            L.set_option('tc_pair', mode)
            y = G.conv_fwd(x, wt, b, stride, pad)
            dx = G.conv_dgrad(dy, wt, (h, w), stride, pad)
            torch.cuda.synchronize()
            res.append((y, dx))
        assert torch.equal(res[0][0], res[1][0])
        assert torch.equal(res[0][1], res[1][1])
        if n <= 8:
            want = F.conv2d(x.double().cpu(), wt.double().cpu(), b.double().cpu(), stride=stride, padding=pad)
            assert G.rel(res[1][0], want) < 3e-5
    finally:
        L.set_option('tc_pair', saved)
        L.set_option('tc_strip', saved_strip)
        L.set_option('tc_dual', saved_dual)
