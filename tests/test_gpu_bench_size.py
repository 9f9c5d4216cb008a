"""Parity at the sizes bench.py actually times (BASELINE configs[1] and [2]: B=8 clips, T=30).

Split-K chain lengths, wave counts, TMA box decompositions, row-strip plans and the cluster
decomposition of the normalisation kernels all depend on the batch (N = 240 screen frames /
480 eye patches), so the toy-size parity tests do not cover the configuration the benchmark
runs.  Here the whole optimisation step is compared with the CPU oracle (fp32, one step takes
some tens of seconds on the host cores) and the convolution passes are compared with fp64
library convolutions at N = 240 / 480.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from eve_b200 import lib as L            # noqa: E402
from oracle import eve_oracle as O       # noqa: E402   (checker only)
from tests import gpu_util as G          # noqa: E402


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _oracle(cfg, sd, inputs, B, seed, dtype):
    """The CPU oracle on the same inputs, weights and kappa draws, evaluated in `dtype`."""
    np.random.seed(seed)
    std = np.radians(cfg.refine_net_offset_augmentation_sigma)
    kap = {k: torch.from_numpy(np.random.normal(size=(B, 2), scale=std).astype(np.float32)).to(dtype)
           for k in ('left', 'right')}
    osd = {k: v.clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
    inp = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in inputs.items()}
    want, wmid = O.eve_forward(osd, cfg, inp, True, kap)
    want['full_loss'].backward()
    both = dict(wmid)
    both.update(want)
    return ({k: v.detach() for k, v in both.items() if torch.is_tensor(v)},
            {k: v.grad for k, v in osd.items() if v.grad is not None})


def _gpu_step(sd, inputs, seed, mode):
    from eve_b200.models import EVE
    lib = L.load()
    prev = lib.eve_get_conv_mode()
    lib.eve_set_conv_mode(mode)
    try:
        model = EVE(output_predictions=True)
        model.load_state_dict(sd, strict=True)
        model = model.cuda().train()
        np.random.seed(seed)
        out = model({'bench': {k: v.cuda() for k, v in inputs.items()}}, current_epoch=0.0)
        out['full_loss'].backward()
        torch.cuda.synchronize()
        got = {k: v.detach().cpu() for k, v in out.items() if torch.is_tensor(v)}
        grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}
        del model, out
        torch.cuda.empty_cache()
    finally:
        lib.eve_set_conv_mode(prev)
    return got, grads


def _step(cfg, B, T, seed, with_refine):
    """One optimisation step four ways: the product (conv mode 1: tcgen05 split operands), the
    library's exact-fp32 CUDA-core mode 0, the CPU oracle in fp32 (= the reference's arithmetic)
    and the CPU oracle in fp64 (the yardstick)."""
    from eve_b200 import synth
    sd = synth.make_state_dict(synth.eye_net_param_shapes(cfg), seed, 'eye_net.')
    if with_refine:
        sd.update(synth.make_state_dict(synth.refine_net_param_shapes(cfg), seed + 1000, 'refine_net.'))
    inputs = synth.make_clip_batch(B, T, seed=seed, with_screen=with_refine)
    got, grads = _gpu_step(sd, inputs, seed, 1)
    _, grads0 = _gpu_step(sd, inputs, seed, 0)
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    want, g32 = _oracle(cfg, sd, inputs, B, seed, torch.float32)
    _, g64 = _oracle(cfg, sd, inputs, B, seed, torch.float64)
    return got, grads, want, (g32, grads0), g64


def _grad_check(grads, fp32_runs, g64, prefix, factor, floor):
    """Weight gradients against the fp64 evaluation of the oracle.  The networks are ill
    conditioned at random weights (InstanceNorm over 4x4 .. 5x8 maps, nearly constant RefineNet
    planes): plain fp32 arithmetic -- the CPU oracle, and this library's exact-fp32 CUDA-core
    mode 0 -- sits 1e-3 .. 1e-2 from the fp64 result, so the bar is relative to that measured
    noise: per parameter tensor, the product's L2 distance from fp64 must stay within `factor` x
    the larger of the two fp32 distances -- or of the median fp32 distance over the network's
    tensors, whichever is larger -- plus a small absolute floor.  tools/grad_precision.py
    shows where the remaining factor comes from: running dgrad / wgrad on fp32 CUDA cores
    (EVE_B200_TC_MASK=1) leaves it unchanged, i.e. it is the FORWARD's 22-bit operand planes and
    TMEM's round-toward-zero accumulation, not the bf16 gradient planes."""
    g32, grads0 = fp32_runs
    assert grads.keys() == g32.keys() == g64.keys() == grads0.keys()
    top = max(float(v.norm()) for v in g64.values())
    rows = []
    for k in grads:
        if not k.startswith(prefix):
            continue
        den = float(g64[k].norm())
        if den < 1e-5 * top:
            # biases in front of an InstanceNorm: exactly zero gradient, only cancellation noise
            assert float((grads[k].double() - g64[k]).norm()) < 1e-5 * top, k
            continue
        rows.append((k, max(_l2(g32[k], g64[k]), _l2(grads0[k], g64[k])), _l2(grads[k], g64[k]),
                     _l2(g32[k], g64[k]), _l2(grads0[k], g64[k])))
    assert rows
    # the yardstick of a tensor is never smaller than the network's median fp32 distance: where plain
    # fp32 happens to land 10x closer to fp64 than it typically does (6e-4 on one norm gain of the
    # outermost decoder block against a median of 7e-3), that is luck of the rounding order -- any
    # legitimate reordering of a sum upstream (a different InstanceNorm reduction tree) moves it --
    # not a precision level the split-operand path could be held to
    med = float(np.median([r[1] for r in rows]))
    bad = [(k, o, e) for k, o, e, _, _ in rows if e > factor * max(o, med) + floor]
    print('%s gradients vs fp64 (median / max): oracle-fp32 %.2e / %.2e | mode 0 %.2e / %.2e | '
          'product %.2e / %.2e' % (
              prefix, np.median([r[3] for r in rows]), max(r[3] for r in rows),
              np.median([r[4] for r in rows]), max(r[4] for r in rows),
              np.median([r[2] for r in rows]), max(r[2] for r in rows)))
    assert not bad, bad[:6]
    assert np.median([r[2] for r in rows]) <= factor * np.median([r[1] for r in rows]) + floor
    return rows


def test_config3_full_eve_step_matches_the_oracle_at_B8_T30(cfg):
    """BASELINE configs[2]: EyeNet x2 + GazeRefineNet (CGRU), B=8, T=30 -- the bench workload.
    Forward: the north_star bar is 1e-3 relative on gaze vectors / PoG; measured 1e-5..2e-4 here.
    Gradients: see _grad_check (measured on B200: ours <= 1.5x the fp32 oracle's own distance
    from fp64)."""
    cfg.override('refine_net_enabled', True)
    cfg.override('load_screen_content', True)
    got, grads, want, fp32_runs, g64 = _step(cfg, 8, 30, 3, True)
    fwd = {}
    for k in ('g_initial', 'g_final', 'PoG_px_initial', 'PoG_px_final', 'PoG_cm_initial',
              'PoG_cm_final', 'left_pupil_size', 'right_pupil_size', 'full_loss',
              'loss_ce_heatmap_final', 'loss_mse_PoG_cm_final', 'metric_euc_PoG_px_final',
              'metric_ang_g_final', 'loss_ang_left_g_initial'):
        assert k in got and k in want, k
        fwd[k] = _rel(got[k], want[k])
    print('config3 forward rel errors:', fwd)
    assert all(e < 1e-3 for e in fwd.values()), fwd          # the north_star bar
    assert fwd['g_initial'] < 2e-5 and fwd['PoG_px_initial'] < 5e-5, fwd
    assert fwd['g_final'] < 5e-4 and fwd['PoG_px_final'] < 5e-4, fwd
    _grad_check(grads, fp32_runs, g64, 'refine_net.', 2.5, 2e-3)
    _grad_check(grads, fp32_runs, g64, 'eye_net.', 2.5, 2e-3)


def test_config2_static_eyenet_step_matches_the_oracle_at_B8_T30(cfg):
    """BASELINE configs[1]: EyeNet static (eye_net_use_rnn=0, no RefineNet), B=8, T=30 = 480 eye
    patches through the ResNet-18/InstanceNorm encoder.  Measured on B200: fp32 oracle 1.3e-3
    median / 5.4e-3 max from fp64, tcgen05 split-operand path 1.7e-3 / 7.5e-3."""
    cfg.override('refine_net_enabled', False)
    cfg.override('load_screen_content', False)
    cfg.override('eye_net_use_rnn', False)
    got, grads, want, fp32_runs, g64 = _step(cfg, 8, 30, 4, False)
    fwd = {k: _rel(got[k], want[k]) for k in ('g_initial', 'PoG_px_initial', 'left_pupil_size',
                                               'full_loss', 'loss_ang_left_g_initial')}
    print('config2 forward rel errors:', fwd)
    assert all(e < 1.5e-4 for e in fwd.values()), fwd        # measured 0 .. 5.6e-5
    rows = _grad_check(grads, fp32_runs, g64, 'eye_net.', 2.0, 1e-5)
    assert max(r[2] for r in rows) < 1.2e-2


# (n, cin, cout, h, w, k, stride): the bench geometry of each kernel family
BIG_CONVS = [(240, 16, 16, 72, 128, 3, 1),     # halo-row kernels, RefineNet level 0
             (240, 64, 16, 72, 128, 1, 1),
             (240, 32, 64, 36, 64, 3, 1),      # generic kernel, level 1
             (240, 256, 256, 9, 16, 3, 1),     # level 3: long split-K chains
             (240, 128, 64, 5, 8, 3, 1),       # bottleneck batched weight gradient
             (480, 64, 64, 32, 32, 3, 1),      # EyeNet layer1
             (480, 64, 128, 32, 32, 3, 2),     # EyeNet layer2 stride 2 + its 1x1 downsample
             (480, 64, 128, 32, 32, 1, 2),
             (480, 512, 512, 4, 4, 3, 1)]      # EyeNet layer4


@pytest.mark.parametrize('case', BIG_CONVS, ids=lambda c: 'x'.join(map(str, c)))
def test_conv_passes_at_bench_batch_match_fp64(case):
    """fwd / dgrad / wgrad / bias gradient at N = 240 (screen frames) and 480 (eye patches) against
    fp64 library convolutions on the same device (cuDNN fp64 is the checker here, nothing more)."""
    n, cin, cout, h, w, k, stride = case
    L.load().eve_set_conv_mode(1)
    g = torch.Generator(device='cuda').manual_seed(n + cin + cout)
    x = torch.randn(n, cin, h, w, generator=g, device='cuda')
    wt = torch.randn(cout, cin, k, k, generator=g, device='cuda') / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g, device='cuda')
    pad = k // 2
    xd, wd, bd = x.double().requires_grad_(True), wt.double().requires_grad_(True), \
        b.double().requires_grad_(True)
    y = F.conv2d(xd, wd, bd, stride=stride, padding=pad)
    dy = torch.randn(y.shape, generator=g, device='cuda')
    y.backward(dy.double())
    got = G.conv_fwd(x, wt, b, stride, pad)
    dx = G.conv_dgrad(dy, wt, (h, w), stride, pad)
    dw, db = G.conv_wgrad(x, dy, k, stride, pad)
    torch.cuda.synchronize()
    assert _rel(got, y) < 2e-5
    assert _rel(dx, xd.grad) < 3e-5
    # weight gradients sum N*OH*OW products per element: relative to the largest element
    assert _rel(dw, wd.grad) < 5e-5
    assert _rel(db, bd.grad) < 5e-5


def test_stem_maxpool_indices_with_ties_match_aten():
    """torchvision ResNet.maxpool (3x3, stride 2, pad 1) behind relu(IN(x)): values and the argmax
    INDEX output are bit-exact against F.max_pool2d(return_indices=True).  The input only takes
    three distinct values, so almost every window holds its maximum more than once: the first
    maximum in row-major window order must win."""
    lib = L.load()
    n, c, h, w = 3, 64, 64, 64
    g = torch.Generator().manual_seed(8)
    x = torch.randint(0, 3, (n, c, h, w), generator=g).float()
    xh = G.nhwc(x.cuda())
    oh, ow = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    y = torch.empty((n, oh, ow, c), device='cuda')
    idx = torch.empty((n, oh, ow, c), dtype=torch.int32, device='cuda')
    mean, rstd = torch.empty((n, c), device='cuda'), torch.empty((n, c), device='cuda')
    L.check(lib.eve_in_relu_maxpool_fwd(L.ptr(xh), n, h, w, c, L.ptr(mean), L.ptr(rstd), L.ptr(y),
                                        L.ptr(idx), L.stream_ptr()), 'in_relu_maxpool')
    torch.cuda.synchronize()
    # the same normalised tensor ATen would pool (statistics from the kernel under test, so that
    # equal inputs stay bit-equal and the comparison is about the pooling rule alone)
    act = F.relu((x.cuda() - mean.view(n, c, 1, 1)) * rstd.view(n, c, 1, 1))
    want, widx = F.max_pool2d(act, 3, 2, 1, return_indices=True)
    assert torch.equal(G.nchw(y), want)
    assert torch.equal(idx.permute(0, 3, 1, 2).long(), widx)
    # how many windows actually hold a tie
    cols = F.unfold(F.pad(act[:1, :4], (1, 1, 1, 1), value=float('-inf')), 3, stride=2)
    cols = cols.reshape(1, 4, 9, -1)
    tied = ((cols == cols.max(dim=2, keepdim=True)[0]).sum(dim=2) > 1).float().mean()
    assert float(tied) > 0.5, float(tied)
    assert _rel(mean, x.double().mean(dim=(2, 3))) < 1e-6
