"""Heatmap raster, soft-argmax and PoG projection kernels against the CPU oracle
(fp64) and the reference's golden vectors."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import eve_oracle as O       # noqa: E402  (checker only)
from tests import gpu_util as G          # noqa: E402
from tests import helpers as H           # noqa: E402


def test_heatmap_and_soft_argmax_match_reference_vectors(cfg):
    from eve_b200.models import common as MC
    gold = H.load_golden('modules')
    px = torch.from_numpy(gold['refine/px_in']).cuda()
    hm = MC.batch_make_heatmaps(px, cfg.gaze_heatmap_sigma_initial)
    assert H.rel_err(hm.cpu().numpy(), gold['refine/heatmap_initial']) < 1e-5
    fin = torch.from_numpy(gold['refine/heatmap_final']).cuda()
    assert H.rel_err(MC.soft_argmax(fin).cpu().numpy(), gold['refine/softargmax']) < 1e-4
    onehot = torch.zeros(1, 1, 72, 128, device='cuda')
    onehot[0, 0, 10, 20] = 1.0
    got = MC.soft_argmax(onehot).cpu().numpy()
    assert H.rel_err(got, gold['refine/softargmax_onehot']) < 1e-5
    assert abs(got[0, 0] - 302.36) < 0.05 and abs(got[0, 1] - 152.11) < 0.05


@pytest.mark.parametrize('sigma', [10.0, 3.0, 5.0])
def test_heatmap_backward(cfg, sigma):
    from eve_b200.models import common as MC
    g = torch.Generator().manual_seed(1)
    c = torch.stack([torch.rand(6, generator=g) * 1920, torch.rand(6, generator=g) * 1080], -1)
    c[0] = torch.tensor([0.0, 0.0])              # on the corner
    c[1] = torch.tensor([1920.0, 1080.0])
    cd = c.double().requires_grad_(True)
    want = O.make_heatmaps(cd, sigma)
    dy = torch.randn(want.shape, generator=g)
    want.backward(dy.double())
    cc = c.cuda().requires_grad_(True)
    got = MC.batch_make_heatmaps(cc, sigma)
    assert G.rel(got, want) < 1e-5
    got.backward(dy.cuda())
    assert G.rel(cc.grad, cd.grad) < 1e-4


def test_soft_argmax_forward_backward(cfg):
    from eve_b200.models import common as MC
    g = torch.Generator().manual_seed(2)
    h = torch.rand(5, 1, 72, 128, generator=g) * 0.05
    h[0, 0, 40, 100] = 0.9                       # a peaked map
    h[1] = 0.5                                   # a flat map (the zero-init RefineNet output)
    hd = h.double().requires_grad_(True)
    want = O.soft_argmax(hd)
    dy = torch.randn(want.shape, generator=g)
    want.backward(dy.double())
    hc = h.cuda().requires_grad_(True)
    got = MC.soft_argmax(hc)
    assert G.rel(got, want) < 1e-5
    got.backward(dy.cuda())
    assert G.rel(hc.grad, hd.grad) < 1e-4


def test_pog_projection_forward_backward(cfg):
    from eve_b200 import synth
    from eve_b200.models import common as MC
    inp = synth.make_clip_batch(3, 5, seed=9, with_screen=False)
    g = torch.Generator().manual_seed(3)
    gaze = (torch.rand(3, 5, 2, generator=g) - 0.5) * 0.8
    gaze[0, 0] = torch.tensor([1.2, -1.3])       # far off screen -> clamped pixels
    gd = gaze.double().requires_grad_(True)
    dd = {k: v.double() if v.is_floating_point() else v for k, v in inp.items()}
    mm, px = O.to_screen_coordinates(dd['left_o'], gd, dd['left_R'],
                                     dd['inv_camera_transformation'],
                                     dd['pixels_per_millimeter'])
    dmm = torch.randn(mm.shape, generator=g)
    dpx = torch.randn(px.shape, generator=g)
    (mm * dmm.double()).sum().backward(retain_graph=True)
    (px * dpx.double()).sum().backward()
    cu = {k: v.cuda() for k, v in inp.items()}
    gc = gaze.cuda().requires_grad_(True)
    gmm, gpx = MC.to_screen_coordinates(cu['left_o'], gc, cu['left_R'], cu)
    assert G.rel(gmm, mm) < 2e-5
    assert G.rel(gpx, px) < 2e-5
    ((gmm * dmm.cuda()).sum() + (gpx * dpx.cuda()).sum()).backward()
    assert G.rel(gc.grad, gd.grad) < 1e-4


def test_geometry_helpers_match_oracle(cfg):
    from eve_b200 import synth
    from eve_b200.models import common as MC
    inp = synth.make_clip_batch(2, 4, seed=5, with_screen=False)
    cu = {k: v.cuda() for k, v in inp.items()}
    o = 0.5 * (inp['left_o'] + inp['right_o'])
    pog = torch.rand(2, 4, 2) * torch.tensor([553.0, 311.0])
    want = O.combined_gaze_direction(o, pog, inp['left_R'], inp['camera_transformation'])
    got = MC.calculate_combined_gaze_direction(o.cuda(), pog.cuda(), cu['left_R'],
                                               cu['camera_transformation'])
    assert G.rel(got, want) < 1e-5
    kappa = torch.randn(2, 1, 2).expand(2, 4, 2) * 0.05
    gz = (torch.rand(2, 4, 2) - 0.5) * 0.5
    want = O.offset_augmentation(gz, inp['head_R'], kappa)
    got = MC.apply_offset_augmentation(gz.cuda(), cu['head_R'], kappa.cuda())
    assert G.rel(got, want) < 1e-5
    # gaze-history maps: all prefixes at once vs the oracle's restatement of common.py:249-287
    hm = torch.rand(2, 4, 1, 72, 128)
    val = torch.tensor([[True, True, False, True], [True, False, True, True]])
    ts = inp['timestamps'].clone()
    ts[1, 3] = 0                                   # a padded frame
    want = O.gaze_history_maps(ts, hm, val, cfg.gaze_history_map_decay_per_ms)
    got = MC.all_gaze_history_maps(ts.cuda(), hm.cuda(), val.cuda())
    assert G.rel(got, want) < 1e-5
    last = MC.batch_make_gaze_history_maps(ts.cuda(), [hm[:, t].cuda() for t in range(3)],
                                           val.cuda())
    assert G.rel(last, want[:, 2]) < 1e-5
