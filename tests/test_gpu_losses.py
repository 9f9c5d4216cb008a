"""Fused label / loss / geometry kernels against the
torch formulas that oracle/check_host_logic.py pins to the unmodified reference functions, against
the oracle, and against the labels stored in the golden fixtures."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from eve_b200 import losses as LS        # noqa: E402
from eve_b200 import ops                 # noqa: E402
from eve_b200.models import common as MC  # noqa: E402
from oracle import eve_oracle as O       # noqa: E402   (checker only)
from tests import gpu_util as G          # noqa: E402
from tests import helpers as H           # noqa: E402


def _rot(n, g):
    return MC.pitchyaw_to_rotation(torch.rand(n, 2, generator=g, dtype=torch.float64) - 0.5)


def test_combined_gaze_kernel_forward_and_gradient():
    g = torch.Generator().manual_seed(0)
    B, T = 3, 7
    n = B * T
    o = (torch.randn(n, 3, generator=g, dtype=torch.float64) * 5 + torch.tensor([0.0, 0.0, 600.0]))
    pog = (torch.rand(n, 2, generator=g, dtype=torch.float64) * 500).requires_grad_(True)
    R = _rot(n, g)
    cam = torch.eye(4, dtype=torch.float64).repeat(n, 1, 1)
    cam[:, :3, :3] = _rot(n, g)
    cam[:, :3, 3] = torch.randn(n, 3, generator=g, dtype=torch.float64) * 20
    want = MC._combined_gaze_torch(o, pog, R, cam)
    dg = torch.randn(n, 2, generator=g, dtype=torch.float64)
    want.backward(dg)
    pc = pog.detach().float().cuda().reshape(B, T, 2).requires_grad_(True)
    got = MC.calculate_combined_gaze_direction(o.float().cuda().reshape(B, T, 3), pc,
                                               R.float().cuda().reshape(B, T, 3, 3),
                                               cam.float().cuda().reshape(B, T, 4, 4))
    assert got.shape == (B, T, 2)
    assert G.rel(got.reshape(n, 2), want) < 2e-6
    got.backward(dg.float().cuda().reshape(B, T, 2))
    assert G.rel(pc.grad.reshape(n, 2), pog.grad) < 2e-5


@pytest.mark.parametrize('inverse', [False, True])
@pytest.mark.parametrize('expanded', [False, True])
def test_offset_augmentation_kernel_forward_and_gradient(inverse, expanded):
    g = torch.Generator().manual_seed(1)
    B, T = 4, 6
    gz = ((torch.rand(B, T, 2, generator=g, dtype=torch.float64) - 0.5) * 0.8).requires_grad_(True)
    R = _rot(B * T, g).reshape(B, T, 3, 3)
    if expanded:      # one kappa per clip, repeated over time (eve.py:466-477)
        kappa = (torch.randn(B, 1, 2, generator=g, dtype=torch.float64) * 0.05).expand(B, T, 2)
    else:
        kappa = torch.randn(B, T, 2, generator=g, dtype=torch.float64) * 0.05
    want = MC._offset_augmentation_torch(gz, R, kappa, inverse_kappa=inverse)
    dout = torch.randn(B, T, 2, generator=g, dtype=torch.float64)
    want.backward(dout)
    gc = gz.detach().float().cuda().requires_grad_(True)
    kc = kappa.float().cuda() if not expanded else kappa[:, :1].float().cuda().expand(B, T, 2)
    got = MC.apply_offset_augmentation(gc, R.float().cuda(), kc, inverse_kappa=inverse)
    assert G.rel(got, want) < 2e-6
    got.backward(dout.float().cuda())
    assert G.rel(gc.grad, gz.grad) < 2e-5


def test_history_recurrence_matches_the_explicit_weights_and_its_gradient(cfg):
    g = torch.Generator().manual_seed(2)
    B, T, Hh, Ww = 3, 9, 72, 128
    ts = torch.cumsum(torch.randint(20_000_000, 50_000_000, (B, T), generator=g), dim=1)
    ts[0, 6:] = 0                 # padded tail
    ts[2, 0] = 0                  # dropped first frame
    ts[1, 4] = 0                  # a hole in the middle
    val = torch.rand(B, T, generator=g) > 0.25
    hm = torch.rand(B, T, 1, Hh, Ww, generator=g, dtype=torch.float64).requires_grad_(True)
    want = MC._all_gaze_history_maps_torch(ts, hm, val)
    dout = torch.randn(want.shape, generator=g, dtype=torch.float64)
    want.backward(dout)
    hc = hm.detach().float().cuda().requires_grad_(True)
    got = MC.all_gaze_history_maps(ts.cuda(), hc, val.cuda())
    assert G.rel(got, want) < 2e-6
    got.backward(dout.float().cuda())
    assert G.rel(hc.grad, hm.grad) < 2e-6
    # the oracle's restatement of the reference's own loops
    want2 = O.gaze_history_maps(ts, hm.detach().float(), val, cfg.gaze_history_map_decay_per_ms)
    assert G.rel(got, want2) < 1e-5


def _validity(B, T, g):
    v = torch.rand(B, T, generator=g) > 0.3
    v[1] = False                  # a clip without any valid frame
    v[2] = False
    v[2, T // 2] = True           # exactly one valid frame: the n_valid > 1 rule
    return v


def test_masked_loss_table_values_and_gradients():
    g = torch.Generator().manual_seed(3)
    B, T = 5, 11
    v, v2 = _validity(B, T, g), torch.rand(B, T, generator=g) > 0.2
    a2 = (torch.rand(B, T, 2, generator=g, dtype=torch.float64) - 0.5).requires_grad_(True)
    b2 = torch.rand(B, T, 2, generator=g, dtype=torch.float64) - 0.5
    p2 = (torch.rand(B, T, 2, generator=g, dtype=torch.float64) * 300).requires_grad_(True)
    q2 = torch.rand(B, T, 2, generator=g, dtype=torch.float64) * 300
    s1 = torch.rand(B, T, generator=g, dtype=torch.float64).requires_grad_(True)
    s2 = torch.rand(B, T, generator=g, dtype=torch.float64)
    ref = {'g': b2, 'g_validity': v, 'p': q2, 'p_validity': v & v2, 's': s2, 's_validity': v2}
    table = [(LS.angular_loss, a2, 'g'), (LS.mse_loss, a2, 'g'), (LS.l1_loss, a2, 'g'),
             (LS.euclidean_loss, p2, 'p'), (LS.mse_loss, p2, 'p'), (LS.l1_loss, s1, 's'),
             (LS.mse_loss, s1, 's')]
    want = [loss.torch_formula(pred, key, ref) for loss, pred, key in table]
    w = torch.randn(len(table), generator=g, dtype=torch.float64)
    sum(wi * li for wi, li in zip(w, want)).backward()
    cu = {id(t): t.detach().float().cuda().requires_grad_(True) for t in (a2, p2, s1)}
    cref = {k: (t.cuda() if t.dtype == torch.bool else t.float().cuda()) for k, t in ref.items()}
    got = LS.evaluate_terms([(loss, cu[id(pred)], cref[key], cref[key + '_validity'], None)
                             for loss, pred, key in table])
    for gi, wi in zip(got, want):
        assert abs(float(gi) - float(wi)) < 2e-6 * max(1.0, abs(float(wi))), (float(gi), float(wi))
    sum(float(wi) * gi for wi, gi in zip(w, got)).backward()
    for t in (a2, p2, s1):
        assert G.rel(cu[id(t)].grad, t.grad) < 2e-5
    # the public per-loss call contract (losses/base_loss_with_validity.py:32-73)
    one = LS.angular_loss(cu[id(a2)].detach(), 'g', cref)
    assert abs(float(one) - float(want[0])) < 2e-6 * max(1.0, abs(float(want[0])))
    # second validity mask (left-right consistency terms)
    both = LS.evaluate_terms([(LS.mse_loss, cu[id(p2)].detach(), cref['p'], cref['g_validity'],
                               cref['s_validity'])])[0]
    assert abs(float(both) - float(want[4])) < 2e-6 * max(1.0, abs(float(want[4])))


def test_heatmap_losses_values_and_gradients():
    g = torch.Generator().manual_seed(4)
    B, T = 4, 5
    v = _validity(B, T, g)
    pred = torch.sigmoid(torch.randn(B, T, 1, 72, 128, generator=g, dtype=torch.float64) * 4)
    pred[0, 0, 0, 0, :4] = torch.tensor([0.0, 1.0, 1e-30, 1.0 - 1e-12])   # saturated sigmoid outputs
    pred = pred.requires_grad_(True)
    gt = (torch.rand(B, T, 1, 72, 128, generator=g, dtype=torch.float64)
          * v.double().view(B, T, 1, 1, 1))
    ref = {'h': gt, 'h_validity': v}
    p32 = pred.detach().float().requires_grad_(True)
    r32 = {'h': gt.float(), 'h_validity': v}
    wce = LS.cross_entropy_loss.torch_formula(p32, 'h', r32)
    wms = LS.mse_loss.torch_formula(p32, 'h', r32)
    (0.7 * wce + 0.3 * wms).backward()
    pc = pred.detach().float().cuda().requires_grad_(True)
    cref = {'h': gt.float().cuda(), 'h_validity': v.cuda()}
    gce, gms = LS.evaluate_terms([(LS.cross_entropy_loss, pc, cref['h'], cref['h_validity'], None),
                                  (LS.mse_loss, pc, cref['h'], cref['h_validity'], None)])
    assert abs(float(gce) - float(wce)) < 1e-5 * max(1.0, abs(float(wce)))
    assert abs(float(gms) - float(wms)) < 1e-5 * max(1.0, abs(float(wms)))
    (0.7 * gce + 0.3 * gms).backward()
    assert torch.isfinite(pc.grad).all()
    assert G.rel(pc.grad, p32.grad) < 1e-5
    # only the BCE term in the loss: the MSE gradient slot stays empty
    pc2 = pred.detach().float().cuda().requires_grad_(True)
    gce2 = LS.cross_entropy_loss(pc2, 'h', cref)
    gce2.backward()
    p33 = pred.detach().float().requires_grad_(True)
    LS.cross_entropy_loss.torch_formula(p33, 'h', r32).backward()
    assert G.rel(pc2.grad, p33.grad) < 1e-5


@pytest.mark.parametrize('name', ['eve_cgru_train', 'eve_cgru_eval_pad', 'eve_clstm_frozen_train'])
def test_label_kernels_match_the_reference_labels_in_the_goldens(cfg, name):
    """calculate_additional_labels (eve.py:441-543): every label/* array the unmodified reference
    produced for the golden cases, from the two label kernels."""
    from eve_b200.models import EVE
    names = H.golden_names()
    if name not in names:
        pytest.skip('golden %s not present' % name)
    gold = H.load_golden(name)
    H.apply_case_config(cfg, gold)
    inputs = {k: v.cuda() for k, v in H.case_inputs(gold, cfg).items()}
    model = EVE(output_predictions=True)
    model.train(bool(gold['meta/training']))
    np.random.seed(int(gold['meta/seed']))
    model.calculate_additional_labels(inputs, current_epoch=0.0)
    checked = 0
    for k in gold:
        if not k.startswith('label/'):
            continue
        key = k[len('label/'):]
        assert key in inputs, key
        got = inputs[key].detach().cpu().numpy()
        want = gold[k]
        assert got.shape == want.shape, (key, got.shape, want.shape)
        if want.dtype == np.bool_:
            assert np.array_equal(got.astype(np.bool_), want), key
        else:
            assert H.rel_err(got, want) < 2e-6, (key, H.rel_err(got, want))
        checked += 1
    assert checked >= 5
