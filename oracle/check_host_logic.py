"""Host-side logic of the hot path (geometry, gaze-history weighting, validity-masked losses)
checked against the UNMODIFIED reference functions, both on CPU tensors.

Test infrastructure (see oracle/__init__.py): run by tests/test_host_logic.py in a subprocess, in
the container where /root/reference exists.  The reference modules are imported from where they
lie (src/models/common.py, src/losses/*.py); nothing is copied.  Prints one JSON object
{check name: max abs error} and exits 0.
"""
import json
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = '/root/reference/src'


def import_reference():
    for name in ['gspread', 'oauth2client', 'oauth2client.service_account', 'tensorboardX',
                 'coloredlogs', 'h5py', 'ffmpeg']:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['oauth2client.service_account'].ServiceAccountCredentials = object
    sys.modules['tensorboardX'].SummaryWriter = object
    sys.argv[0] = os.path.abspath(__file__)
    os.chdir(REF_SRC)
    sys.path.insert(0, REF_SRC)
    sys.path.insert(1, REPO)
    from core import DefaultConfig
    return DefaultConfig()


def err(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max()) if a.numel() else 0.0


def rand_rotation(g, n):
    q, _ = torch.linalg.qr(torch.randn(n, 3, 3, generator=g))
    return q * torch.sign(torch.linalg.det(q)).view(n, 1, 1)


def main():
    config = import_reference()
    import models.common as RC                       # the reference's src/models/common.py
    from losses.angular import AngularLoss as RAng
    from losses.cross_entropy import CrossEntropyLoss as RCe
    from losses.euclidean import EuclideanLoss as REuc
    from losses.l1 import L1Loss as RL1
    from losses.mse import MSELoss as RMse
    import eve_b200.models.common as PC              # the product's host mirror
    from eve_b200 import losses as PL

    g = torch.Generator().manual_seed(2024)
    out = {}
    n = 17
    py = (torch.rand(n, 2, generator=g) - 0.5) * 1.2
    v3 = torch.randn(n, 3, generator=g)
    out['pitchyaw_to_vector'] = err(PC.pitchyaw_to_vector(py), RC.pitchyaw_to_vector(py))
    out['vector_to_pitchyaw'] = err(PC.vector_to_pitchyaw(v3), RC.vector_to_pitchyaw(v3))
    out['pitchyaw_to_rotation'] = err(PC.pitchyaw_to_rotation(py), RC.pitchyaw_to_rotation(py))
    R = rand_rotation(g, n)
    T = torch.eye(4).repeat(n, 1, 1)
    T[:, :3, :3] = rand_rotation(g, n)
    T[:, :3, 3] = torch.randn(n, 3, generator=g) * 50.0
    out['apply_transformation'] = err(PC.apply_transformation(T, v3), RC.apply_transformation(T, v3))
    out['apply_rotation'] = err(PC.apply_rotation(T, v3), RC.apply_rotation(T, v3))
    o3 = torch.randn(n, 3, generator=g) * 100.0 + torch.tensor([0.0, 0.0, 500.0])
    g3 = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) - torch.tensor([0.0, 0.0, 2.0]), dim=-1)
    out['get_intersect_with_zero'] = err(PC.get_intersect_with_zero(o3, g3),
                                         RC.get_intersect_with_zero(o3, g3))
    pog = torch.randn(n, 2, generator=g) * 150.0
    out['calculate_combined_gaze_direction'] = err(
        PC._combined_gaze_torch(o3, pog, R, T),
        RC.calculate_combined_gaze_direction(o3, pog, R, T))
    kappa = torch.randn(n, 2, generator=g) * 0.05
    for inv in (False, True):
        out['apply_offset_augmentation/inverse=%d' % inv] = err(
            PC._offset_augmentation_torch(py, R, kappa, inverse_kappa=inv),
            RC.apply_offset_augmentation(py, R, kappa, inverse_kappa=inv))

    # gaze history maps: the reference's O(T^2) Python loops against the batched weights
    B, Tn, H, W = 3, 7, 6, 8
    ts = torch.cumsum(torch.randint(20_000_000, 50_000_000, (B, Tn), generator=g), dim=1)
    ts[0, 5:] = 0                                    # a padded tail
    ts[2, 0] = 0                                     # a dropped first frame
    val = torch.rand(B, Tn, generator=g) > 0.25
    hms = torch.rand(B, Tn, 1, H, W, generator=g)
    worst_b, worst_all = 0.0, 0.0
    allmaps = PC._all_gaze_history_maps_torch(ts, hms, val)
    for t in range(1, Tn + 1):
        if bool((ts[:, :t] == 0).all(dim=1).any()):
            continue                                 # the reference cannot index an all-zero prefix
        lst = [hms[:, i] for i in range(t)]
        want = RC.batch_make_gaze_history_maps(ts, lst, val)
        worst_b = max(worst_b, err(PC._batch_make_gaze_history_maps_torch(ts, lst, val), want))
        worst_all = max(worst_all, err(allmaps[:, t - 1], want))
    out['batch_make_gaze_history_maps'] = worst_b
    out['all_gaze_history_maps'] = worst_all
    one = RC.make_gaze_history_map(ts[1], [hms[1, i] for i in range(Tn)], val[1])
    out['make_gaze_history_map'] = err(PC._batch_make_gaze_history_maps_torch(
        ts[1:2], [hms[1:2, i] for i in range(Tn)], val[1:2])[0], one)

    # validity-masked sequence losses (loop over the batch in the reference, one expression here)
    B, Tn = 5, 9
    valid = torch.rand(B, Tn, generator=g) > 0.3
    valid[1] = False                                 # no valid frame at all
    valid[2] = False
    valid[2, 4] = True                               # exactly one valid frame (the n_valid > 1 rule)
    ref = {'g_validity': valid, 'p_validity': valid, 'h_validity': valid, 's_validity': valid}
    a2, b2 = (torch.rand(B, Tn, 2, generator=g) - 0.5), (torch.rand(B, Tn, 2, generator=g) - 0.5)
    ref['g'] = b2
    out['loss/angular'] = err(PL.angular_loss.torch_formula(a2, 'g', ref), RAng()(a2, 'g', ref))
    ref['p'] = b2 * 300.0
    out['loss/euclidean'] = err(PL.euclidean_loss.torch_formula(a2 * 300.0, 'p', ref), REuc()(a2 * 300.0, 'p', ref))
    out['loss/mse'] = err(PL.mse_loss.torch_formula(a2, 'g', ref), RMse()(a2, 'g', ref))
    out['loss/l1'] = err(PL.l1_loss.torch_formula(a2, 'g', ref), RL1()(a2, 'g', ref))
    s1, s2 = torch.rand(B, Tn, generator=g), torch.rand(B, Tn, generator=g)
    ref['s'] = s2
    out['loss/mse_scalar'] = err(PL.mse_loss.torch_formula(s1, 's', ref), RMse()(s1, 's', ref))
    out['loss/l1_scalar'] = err(PL.l1_loss.torch_formula(s1, 's', ref), RL1()(s1, 's', ref))
    h1 = torch.rand(B, Tn, 1, 6, 8, generator=g).clamp(1e-4, 1 - 1e-4)
    h2 = torch.rand(B, Tn, 1, 6, 8, generator=g)
    ref['h'] = h2
    out['loss/cross_entropy'] = err(PL.cross_entropy_loss.torch_formula(h1, 'h', ref), RCe()(h1, 'h', ref))
    out['loss/mse_heatmap'] = err(PL.mse_loss.torch_formula(h1, 'h', ref), RMse()(h1, 'h', ref))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
