import sys, torch, numpy as np
sys.path.insert(0, '.')
from eve_b200.config import DefaultConfig
from eve_b200 import synth, lib as L
from eve_b200.models import EVE
lib = L.load()
cfg = DefaultConfig(); cfg.reset()
cfg.override('refine_net_enabled', True); cfg.override('load_screen_content', True)
B, T = int(sys.argv[1]), int(sys.argv[2])
sd = synth.make_state_dict(synth.eye_net_param_shapes(cfg), 0, 'eye_net.')
sd.update(synth.make_state_dict(synth.refine_net_param_shapes(cfg), 1000, 'refine_net.'))
inputs = {k: v.cuda() for k, v in synth.make_clip_batch(B, T, seed=3).items()}
res = {}
for mode in (0, 1):
    lib.eve_set_conv_mode(mode)
    model = EVE(); model.load_state_dict(sd); model = model.cuda().train()
    np.random.seed(0)
    out = model({'x': dict(inputs)}, current_epoch=0.0)
    out['full_loss'].backward()
    mid = model.last_intermediates
    res[mode] = (float(out['full_loss']), {k: v.detach().clone() for k, v in mid.items() if torch.is_tensor(v)},
                 {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None})
    print('mode', mode, 'loss', res[mode][0])
def rel(a, b): return float((a.double()-b.double()).abs().max() / (b.double().abs().max()+1e-30))
def l2(a, b): return float((a.double()-b.double()).norm() / (b.double().norm()+1e-30))
for k in ('left_g_initial', 'left_pupil_size', 'PoG_px_initial', 'heatmap_final', 'PoG_px_final'):
    a, b = res[1][1][k], res[0][1][k]
    print(k, 'rel', rel(a, b), 'nan', bool(torch.isnan(a).any()))
worst = sorted(((l2(res[1][2][n], res[0][2][n]), n) for n in res[0][2]), reverse=True)[:12]
for e, n in worst: print('%.3e %s nan=%s' % (e, n, bool(torch.isnan(res[1][2][n]).any())))
