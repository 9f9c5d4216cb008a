"""Where does the weight-gradient error at the bench size come from?  One optimisation step at
B=8, T=30 (config 2 = static EyeNet, config 3 = EyeNet + GazeRefineNet) evaluated four ways:
CPU oracle in fp64 (the yardstick), CPU oracle in fp32 (= the reference's own arithmetic), this
library in conv mode 0 (fp32 CUDA cores) and in conv mode 1 (tcgen05 split operands, the product).
Prints, per parameter group, the L2 distance of each fp32 evaluation from the fp64 one.
Usage: python tools/grad_precision.py [2|3] [B] [T]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from eve_b200 import lib as L, synth                    # noqa: E402
from eve_b200.config import DefaultConfig               # noqa: E402
from eve_b200.models import EVE                         # noqa: E402
from oracle import eve_oracle as O                      # noqa: E402  (checker only)

which = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
T = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cfg = DefaultConfig()
cfg.reset()
refine = which == 3
cfg.override('refine_net_enabled', refine)
cfg.override('load_screen_content', refine)
if not refine:
    cfg.override('eye_net_use_rnn', False)
seed = 3 if refine else 4
sd = synth.make_state_dict(synth.eye_net_param_shapes(cfg), seed, 'eye_net.')
if refine:
    sd.update(synth.make_state_dict(synth.refine_net_param_shapes(cfg), seed + 1000, 'refine_net.'))
inputs = synth.make_clip_batch(B, T, seed=seed, with_screen=refine)
lib = L.load()


def gpu(mode):
    lib.eve_set_conv_mode(mode)
    model = EVE(output_predictions=True)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    np.random.seed(seed)
    out = model({'bench': {k: v.cuda() for k, v in inputs.items()}}, current_epoch=0.0)
    out['full_loss'].backward()
    torch.cuda.synchronize()
    g = {k: p.grad.detach().double().cpu() for k, p in model.named_parameters() if p.grad is not None}
    o = {k: v.detach().double().cpu() for k, v in out.items() if torch.is_tensor(v)}
    del model, out
    torch.cuda.empty_cache()
    return o, g


def oracle(dtype):
    np.random.seed(seed)
    std = np.radians(cfg.refine_net_offset_augmentation_sigma)
    kap = {'left': torch.from_numpy(np.random.normal(size=(B, 2), scale=std).astype(np.float32)).to(dtype),
           'right': torch.from_numpy(np.random.normal(size=(B, 2), scale=std).astype(np.float32)).to(dtype)}
    osd = {k: v.clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
    inp = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in inputs.items()}
    t0 = time.time()
    want, mid = O.eve_forward(osd, cfg, inp, True, kap)
    want['full_loss'].backward()
    print('oracle %s: %.1f s' % (dtype, time.time() - t0), flush=True)
    o = dict(mid)
    o.update(want)
    return ({k: v.detach().double() for k, v in o.items() if torch.is_tensor(v)},
            {k: v.grad.double() for k, v in osd.items() if v.grad is not None})


torch.set_num_threads(max(torch.get_num_threads(), 8))
o1, g1 = gpu(1)
o0, g0 = gpu(0)
lib.eve_set_conv_mode(1)
o32, g32 = oracle(torch.float32)
o64, g64 = oracle(torch.float64)


def l2(a, b):
    return float((a - b).norm() / (b.norm() + 1e-300))


top = max(float(v.norm()) for v in g64.values())
print('%-78s %9s %9s %9s %9s' % ('parameter', '|g|/top', 'cpu fp32', 'mode 0', 'mode 1'))
rows = []
for k in g64:
    if float(g64[k].norm()) < 1e-5 * top:
        continue
    rows.append((k, float(g64[k].norm()) / top, l2(g32[k], g64[k]), l2(g0[k], g64[k]), l2(g1[k], g64[k])))
for r in rows:
    print('%-78s %9.2e %9.2e %9.2e %9.2e' % r)
for name, sel in (('eye_net', [r for r in rows if r[0].startswith('eye_net.')]),
                  ('refine_net', [r for r in rows if r[0].startswith('refine_net.')])):
    if not sel:
        continue
    a = np.array([[r[2], r[3], r[4]] for r in sel])
    print('%s: median  cpu-fp32 %.2e  mode0 %.2e  mode1 %.2e   max  %.2e  %.2e  %.2e' % (
        (name,) + tuple(np.median(a, 0)) + tuple(a.max(0))))
for k in ('g_initial', 'g_final', 'PoG_px_initial', 'PoG_px_final', 'full_loss'):
    if k in o64 and k in o1:
        den = float(o64[k].abs().max()) + 1e-300
        print('%-16s max-rel vs fp64: cpu fp32 %.2e  mode0 %.2e  mode1 %.2e' % (
            k, float((o32[k] - o64[k]).abs().max()) / den, float((o0[k] - o64[k]).abs().max()) / den,
            float((o1[k] - o64[k]).abs().max()) / den))
