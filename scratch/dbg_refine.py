import sys, torch, numpy as np
sys.path.insert(0, '.')
from eve_b200.config import DefaultConfig
from eve_b200 import synth
from eve_b200.models import RefineNet
from oracle import eve_oracle as O
cfg = DefaultConfig(); cfg.reset()
cfg.override('refine_net_enabled', True); cfg.override('load_screen_content', True)
sd = synth.make_state_dict(synth.refine_net_param_shapes(cfg), 79)
net = RefineNet(); net.load_state_dict(sd); net = net.cuda()
B, T = 2, 3
g = torch.Generator().manual_seed(10)
px = torch.stack([torch.rand(B, T, generator=g) * 1920, torch.rand(B, T, generator=g) * 1080], -1)
hm = O.make_heatmaps(px, 10.0)
scr = torch.rand(B, T, 3, 72, 128, generator=g)
wo = torch.randn(B, T, 1, 72, 128, generator=g)
def oracle(dtype):
    osd = {'refine_net.' + k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
    hmo = hm.detach().clone().to(dtype).requires_grad_(True)
    states = None; outs = []
    for t in range(T):
        o, states = O.refine_net_step(osd, cfg, scr[:, t].to(dtype), hmo[:, t], states or None)
        outs.append(o)
    want = torch.stack(outs, 1)
    (want * wo.to(dtype)).sum().backward()
    return want.detach(), hmo.grad, {k: v.grad for k, v in osd.items()}
w64, dh64, g64 = oracle(torch.float64)
w32, dh32, g32 = oracle(torch.float32)
def l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
hmc = hm.detach().cuda().requires_grad_(True)
got, hT, cT = net.sequence(scr.cuda(), hmc, None, None)
(got * wo.cuda()).sum().backward()
print('fwd', l2(got, w64), l2(w32, w64))
print('dhm', l2(hmc.grad, dh64), l2(dh32, dh64))
for name, p in net.named_parameters():
    ref = g64['refine_net.' + name]
    print('%-70s %.2e %.2e  norm %.3e' % (name, l2(p.grad, ref), l2(g32['refine_net.' + name], ref), float(ref.norm())))
