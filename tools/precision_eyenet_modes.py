import sys, torch, numpy as np
sys.path.insert(0, '.')
from eve_b200.config import DefaultConfig
from eve_b200 import synth, lib as L
from eve_b200.models import EyeNet
lib = L.load()
cfg = DefaultConfig(); cfg.reset()
N = int(sys.argv[1])
sd = synth.make_state_dict(synth.eye_net_param_shapes(cfg), 78)
g = torch.Generator().manual_seed(9)
x = (torch.rand(N, 3, 128, 128, generator=g) * 2 - 1).cuda()
wf = torch.randn(N, 128, generator=g).cuda()
res = {}
for mode in (0, 1):
    lib.eve_set_conv_mode(mode)
    net = EyeNet(); net.load_state_dict(sd); net = net.cuda()
    got = net.cnn_features(x)
    (got * wf).sum().backward()
    res[mode] = (got.detach(), {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None})
def l2(a, b): return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
print('feat', l2(res[1][0], res[0][0]))
for n in res[0][1]:
    print('%-45s %.3e' % (n, l2(res[1][1][n], res[0][1][n])))
