import sys, torch
sys.path.insert(0, '.')
import torch.nn.functional as F
from tests import gpu_util as G
for shape, affine, act in [((3, 64, 64, 64), False, 1), ((3, 64, 32, 32), False, 1), ((3, 128, 16, 16), False, 1),
                           ((3, 256, 8, 8), False, 1), ((3, 512, 4, 4), False, 1), ((3, 512, 4, 4), False, 0),
                           ((6, 16, 72, 128), True, 2), ((6, 64, 5, 8), True, 1)]:
    g = torch.Generator().manual_seed(1)
    n, c, h, w = shape
    x = torch.randn(shape, generator=g) * 0.7 + 0.3
    x = F.relu(x) if act == 1 else x           # sparse, like real activations
    gamma = (1 + 0.1 * torch.randn(c, generator=g)) if affine else None
    beta = 0.1 * torch.randn(c, generator=g) if affine else None
    xd = x.double().requires_grad_(True)
    y = F.instance_norm(xd, weight=gamma.double() if affine else None, bias=beta.double() if affine else None, eps=1e-5)
    y = {0: lambda t: t, 1: F.relu, 2: lambda t: F.leaky_relu(t, 0.01)}[act](y)
    dy = torch.randn(shape, generator=g)
    y.backward(dy.double())
    cg, cb = (gamma.cuda(), beta.cuda()) if affine else (None, None)
    got, mean, rstd = G.instnorm_fwd(x.cuda(), cg, cb, act)
    dx, dgm, dbt = G.instnorm_bwd(dy.cuda(), got, x.cuda(), mean, rstd, cg, act)
    dx2, _, _ = G.instnorm_bwd(dy.cuda(), got, x.cuda(), mean, rstd, cg, act)
    var = x.double().var(dim=(2, 3), unbiased=False)
    print(shape, act, 'fwd %.1e mean %.1e rstd %.1e dx %.1e deterministic %s' % (
        G.rel(got, y), G.rel(mean, x.double().mean(dim=(2, 3))), G.rel(rstd, torch.rsqrt(var + 1e-5)),
        G.rel(dx, xd.grad), bool(torch.equal(dx, dx2))))
