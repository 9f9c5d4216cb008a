"""Per-geometry table of the convolution passes of ONE eager training step at the bench size
(B=8, T=30): launches, kernel ms (CUDA events on the launching stream, library profiler) and
algorithmic TFLOP/s per (pass, N, H, W, Cin, Cout, k, stride).
Usage: python tools/conv_table.py [refine|static] [steps]"""
import collections
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import bench                                            # noqa: E402
from eve_b200 import lib as L                           # noqa: E402
from eve_b200.models import EVE                         # noqa: E402
from eve_b200.parallel import FlatAdamTrainer           # noqa: E402

wl = 'eve_refine' if (len(sys.argv) < 2 or sys.argv[1] == 'refine') else 'eyenet_static'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = L.load()
cfg = bench.configure(wl)
dev = torch.device('cuda', 0)
np.random.seed(0)
model = EVE()
model.load_state_dict(bench.build_state_dict(cfg), strict=True)
model = model.to(dev).train()
tr = FlatAdamTrainer(model)
batch = {k: v.to(dev) for k, v in bench.make_batch(8, 30, cfg, seed=0, pinned=False).items()}


def one():
    out = model({'bench': dict(batch)}, current_epoch=0.0)
    tr.step(out['full_loss'])


for _ in range(2):
    one()
torch.cuda.synchronize()
lib.eve_profile_reset()
lib.eve_profile_enable(1)
for _ in range(steps):
    one()
torch.cuda.synchronize()
lib.eve_profile_enable(0)
need = lib.eve_profile_dump(None, 0)
buf = C.create_string_buffer(need + 16)
lib.eve_profile_dump(buf, need + 16)
agg = collections.OrderedDict()
for line in buf.value.decode().splitlines():
    f = line.split()
    key = tuple(int(v) for v in f[:8])
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += float(f[8])
    a[2] += float(f[9])
names = {0: 'fwd', 1: 'dgrad', 2: 'wgrad'}
tot = sum(a[1] for a in agg.values()) / steps
print('conv kernel time %.3f ms/step over %d geometries' % (tot, len(agg)))
print('%-6s %5s %4s %4s %4s %4s %2s %2s %6s %9s %8s %6s' % ('pass', 'N', 'H', 'W', 'Cin', 'Cout', 'k', 's',
                                                          'n/step', 'ms/step', 'TFLOP/s', 'share'))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    kind, n, h, w, ci, co, k, s = key
    print('%-6s %5d %4d %4d %4d %4d %2d %2d %6.1f %9.4f %8.1f %5.1f%%' % (
        names[kind], n, h, w, ci, co, k, s, a[0] / steps, a[1] / steps, a[2] / a[1] * 1e-9,
        100 * a[1] / steps / tot))
