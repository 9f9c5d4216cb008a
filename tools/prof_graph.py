"""Timeline of ONE graph-replayed training step (B=8, T=30): per-kernel busy time AND the idle gap
in front of each kernel, grouped by kernel name.  Usage: python tools/prof_graph.py [refine|static]"""
import collections
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from torch.profiler import profile, ProfilerActivity   # noqa: E402
import bench                                            # noqa: E402
from eve_b200 import lib as L                           # noqa: E402
from eve_b200.graph import GraphedTrainStep             # noqa: E402
from eve_b200.models import EVE                         # noqa: E402
from eve_b200.parallel import FlatAdamTrainer           # noqa: E402

wl = 'eve_refine' if (len(sys.argv) < 2 or sys.argv[1] == 'refine') else 'eyenet_static'
L.load()
cfg = bench.configure(wl)
dev = torch.device('cuda', 0)
np.random.seed(0)
model = EVE()
model.load_state_dict(bench.build_state_dict(cfg), strict=True)
model = model.to(dev).train()
tr = FlatAdamTrainer(model)
batch = {k: v.to(dev) for k, v in bench.make_batch(8, 30, cfg, seed=0, pinned=False).items()}
step = GraphedTrainStep(model, tr, batch, warmup=3, tag='bench')
for _ in range(3):
    step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(batch)
    torch.cuda.synchronize()
ks = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks.sort(key=lambda e: e.time_range.start)
busy = collections.defaultdict(float)
gap = collections.defaultdict(float)
cnt = collections.Counter()
prev_end = None
prev_name = None
pairs = collections.defaultdict(lambda: [0, 0.0])
for e in ks:
    name = e.name.replace('(anonymous namespace)::', '').replace('void ', '').split('(')[0][:70]
    d = e.time_range.end - e.time_range.start
    busy[name] += d
    cnt[name] += 1
    if prev_end is not None:
        gp = max(0.0, e.time_range.start - prev_end)
        gap[name] += gp
        pr = pairs[(prev_name, name)]
        pr[0] += 1
        pr[1] += gp
    prev_end = max(prev_end or 0, e.time_range.end)
    prev_name = name
span = ks[-1].time_range.end - ks[0].time_range.start
tb, tg = sum(busy.values()), sum(gap.values())
print('kernels %d  span %.3f ms  busy %.3f ms  gaps %.3f ms' % (len(ks), span / 1e3, tb / 1e3, tg / 1e3))
print('%-72s %6s %9s %9s %8s' % ('kernel', 'count', 'busy ms', 'gap ms', 'gap/k us'))
for name in sorted(busy, key=lambda n: -(busy[n] + gap[n]))[:45]:
    print('%-72s %6d %9.3f %9.3f %8.2f' % (name, cnt[name], busy[name] / 1e3, gap[name] / 1e3,
                                            gap[name] / cnt[name]))
aten = [n for n in busy if not n.startswith('eve::') and not n.startswith('Memcpy') and not n.startswith('Memset')]
print('ATen kernels in the step: %d launches, %.3f ms busy; library kernels: %d launches' % (
    sum(cnt[n] for n in aten), sum(busy[n] for n in aten) / 1e3, sum(cnt[n] for n in busy if n not in aten)))
for n in sorted(aten, key=lambda n: -cnt[n])[:12]:
    print('   %5d x %s' % (cnt[n], n[:100]))
print()
print('largest idle gaps by (previous kernel -> next kernel)')
for (a, b), (c, gsum) in sorted(pairs.items(), key=lambda kv: -kv[1][1])[:25]:
    print('%9.3f ms %5d x %7.1f us   %s -> %s' % (gsum / 1e3, c, gsum / c, a[:48], b[:48]))
