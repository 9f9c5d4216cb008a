"""BASELINE config 5: inference-only RefineNet stream, seq_len=900 (30 s @ 30 fps), batch=1,
frames/sec on one B200 (ConvGRU state carried across all 900 steps), next to the CPU oracle on a
bounded sample.  Usage: python tools/bench_stream.py [--cpu]"""
import sys
import time

import torch

sys.path.insert(0, '.')
from eve_b200 import synth                      # noqa: E402
from eve_b200.config import DefaultConfig       # noqa: E402
from eve_b200.models import RefineNet           # noqa: E402
from eve_b200.models.common import batch_make_heatmaps, soft_argmax   # noqa: E402

cfg = DefaultConfig()
cfg.reset()
cfg.override('refine_net_enabled', True)
cfg.override('load_screen_content', True)
T = 900
sd = synth.make_state_dict(synth.refine_net_param_shapes(cfg), 1000)
net = RefineNet()
net.load_state_dict(sd)
net = net.cuda().eval()
g = torch.Generator().manual_seed(0)
px = torch.stack([torch.rand(1, T, generator=g) * 1920, torch.rand(1, T, generator=g) * 1080], -1).cuda()
screen = torch.rand(1, T, 3, 72, 128, generator=g).cuda()


def run():
    with torch.no_grad():
        hm = batch_make_heatmaps(px, cfg.gaze_heatmap_sigma_initial)
        out, hT, _ = net.sequence(screen, hm)
        return soft_argmax(out.reshape(T, 1, 72, 128))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    pog = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print('B200: %d-frame stream in %.2f ms -> %.0f screen frames/s (heatmap raster + RefineNet + soft-argmax)'
      % (T, ms, T / ms * 1e3))
if '--cpu' in sys.argv:
    from oracle import eve_oracle as O
    n = 30
    osd = {'refine_net.' + k: v for k, v in sd.items()}
    with torch.no_grad():
        hm = O.make_heatmaps(px[:, :n].cpu(), cfg.gaze_heatmap_sigma_initial)
        O.refine_net_sequence(osd, cfg, screen[:, :4].cpu(), hm[:, :4])
        t0 = time.perf_counter()
        O.soft_argmax(O.refine_net_sequence(osd, cfg, screen[:, :n].cpu(), hm).reshape(n, 1, 72, 128))
        dt = time.perf_counter() - t0
    print('CPU oracle (%d threads): %d frames in %.2f s -> %.1f frames/s' % (torch.get_num_threads(), n, dt, n / dt))
