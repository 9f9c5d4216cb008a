import sys, torch, numpy as np
sys.path.insert(0, '.')
from torch.profiler import profile, ProfilerActivity
from eve_b200.config import DefaultConfig
from eve_b200 import synth, lib as L
from eve_b200.models import EVE
from eve_b200.parallel import FlatAdamTrainer
lib = L.load()
cfg = DefaultConfig(); cfg.reset()
wl = sys.argv[1] if len(sys.argv) > 1 else 'refine'
if wl == 'refine':
    cfg.override('refine_net_enabled', True); cfg.override('load_screen_content', True)
B, T = 8, 30
sd = synth.make_state_dict(synth.eye_net_param_shapes(cfg), 0, 'eye_net.')
if cfg.refine_net_enabled:
    sd.update(synth.make_state_dict(synth.refine_net_param_shapes(cfg), 1000, 'refine_net.'))
batch = {k: v.cuda() for k, v in synth.make_clip_batch(B, T, seed=0, with_screen=bool(cfg.load_screen_content)).items()}
model = EVE(); model.load_state_dict(sd); model = model.cuda().train()
tr = FlatAdamTrainer(model)
np.random.seed(0)
def step():
    out = model({'x': dict(batch)}, current_epoch=0.0)
    tr.step(out['full_loss'])
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print('total device ms %.2f' % tot)
for k, c, t in rows[:45]:
    print('%8.3f ms %5.1f%% %5d  %s' % (t, 100 * t / tot, c, k[:110]))
import time
t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('host enqueue ms %.1f, total %.1f' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
