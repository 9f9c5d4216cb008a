import sys, torch, numpy as np
sys.path.insert(0, '.')
from eve_b200.config import DefaultConfig
from eve_b200 import synth, lib as L
from eve_b200.models import EVE
from eve_b200.parallel import FlatAdamTrainer
lib = L.load()
cfg = DefaultConfig(); cfg.reset()
cfg.override('refine_net_enabled', True); cfg.override('load_screen_content', True)
B, T, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
lib.eve_set_conv_mode(mode)
sd = synth.make_state_dict(synth.eye_net_param_shapes(cfg), 0, 'eye_net.')
sd.update(synth.make_state_dict(synth.refine_net_param_shapes(cfg), 1000, 'refine_net.'))
batches = [{k: v.cuda() for k, v in synth.make_clip_batch(B, T, seed=i).items()} for i in range(2)]
model = EVE(); model.load_state_dict(sd); model = model.cuda().train()
tr = FlatAdamTrainer(model)
np.random.seed(1234)
for i in range(10):
    out = model({'x': dict(batches[i % 2])}, current_epoch=0.0)
    loss = out['full_loss']
    tr.step(loss)
    bad = [k for k, v in model.last_intermediates.items() if torch.is_tensor(v) and v.is_floating_point() and not bool(torch.isfinite(v).all())]
    print(i, float(loss.detach()), 'gnorm', float(tr.last_grad_norm), 'pmax', float(tr.flat.abs().max()), 'nonfinite', bad[:6], {k: float(v) for k, v in out.items() if k.startswith('loss_') and not bool(torch.isfinite(v))})
