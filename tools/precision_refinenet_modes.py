import sys, torch, numpy as np
sys.path.insert(0, '.')
from eve_b200.config import DefaultConfig
from eve_b200 import synth, lib as L
from eve_b200.models import RefineNet
from oracle import eve_oracle as O
lib = L.load()
def run(rnn, cells, skip, screen):
    cfg = DefaultConfig(); cfg.reset()
    cfg.override('refine_net_enabled', True); cfg.override('load_screen_content', screen)
    cfg.override('refine_net_use_skip_connections', skip)
    cfg.override('refine_net_rnn_type', rnn); cfg.override('refine_net_rnn_num_cells', cells)
    sd = synth.make_state_dict(synth.refine_net_param_shapes(cfg), 79)
    B, T = 2, 3
    g = torch.Generator().manual_seed(10)
    px = torch.stack([torch.rand(B, T, generator=g) * 1920, torch.rand(B, T, generator=g) * 1080], -1)
    hm = O.make_heatmaps(px, 10.0)
    scr = torch.rand(B, T, 3, 72, 128, generator=g) if screen else None
    h0 = torch.randn(cells, B, 64, 5, 8, generator=g) * 0.5
    def oracle(dtype):
        osd = {'refine_net.' + k: v.to(dtype) for k, v in sd.items()}
        states = [h0[i].to(dtype) for i in range(cells)]
        outs = []
        for t in range(T):
            o, states = O.refine_net_step(osd, cfg, scr[:, t].to(dtype) if screen else None, hm[:, t].to(dtype), states)
            outs.append(o)
        return torch.stack(outs, 1), torch.stack(states, 0)
    with torch.no_grad():
        w64, f64 = oracle(torch.float64)
        w32, f32 = oracle(torch.float32)
    def rel(a, b): return float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max())
    print(rnn, cells, skip, screen, 'fp32-oracle: out %.2e hT %.2e' % (rel(w32, w64), rel(f32, f64)))
    for mode in (0, 1, 2):
        lib.eve_set_conv_mode(mode)
        net = RefineNet(); net.load_state_dict(sd); net = net.cuda()
        with torch.no_grad():
            got, hT, _ = net.sequence(scr.cuda() if screen else None, hm.cuda(), h0.cuda(), None)
        print('  mode', mode, 'out %.2e hT %.2e' % (rel(got, w64), rel(hT, f64)))
run('CRNN', 2, True, True)
run('CGRU', 2, False, False)
run('CGRU', 1, True, True)
