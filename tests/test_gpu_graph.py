"""The CUDA-graph replay of a training step is the eager step, bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(cfg, seed, capture, x0):
    from eve_b200 import synth
    from eve_b200.graph import GraphedTrainStep
    from eve_b200.models import EVE
    from eve_b200.parallel import FlatAdamTrainer
    sd = synth.make_state_dict(synth.eye_net_param_shapes(cfg), seed, 'eye_net.')
    sd.update(synth.make_state_dict(synth.refine_net_param_shapes(cfg), seed + 1000, 'refine_net.'))
    model = EVE()
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    trainer = FlatAdamTrainer(model, lr=1e-4)
    np.random.seed(123)
    return model, trainer, GraphedTrainStep(model, trainer, x0, warmup=2, capture=capture)


def test_graph_replay_equals_eager_steps(cfg):
    from eve_b200 import synth
    cfg.override('refine_net_enabled', True)
    cfg.override('load_screen_content', True)
    B, T = 2, 3
    xs = [synth.make_clip_batch(B, T, seed=40 + i) for i in range(4)]
    ma, ta, ga = _make(cfg, 7, True, xs[0])
    la = [float(ga(x)) for x in xs[1:]]
    pa = ta.flat.clone()
    step_a = int(ta.step_dev)
    mb, tb, gb = _make(cfg, 7, False, xs[0])
    lb = [float(gb(x)) for x in xs[1:]]
    assert la == lb, (la, lb)
    assert torch.equal(pa, tb.flat)
    assert step_a == int(tb.step_dev) == 2 + 3      # warm-up steps + replays; capture runs nothing
    assert all(np.isfinite(la))
    ga.close()
    gb.close()


def test_prefetched_inputs_give_the_same_steps(cfg):
    """GraphedTrainStep.prefetch() (next batch staged H2D on a second stream, consumed by
    __call__(None)) is only a different route for the same bytes: losses and parameters equal the
    direct path's bit for bit."""
    from eve_b200 import synth
    cfg.override('refine_net_enabled', True)
    cfg.override('load_screen_content', True)
    B, T = 2, 3
    xs = [{k: v.pin_memory() for k, v in synth.make_clip_batch(B, T, seed=60 + i).items()}
          for i in range(4)]
    ma, ta, ga = _make(cfg, 9, True, xs[0])
    la = [float(ga(x)) for x in xs[1:]]
    pa = ta.flat.clone()
    mb, tb, gb = _make(cfg, 9, True, xs[0])
    lb = []
    gb.prefetch(xs[1])
    for i in (1, 2, 3):
        loss = gb(None)
        if i < 3:
            gb.prefetch(xs[i + 1])
        lb.append(float(loss))
    assert la == lb, (la, lb)
    assert torch.equal(pa, tb.flat)
    ga.close()
    gb.close()
