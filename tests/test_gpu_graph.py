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


def test_replay_survives_workspace_growth(cfg):
    """A captured step has the shared scratch buffers' addresses baked in.  A later eager call
    with a bigger batch (the reference evaluates every 500 steps with test_batch_size,
    training.py:548-567) outgrows those buffers: the old ones must stay allocated, otherwise the
    replay writes into memory the allocator has handed to someone else."""
    from eve_b200 import lib as L
    from eve_b200 import synth
    cfg.override('refine_net_enabled', True)
    cfg.override('load_screen_content', True)
    B, T = 1, 2
    xs = [synth.make_clip_batch(B, T, seed=80 + i) for i in range(4)]
    dev = torch.device('cuda', torch.cuda.current_device())
    L._workspaces.clear()      # earlier tests may have left buffers big enough for anything
    ma, ta, ga = _make(cfg, 11, True, xs[0])
    la = [float(ga(xs[1]))]
    before = L.workspace(1, dev).data_ptr()
    # bigger eager evaluation pass: the scratch buffers have to grow
    big = {k: v.cuda() for k, v in synth.make_clip_batch(3, 4, seed=99).items()}
    ma.eval()
    with torch.no_grad():
        out = ma(big)
    ma.train()
    assert np.isfinite(float(out['full_loss']))
    assert L.workspace(1, dev).data_ptr() != before, 'the evaluation pass was meant to outgrow ws'
    # allocate and scribble over fresh memory: a freed workspace would be handed out here
    junk = [torch.full((64 * 1024 * 1024,), float('nan'), device='cuda') for _ in range(4)]
    la += [float(ga(x)) for x in xs[2:]]
    pa = ta.flat.clone()
    del junk
    ga.close()
    mb, tb, gb = _make(cfg, 11, False, xs[0])
    lb = [float(gb(x)) for x in xs[1:]]
    assert la == lb, (la, lb)
    assert torch.equal(pa, tb.flat)
    gb.close()


def test_learning_rate_and_optimizer_state_follow_the_trainer(cfg):
    """The captured step reads the learning rate from a device scalar: a scheduler changing
    trainer.lr between replays takes effect (training.py:382-440,576); state_dict /
    load_state_dict carry the Adam moments and the step count (checkpoint_manager.py:70-72)."""
    from eve_b200 import synth
    cfg.override('refine_net_enabled', True)
    cfg.override('load_screen_content', True)
    B, T = 1, 2
    xs = [synth.make_clip_batch(B, T, seed=90 + i) for i in range(3)]
    ma, ta, ga = _make(cfg, 13, True, xs[0])
    p0 = ta.flat.clone()
    ta.lr = 0.0                      # a zero rate freezes the parameters from the next step on
    ga(xs[1])
    assert torch.equal(ta.flat, p0)
    ta.param_groups[0]['lr'] = 1e-4  # the torch.optim way schedulers write it
    ga(xs[2])
    assert not torch.equal(ta.flat, p0)
    state = ta.state_dict()
    assert state['steps'] == int(ta.step_dev) == ta.steps
    flat_a = ta.flat.clone()
    np.random.seed(5)
    la = float(ga(xs[0]))
    final_a = ta.flat.clone()
    ga.close()
    mb, tb, gb = _make(cfg, 13, False, xs[0])
    tb.flat.copy_(flat_a)
    tb.load_state_dict(state)
    np.random.seed(5)
    lb = float(gb(xs[0]))
    assert la == lb
    assert torch.equal(final_a, tb.flat)
    gb.close()
