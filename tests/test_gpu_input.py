"""Device side of the input pipeline: uint8 frames -> float32 NCHW patches,
bit-exact against the numpy restatement of datasources/eve_sequences.py:196-211,283-299."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from eve_b200 import input_pipeline as IP   # noqa: E402
from oracle import eve_oracle as O          # noqa: E402   (checker only)


def test_preprocess_is_bit_exact_for_every_byte_value():
    # all 256 byte values in every channel / position class
    base = np.arange(256, dtype=np.uint8)
    frames = np.resize(base, (2, 3, 8, 32, 3)).copy()           # [B,T,H,W,C]
    frames[1] = frames[1][..., ::-1, :]
    got = IP.preprocess_frames(torch.from_numpy(frames).cuda()).cpu().numpy()
    want = O.preprocess_frames(frames.reshape(6, 8, 32, 3)).reshape(2, 3, 3, 8, 32)
    assert got.dtype == np.float32 and np.array_equal(got, want)
    got = IP.preprocess_screen_frames(torch.from_numpy(frames).cuda()).cpu().numpy()
    want = O.preprocess_screen_frames(frames.reshape(6, 8, 32, 3)).reshape(2, 3, 3, 8, 32)
    assert np.array_equal(got, want)


def test_eye_patch_split_and_padded_clips():
    rs = np.random.RandomState(0)
    B, T, H, ew = 3, 5, 128, 128
    frames = rs.randint(0, 256, size=(B, T, H, 2 * ew, 3)).astype(np.uint8)
    lens = torch.tensor([5, 3, 0], dtype=torch.int32)            # ragged clips, one empty
    left, right = IP.preprocess_eye_frames(torch.from_numpy(frames).cuda(), lens)
    full = O.preprocess_frames(frames.reshape(B * T, H, 2 * ew, 3))
    wl, wr = O.split_eye_patches(full, ew)
    wl, wr = wl.reshape(B, T, 3, H, ew).copy(), wr.reshape(B, T, 3, H, ew).copy()
    for b in range(B):                                           # zero padding AFTER preprocessing
        wl[b, int(lens[b]):] = 0.0
        wr[b, int(lens[b]):] = 0.0
    assert np.array_equal(left.cpu().numpy(), wl)
    assert np.array_equal(right.cpu().numpy(), wr)
    with pytest.raises(TypeError):
        IP.preprocess_frames(torch.zeros(1, 4, 4, 3).cuda())
    with pytest.raises(RuntimeError):
        IP.preprocess_frames(torch.zeros(1, 4, 4, 3, dtype=torch.uint8))   # CPU tensor: no fallback


def test_graphed_step_accepts_uint8_frames(cfg):
    """The same step from float patches and from the uint8 frames they were made of."""
    from eve_b200 import synth
    from eve_b200.graph import GraphedTrainStep
    from eve_b200.models import EVE
    from eve_b200.parallel import FlatAdamTrainer
    cfg.override('refine_net_enabled', True)
    cfg.override('load_screen_content', True)
    B, T = 2, 3
    rs = np.random.RandomState(3)
    eyes = torch.from_numpy(rs.randint(0, 256, size=(B, T, 128, 256, 3)).astype(np.uint8))
    screen = torch.from_numpy(rs.randint(0, 256, size=(B, T, 72, 128, 3)).astype(np.uint8))
    base = synth.make_clip_batch(B, T, seed=3, with_screen=True)
    left, right = IP.preprocess_eye_frames(eyes.cuda())
    floats = dict(base)
    floats.update(left_eye_patch=left.cpu(), right_eye_patch=right.cpu(),
                  screen_frame=IP.preprocess_screen_frames(screen.cuda()).cpu())
    raw = {k: v for k, v in base.items()
           if k not in ('left_eye_patch', 'right_eye_patch', 'screen_frame')}
    raw.update(eyes_frames=eyes.pin_memory(), screen_frames=screen.pin_memory())
    losses = []
    for batch in (floats, raw):
        sd = synth.make_state_dict(synth.eye_net_param_shapes(cfg), 3, 'eye_net.')
        sd.update(synth.make_state_dict(synth.refine_net_param_shapes(cfg), 1003, 'refine_net.'))
        model = EVE()
        model.load_state_dict(sd, strict=True)
        model = model.cuda().train()
        tr = FlatAdamTrainer(model)
        np.random.seed(11)
        step = GraphedTrainStep(model, tr, {k: v.cuda() for k, v in floats.items()}, warmup=1,
                                tag='t', capture=False)
        np.random.seed(12)
        losses.append(float(step(batch)))
        step.prefetch(batch)
        losses.append(float(step(None)))
        step.close()
    assert losses[0] == losses[2] and losses[1] == losses[3], losses
