"""Shared helpers for the parity tests (test infrastructure)."""
import ast
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
GRAD_STRIDE = 997          # must match oracle/gen_golden.py


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith('.npz') and f != 'modules.npz')


def apply_case_config(cfg, gold):
    cfg.reset()
    for k, v in ast.literal_eval(str(gold['meta/overrides'])):
        cfg.override(k, v)
    return cfg


def case_inputs(gold, cfg):
    from eve_b200 import synth
    return synth.make_clip_batch(int(gold['meta/B']), int(gold['meta/T']),
                                 seed=int(gold['meta/seed']),
                                 with_screen=bool(cfg.load_screen_content),
                                 pad_last=int(gold['meta/pad_last']))


def case_state_dict(gold, cfg):
    from eve_b200 import synth
    seed = int(gold['meta/seed'])
    sd = synth.make_state_dict(synth.eye_net_param_shapes(cfg), seed, 'eye_net.')
    if cfg.refine_net_enabled:
        sd.update(synth.make_state_dict(synth.refine_net_param_shapes(cfg), seed + 1000,
                                        'refine_net.'))
    return sd


def case_kappas(gold, B):
    """The kappas the reference drew: np.random.seed(seed) then left, right (eve.py:468-469)."""
    seed = int(gold['meta/seed'])
    np.random.seed(seed)
    std = np.radians(3.0)
    left = np.random.normal(size=(B, 2), loc=0.0, scale=std)
    right = np.random.normal(size=(B, 2), loc=0.0, scale=std)
    return {'left': torch.from_numpy(left.astype(np.float32)),
            'right': torch.from_numpy(right.astype(np.float32))}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))
