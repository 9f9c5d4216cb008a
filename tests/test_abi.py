"""The C-ABI library loads and exports every symbol include/eve_b200.h declares; host-side
entry points (size queries, weight names, argument validation) behave.  No GPU needed."""
import ctypes as C
import os
import re

import pytest

from eve_b200 import lib as L

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include',
                      'eve_b200.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(eve_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = L.load()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), 'missing export: ' + n
        assert n in L.SIGNATURES, 'no ctypes signature for ' + n
    assert sorted(L.SIGNATURES) == names
    assert lib.eve_version() >= 100


def test_size_queries_scale_with_batch():
    lib = L.load()
    p1 = L.EyeNetCnnParams(2, 128, 128, 128)
    p2 = L.EyeNetCnnParams(4, 128, 128, 128)
    s1, s2 = lib.eve_eyenet_cnn_saved_bytes(C.byref(p1)), lib.eve_eyenet_cnn_saved_bytes(C.byref(p2))
    assert 0 < s1 < s2 <= 2 * s1 + 4096 * 64
    assert lib.eve_eyenet_cnn_workspace_bytes(C.byref(p1)) > 0
    # ~6.5 MB of fp32 activations per eye patch
    assert 4e6 < s2 / 4 < 12e6
    r = L.RefineNetParams(2, 3, 4, 1, 3, 1, 64)
    assert lib.eve_refinenet_saved_bytes(C.byref(r)) > 6 * 20e6
    assert lib.eve_refinenet_workspace_bytes(C.byref(r)) > 0
    t = L.EyeNetTailParams(4, 3, 128, 1, 3, 1)
    assert lib.eve_eyenet_tail_saved_bytes(C.byref(t)) > 0
    assert lib.eve_eyenet_tail_num_weights(C.byref(t)) == 15
    t = L.EyeNetTailParams(4, 3, 128, 1, 0, 1)
    assert lib.eve_eyenet_tail_num_weights(C.byref(t)) == 13


@pytest.mark.parametrize('screen,skip,rnn,cells', [(True, True, 'CGRU', 1), (False, False, 'CRNN', 1),
                                                   (True, True, 'CLSTM', 2), (True, True, None, 1)])
def test_refinenet_weight_names_match_state_dict(cfg, screen, skip, rnn, cells):
    from eve_b200 import ops, synth
    cfg.override('refine_net_enabled', True)
    cfg.override('load_screen_content', screen)
    cfg.override('refine_net_use_skip_connections', skip)
    cfg.override('refine_net_use_rnn', rnn is not None)
    if rnn:
        cfg.override('refine_net_rnn_type', rnn)
    cfg.override('refine_net_rnn_num_cells', cells)
    p = L.RefineNetParams(1, 1, 4 if screen else 1, int(skip), L.REFINE_RNN_TYPES[rnn], cells, 64)
    names = ops.refinenet_weight_names(p)
    assert sorted(names) == sorted(synth.refine_net_param_shapes(cfg))
    assert len(set(names)) == len(names)


def test_bad_configuration_is_reported_not_ignored():
    lib = L.load()
    bad = L.RefineNetParams(1, 1, 3, 1, 3, 1, 64)       # in_channels must be 4 or 1
    assert lib.eve_refinenet_num_weights(C.byref(bad)) == -1
    assert b'in_channels' in lib.eve_last_error()
    bad = L.EyeNetTailParams(1, 1, 128, 1, 7, 1)
    assert lib.eve_eyenet_tail_num_weights(C.byref(bad)) == -1
    assert b'Unknown RNN type for EyeNet' in lib.eve_last_error()
    c = L.ConvParams(1, 8, 8, 4, 4, 3, 0, 1)            # stride 0
    assert lib.eve_conv2d_workspace_bytes(C.byref(c)) == 0
    # NULL pointers are refused before anything is launched
    ok = L.ConvParams(1, 8, 8, 4, 4, 3, 1, 1)
    rc = lib.eve_conv2d_fwd(C.byref(ok), None, None, None, None, None, 0, None)
    assert rc == 5 and b'NULL' in lib.eve_last_error()


def test_models_keep_reference_state_dict_keys(cfg):
    from eve_b200 import synth
    from eve_b200.models import EVE
    cfg.override('refine_net_enabled', True)
    cfg.override('load_screen_content', True)
    m = EVE()
    want = {'eye_net.' + k: v for k, v in synth.eye_net_param_shapes(cfg).items()}
    want.update({'refine_net.' + k: v for k, v in synth.refine_net_param_shapes(cfg).items()})
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}
    assert sum(p.numel() for p in m.eye_net.parameters()) == 11398337
    assert sum(p.numel() for p in m.refine_net.parameters()) == 5266289
    # the layers the reference zero-initialises (eye_net.py:96, refine_net.py:235)
    assert float(m.state_dict()['eye_net.fc_to_gaze.2.weight'].abs().max()) == 0.0
    assert float(m.state_dict()['refine_net.final.2.weight'].abs().max()) == 0.0


def test_cpu_tensors_are_refused(cfg):
    import torch
    from eve_b200.models import EyeNet
    net = EyeNet()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        net.cnn_features(torch.zeros(1, 3, 128, 128))


def test_integration_shim_routes_reference_imports(tmp_path, monkeypatch):
    """The three-line `src/models/__init__.py` of INTEGRATION.md makes the reference's own
    import statements (`from models.eve import EVE`, `from models.common import ...`) resolve
    to this package."""
    import importlib
    import sys
    text = open(os.path.join(os.path.dirname(HEADER), '..', 'INTEGRATION.md')).read()
    start = text.index('# src/models/__init__.py')
    shim = text[start:text.index('```', start)]
    pkg = tmp_path / 'models'
    pkg.mkdir()
    (pkg / '__init__.py').write_text(shim)
    monkeypatch.syspath_prepend(str(tmp_path))
    for k in [k for k in sys.modules if k == 'models' or k.startswith('models.')]:
        monkeypatch.delitem(sys.modules, k)
    importlib.invalidate_caches()
    from models.eve import EVE as RefPathEVE                      # noqa: E402
    from models.common import pitchyaw_to_vector, soft_argmax     # noqa: E402,F401
    from models.eye_net import EyeNet as RefPathEyeNet            # noqa: E402
    import eve_b200.models as M
    assert RefPathEVE is M.EVE and RefPathEyeNet is M.EyeNet
    for k in [k for k in sys.modules if k == 'models' or k.startswith('models.')]:
        monkeypatch.delitem(sys.modules, k)


def test_tuning_options_round_trip_without_a_gpu():
    """eve_set_option / eve_get_option are plain host state: defaults, range checks and unknown
    names behave as include/eve_b200.h documents (no CUDA call involved)."""
    from eve_b200 import lib as L
    lib = L.load()
    for name, lo, hi in [('tc_stage_cap', 2, 24), ('tc_row_kernel', 0, 1), ('tc_row_strips', 0, 128),
                         ('tc_row_wgrad', 0, 2), ('tc_wgrad_waves', 1, 8), ('fused_planes', 0, 1)]:
        prev = L.get_option(name)
        assert lo <= prev <= hi
        L.set_option(name, lo)
        assert L.get_option(name) == lo
        assert lib.eve_set_option(name.encode(), hi + 1) != 0
        assert L.get_option(name) == lo
        L.set_option(name, prev)
    assert lib.eve_set_option(b'bogus', 0) != 0
    assert 'unknown option' in L.last_error()
