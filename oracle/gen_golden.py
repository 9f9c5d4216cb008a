"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (``/root/reference`` is not present on the GPU box):

    python oracle/gen_golden.py [--out tests/golden]

The reference is imported from where it lies (never copied); the logging / data-loading
dependencies that are absent from this image are replaced by empty stub modules exactly
as SURVEY.md appendix A documents.  Inputs and weights come from eve_b200/synth.py
(numpy RandomState => reproducible from the seed stored in each fixture), are loaded
into the reference modules with ``load_state_dict(strict=True)`` (which also pins the
parameter names/shapes), and the reference's outputs, captured intermediates and
gradients are written out.  Large gradients are stored as (norm, sum, strided sample).
"""
import argparse
import os
import sys
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = '/root/reference/src'
GRAD_STRIDE = 997


def import_reference():
    for name in ['gspread', 'oauth2client', 'oauth2client.service_account', 'tensorboardX',
                 'coloredlogs', 'h5py', 'ffmpeg']:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['oauth2client.service_account'].ServiceAccountCredentials = object
    sys.modules['tensorboardX'].SummaryWriter = object
    sys.argv[0] = os.path.abspath(__file__)
    os.chdir(REF_SRC)
    sys.path.insert(0, REF_SRC)
    sys.path.insert(1, REPO)
    from core import DefaultConfig
    return DefaultConfig()


CASES = {
    # name: (config overrides, B, T, training, seed, pad_last)
    'eyenet_gru_train': (dict(refine_net_enabled=False, load_screen_content=False), 2, 3, True, 11, 0),
    'eyenet_static_eval': (dict(refine_net_enabled=False, load_screen_content=False,
                                eye_net_use_rnn=False), 2, 2, False, 12, 0),
    'eyenet_lstm_eval': (dict(refine_net_enabled=False, load_screen_content=False,
                              eye_net_rnn_type='LSTM'), 1, 3, False, 13, 0),
    'eyenet_rnn2_eval': (dict(refine_net_enabled=False, load_screen_content=False,
                              eye_net_rnn_type='RNN', eye_net_rnn_num_cells=2), 1, 3, False, 14, 0),
    'eve_cgru_train': (dict(refine_net_enabled=True, load_screen_content=True), 2, 3, True, 21, 0),
    'eve_cgru_train_pogloss': (dict(refine_net_enabled=True, load_screen_content=True,
                                    loss_coeff_PoG_cm_initial=0.01,
                                    loss_coeff_heatmap_ce_initial=0.5,
                                    loss_coeff_heatmap_mse_final=2.0), 1, 2, True, 25, 0),
    'eve_cgru_eval_pad': (dict(refine_net_enabled=True, load_screen_content=True), 2, 4, False, 22, 2),
    'eve_clstm_frozen_train': (dict(refine_net_enabled=True, load_screen_content=True,
                                    refine_net_rnn_type='CLSTM', eye_net_frozen=True), 1, 3, True, 23, 0),
    'eve_crnn_noskip_noscreen_eval': (dict(refine_net_enabled=True, load_screen_content=False,
                                           refine_net_rnn_type='CRNN',
                                           refine_net_use_skip_connections=False), 1, 2, False, 24, 0),
}

DEFAULTS = dict(
    refine_net_enabled=False, load_screen_content=False, eye_net_use_rnn=True,
    eye_net_rnn_type='GRU', eye_net_rnn_num_cells=1, eye_net_frozen=False,
    refine_net_rnn_type='CGRU', refine_net_use_skip_connections=True,
    loss_coeff_PoG_cm_initial=0.0, loss_coeff_heatmap_ce_initial=0.0,
    loss_coeff_heatmap_mse_final=0.0,
)


def run_case(name, spec, config, outdir):
    import torch
    from eve_b200 import synth
    overrides, B, T, training, seed, pad_last = spec
    for k, v in DEFAULTS.items():
        config.override(k, v)
    for k, v in overrides.items():
        config.override(k, v)
    from models.eve import EVE

    torch.manual_seed(0)
    model = EVE(output_predictions=True)
    sd = synth.make_state_dict(synth.eye_net_param_shapes(config), seed, 'eye_net.')
    if config.refine_net_enabled:
        sd.update(synth.make_state_dict(synth.refine_net_param_shapes(config), seed + 1000,
                                        'refine_net.'))
    model.load_state_dict(sd, strict=True)           # pins names and shapes
    model.train(training)

    inputs = synth.make_clip_batch(B, T, seed=seed, with_screen=config.load_screen_content,
                                   pad_last=pad_last)
    captured = {}
    orig = EVE.calculate_losses_and_metrics

    def spy(self, input_dict, intermediate_dict, output_dict):
        for k, v in intermediate_dict.items():
            if isinstance(v, torch.Tensor):
                captured['mid/' + k] = v
        for k in ('g', 'o', 'PoG_px_tobii', 'PoG_cm_tobii', 'PoG_px_tobii_validity',
                  'heatmap_initial', 'heatmap_final', 'left_kappa_fake', 'right_kappa_fake'):
            if k in input_dict:
                captured['label/' + k] = input_dict[k]
        return orig(self, input_dict, intermediate_dict, output_dict)

    EVE.calculate_losses_and_metrics = spy
    np.random.seed(seed)                             # kappas come from np.random (eve.py:468)
    try:
        if training:
            out = model({'synthetic': dict(inputs)}, create_images=True, current_epoch=0.0)
        else:
            with torch.no_grad():
                out = model(dict(inputs), create_images=True)
    finally:
        EVE.calculate_losses_and_metrics = orig

    arrays = {'meta/B': np.int64(B), 'meta/T': np.int64(T), 'meta/seed': np.int64(seed),
              'meta/training': np.bool_(training), 'meta/pad_last': np.int64(pad_last),
              'meta/overrides': np.array(repr(sorted(overrides.items())))}
    skip_big = ('_eye_patch', 'screen_frame', 'both_eye_patch')
    for k, v in out.items():
        if isinstance(v, torch.Tensor) and not k.endswith(skip_big):
            arrays['out/' + k] = v.detach().cpu().numpy()
    for k, v in captured.items():
        if k.endswith(skip_big):
            continue
        if 'rnn_states' in k:
            continue
        if k.startswith('mid/history_') and v.ndim >= 4:
            v = v[:, -1]
        if k in ('mid/heatmap_initial_unaugmented', 'mid/heatmap_initial_augmented'):
            continue
        arrays[k] = v.detach().cpu().numpy()

    if training:
        out['full_loss'].backward()
        for pname, p in model.named_parameters():
            if p.grad is None:
                arrays['gradnone/' + pname] = np.bool_(True)
                continue
            g = p.grad.detach().double().reshape(-1)
            arrays['gradnorm/' + pname] = np.float64(g.norm().item())
            arrays['gradsum/' + pname] = np.float64(g.sum().item())
            gf = p.grad.detach().reshape(-1).numpy()
            arrays['grad/' + pname] = gf.copy() if gf.size <= 20000 else gf[::GRAD_STRIDE].copy()

    path = os.path.join(outdir, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('%-32s %4d arrays  %7.1f KB  full_loss=%.6f' % (
        name, len(arrays), os.path.getsize(path) / 1024.0, float(out['full_loss'])))


def run_module_cases(config, outdir):
    """Direct EyeNet.forward / RefineNet.forward calls with explicit previous states
    (the per-step entry points of SURVEY.md section 8b)."""
    import torch
    from eve_b200 import synth
    for k, v in DEFAULTS.items():
        config.override(k, v)
    config.override('refine_net_enabled', True)
    config.override('load_screen_content', True)
    from models.eye_net import EyeNet
    from models.refine_net import RefineNet
    from models.common import soft_argmax, batch_make_heatmaps
    arrays = {}
    rs = np.random.RandomState(31)
    with torch.no_grad():
        net = EyeNet()
        net.load_state_dict(synth.make_state_dict(synth.eye_net_param_shapes(config), 31), strict=True)
        inp = {'left_eye_patch': torch.from_numpy(rs.uniform(-1, 1, (2, 3, 128, 128)).astype(np.float32)),
               'left_h': torch.from_numpy(rs.uniform(-.2, .2, (2, 2)).astype(np.float32))}
        prev = {'left_eye_rnn_states_0': torch.from_numpy(rs.normal(size=(2, 128)).astype(np.float32))}
        out = {}
        net(inp, out, side='left', previous_output_dict=prev)
        arrays['eyenet/g'] = out['left_g_initial'].numpy()
        arrays['eyenet/pupil'] = out['left_pupil_size'].numpy()
        arrays['eyenet/state'] = out['left_eye_rnn_states_0'].numpy()
        # stage-wise CNN features, to localise errors
        x = inp['left_eye_patch']
        c = net.cnn_layers
        x = c.maxpool(c.relu(c.bn1(c.conv1(x))))
        arrays['eyenet/stem'] = x.numpy()
        for li, layer in enumerate((c.layer1, c.layer2, c.layer3, c.layer4), start=1):
            x = layer(x)
            arrays['eyenet/layer%d' % li] = x.numpy()
        arrays['eyenet/fc'] = c.fc(torch.flatten(c.avgpool(x), 1)).numpy()

        rnet = RefineNet()
        rnet.load_state_dict(synth.make_state_dict(synth.refine_net_param_shapes(config), 1031),
                             strict=True)
        px = torch.from_numpy(np.stack([rs.uniform(0, 1920, 2), rs.uniform(0, 1080, 2)], -1)
                              .astype(np.float32))
        hm = batch_make_heatmaps(px, config.gaze_heatmap_sigma_initial)
        rin = {'screen_frame': torch.from_numpy(rs.uniform(0, 1, (2, 3, 72, 128)).astype(np.float32))}
        prev = {'refinenet_rnn_states_0': torch.from_numpy(
            (0.5 * rs.normal(size=(2, 64, 5, 8))).astype(np.float32))}
        out = {'heatmap_initial': hm}
        rnet(rin, out, previous_output_dict=prev)
        arrays['refine/px_in'] = px.numpy()
        arrays['refine/heatmap_initial'] = hm.numpy()
        arrays['refine/heatmap_final'] = out['heatmap_final'].numpy()
        arrays['refine/state'] = out['refinenet_rnn_states_0'].numpy()
        arrays['refine/softargmax'] = soft_argmax(out['heatmap_final']).numpy()
        # soft-argmax of a one-hot (SURVEY appendix B known answer)
        onehot = torch.zeros(1, 1, 72, 128)
        onehot[0, 0, 10, 20] = 1.0
        arrays['refine/softargmax_onehot'] = soft_argmax(onehot).numpy()
    path = os.path.join(outdir, 'modules.npz')
    np.savez_compressed(path, **arrays)
    print('%-32s %4d arrays  %7.1f KB' % ('modules', len(arrays), os.path.getsize(path) / 1024.0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(REPO, 'tests', 'golden'))
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    outdir = os.path.abspath(args.out)
    os.makedirs(outdir, exist_ok=True)
    config = import_reference()
    import torch
    torch.set_num_threads(os.cpu_count())
    for name, spec in CASES.items():
        if args.only and args.only not in name:
            continue
        run_case(name, spec, config, outdir)
    if not args.only or args.only in 'modules':
        run_module_cases(config, outdir)


if __name__ == '__main__':
    main()
