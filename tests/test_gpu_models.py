"""Parity of the B200 path, called through the drop-in modules (which call the C ABI),
against (1) golden vectors produced by the unmodified reference and (2) the CPU oracle on
seeded inputs for the knob combinations the goldens do not carry gradients for.

Tolerances: BASELINE.json's north_star asks for 1e-3 relative on fp32 gaze vectors / PoG;
the forward comparisons below hold 2e-4 (EyeNet side) and 2e-3 on quantities that pass
through RefineNet + soft-argmax at random weights (the reference's own fp32 noise floor there,
SURVEY.md 7.2); weight gradients are compared in L2 at 2e-2 like tests/test_oracle_golden.py
documents.  Index outputs (validity masks, timestamps) are bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import eve_oracle as O       # noqa: E402  (checker only)
from tests import gpu_util as G          # noqa: E402
from tests import helpers as H           # noqa: E402


@pytest.fixture(params=[0, 1], ids=['fp32-simt', 'tcgen05-bf16x3'])
def conv_mode(request):
    """Run under both kernel selections: 0 = fp32 CUDA cores everywhere (exact products),
    1 = tcgen05 split operands (the default product path: forward convolutions on fp16
    hi + lo planes = 22 mantissa bits, gradients on bf16 hi + lo planes).  Yields a tolerance
    multiplier: measured, mode 1 sits at 1.5-3x the fp32 noise of the reference arithmetic
    itself on RefineNet heatmaps / states and at 4e-6 on EyeNet features."""
    from eve_b200 import lib as L
    lib = L.load()
    prev = lib.eve_get_conv_mode()
    lib.eve_set_conv_mode(request.param)
    yield {0: 1.0, 1: 3.0}[request.param]
    lib.eve_set_conv_mode(prev)


def _load(model, sd):
    model.load_state_dict(sd, strict=True)
    return model.cuda()


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


# ------------------------------------------------------------------ module entry points --
def test_module_entry_points_match_reference(cfg, conv_mode):
    tolx = conv_mode
    from eve_b200 import synth
    from eve_b200.models import EyeNet, RefineNet
    from eve_b200.models.common import batch_make_heatmaps, soft_argmax
    gold = H.load_golden('modules')
    cfg.override('refine_net_enabled', True)
    cfg.override('load_screen_content', True)
    rs = np.random.RandomState(31)
    net = _load(EyeNet(), synth.make_state_dict(synth.eye_net_param_shapes(cfg), 31))
    patch = torch.from_numpy(rs.uniform(-1, 1, (2, 3, 128, 128)).astype(np.float32)).cuda()
    hp = torch.from_numpy(rs.uniform(-.2, .2, (2, 2)).astype(np.float32)).cuda()
    st = torch.from_numpy(rs.normal(size=(2, 128)).astype(np.float32)).cuda()
    with torch.no_grad():
        assert H.rel_err(net.cnn_features(patch).cpu().numpy(), gold['eyenet/fc']) < 1e-4
        out = {}
        net({'left_eye_patch': patch, 'left_h': hp}, out, side='left',
            previous_output_dict={'left_eye_rnn_states_0': st})
        assert H.rel_err(out['left_g_initial'].cpu().numpy(), gold['eyenet/g']) < 1e-4
        assert H.rel_err(out['left_pupil_size'].cpu().numpy(), gold['eyenet/pupil']) < 1e-4
        assert H.rel_err(out['left_eye_rnn_states_0'].cpu().numpy(), gold['eyenet/state']) < 1e-4

        rnet = _load(RefineNet(), synth.make_state_dict(synth.refine_net_param_shapes(cfg), 1031))
        px = torch.from_numpy(np.stack([rs.uniform(0, 1920, 2), rs.uniform(0, 1080, 2)], -1)
                              .astype(np.float32)).cuda()
        hm = batch_make_heatmaps(px, cfg.gaze_heatmap_sigma_initial)
        screen = torch.from_numpy(rs.uniform(0, 1, (2, 3, 72, 128)).astype(np.float32)).cuda()
        prev = torch.from_numpy((0.5 * rs.normal(size=(2, 64, 5, 8))).astype(np.float32)).cuda()
        out = {'heatmap_initial': hm}
        rnet({'screen_frame': screen}, out, previous_output_dict={'refinenet_rnn_states_0': prev})
        assert H.rel_err(out['heatmap_final'].cpu().numpy(), gold['refine/heatmap_final']) < 1e-4 * tolx
        assert H.rel_err(out['refinenet_rnn_states_0'].cpu().numpy(), gold['refine/state']) < 1e-4 * tolx
        assert H.rel_err(soft_argmax(out['heatmap_final']).cpu().numpy(),
                         gold['refine/softargmax']) < 1e-3


# ----------------------------------------------------------------------- golden EVE cases --
@pytest.mark.parametrize('name', H.golden_names())
def test_eve_forward_backward_matches_reference(name, cfg, conv_mode):
    tolx = conv_mode
    from eve_b200.models import EVE
    gold = H.load_golden(name)
    H.apply_case_config(cfg, gold)
    training = bool(gold['meta/training'])
    B, pad_last = int(gold['meta/B']), int(gold['meta/pad_last'])
    model = _load(EVE(output_predictions=True), H.case_state_dict(gold, cfg))
    model.train(training)
    inputs = _cuda(H.case_inputs(gold, cfg))
    np.random.seed(int(gold['meta/seed']))       # kappas come from np.random (eve.py:468)
    if training:
        out = model({'synthetic': inputs}, create_images=True, current_epoch=0.0)
    else:
        with torch.no_grad():
            out = model(inputs, create_images=True)
    mid = model.last_intermediates

    checked = 0
    for k, ref in gold.items():
        if k.startswith('out/'):
            key = k[4:]
            if key not in out or out[key] is None:
                continue
            got = out[key]
        elif k.startswith('mid/'):
            key = k[4:]
            if key not in mid:
                continue
            got = mid[key]
            if key.startswith('history_'):
                got = got[:, -1]
        else:
            continue
        got = got.detach().cpu().numpy()
        if ref.dtype == np.bool_ or np.issubdtype(ref.dtype, np.integer):
            assert np.array_equal(got, ref), k
        else:
            assert got.shape == ref.shape, (k, got.shape, ref.shape)
            tol = 2e-3 if ('final' in key or 'refined' in key or key == 'full_loss') else 2e-4
            if tolx > 1:            # split-operand products; see DESIGN.md (precision)
                tol = 4e-3 if tol == 2e-3 else 4e-4
            if pad_last and got.ndim >= 1 and got.shape[0] == B:
                # Zero-padded frames (all-zero images, validity 0) put InstanceNorm at
                # var ~ 0, where rstd = 316 amplifies fp32 summation-order noise: hold the
                # padded clip to 3e-2 and every other clip to the normal bar.
                assert H.rel_err(got[-1], ref[-1]) < 3e-2, (k, H.rel_err(got[-1], ref[-1]))
                got, ref = got[:-1], ref[:-1]
            assert H.rel_err(got, ref) < tol, (k, H.rel_err(got, ref))
        checked += 1
    assert checked >= 20, checked
    for must in ('full_loss', 'left_pupil_size', 'g_initial', 'PoG_px_initial'):
        assert must in out
    # derived labels (EVE.calculate_additional_labels, eve.py:441-543) against the reference's
    labels = 0
    for k, ref in gold.items():
        if not k.startswith('label/'):
            continue
        key = k[len('label/'):]
        assert key in inputs, key
        got = inputs[key].detach().cpu().numpy()
        if ref.dtype == np.bool_ or np.issubdtype(ref.dtype, np.integer):
            assert np.array_equal(got.astype(ref.dtype), ref), k     # validity masks: bit-exact
        else:
            assert got.shape == ref.shape, (k, got.shape, ref.shape)
            assert H.rel_err(got, ref) < 2e-6, (k, H.rel_err(got, ref))
        labels += 1
    if any(k.startswith('label/') for k in gold):
        assert labels >= 5, labels

    if training:
        out['full_loss'].backward()
        params = dict(model.named_parameters())
        floor = 1e-5 * max(float(v) for k, v in gold.items() if k.startswith('gradnorm/'))

        def grad_errors(named):
            errs = {}
            for k, ref in gold.items():
                if not k.startswith('gradnorm/'):
                    continue
                pname = k[len('gradnorm/'):]
                g = named[pname].grad
                assert g is not None, pname
                sample = gold['grad/' + pname]
                gf = g.reshape(-1).cpu().numpy()
                gs = gf if gf.size <= 20000 else gf[::H.GRAD_STRIDE]
                errs[pname] = (abs(float(g.double().norm()) - float(ref)), float(ref),
                               float(np.linalg.norm(gs.astype(np.float64) - sample)),
                               float(np.linalg.norm(sample.astype(np.float64))))
            return errs

        errs = grad_errors(params)
        # RefineNet gradients at random weights carry several % of pure fp32 ordering noise
        # (the reference's own fp32 run sits 0.4-1.2e-2 from an fp64 evaluation, see
        # test_refinenet_sequences_* for the noise-relative check); observed up to 3.3e-2.
        gtol = 5e-2
        base = None
        if tolx > 1:
            # gradients use bf16 hi+lo planes (16 mantissa bits): where RefineNet's ill-conditioned
            # heatmap gradient flows back into EyeNet the deviation reaches 5-7 % L2 (2 % in
            # exact-fp32 mode); forward outputs -- the north_star bar -- are unaffected
            gtol = 8e-2
            # split-operand mode: additionally allow 3x whatever discrepancy the exact-fp32
            # kernels show on the same tensor (i.e. summation-order noise alone) -- the EyeNet
            # tail tensors that receive RefineNet's gradient through the heatmap are the
            # noisiest (2 % in fp32 mode)
            from eve_b200 import lib as L
            lib = L.load()
            lib.eve_set_conv_mode(0)
            try:
                twin = _load(EVE(output_predictions=True), H.case_state_dict(gold, cfg)).train()
                np.random.seed(int(gold['meta/seed']))
                twin({'synthetic': _cuda(H.case_inputs(gold, cfg))}, create_images=True,
                     current_epoch=0.0)['full_loss'].backward()
                base = grad_errors(dict(twin.named_parameters()))
            finally:
                lib.eve_set_conv_mode(1)
        bad = []
        for pname, (dn, rn, dl2, rl2) in errs.items():
            lim_n = gtol * max(rn, 1e-6) + floor
            lim_l = gtol * rl2 + floor
            if base is not None:
                lim_n = max(lim_n, 3.0 * base[pname][0])
                lim_l = max(lim_l, 3.0 * base[pname][2])
            if dn > lim_n:
                bad.append((pname, 'norm', dn, lim_n))
            if dl2 > lim_l:
                bad.append((pname, 'l2', dl2, lim_l))
        n = len(errs)
        assert not bad, bad
        for k in gold:
            if k.startswith('gradnone/'):
                g = params[k[len('gradnone/'):]].grad
                assert g is None or float(g.abs().max()) == 0.0, k
        assert n > 30


# --------------------------------------------------- oracle comparisons with gradients --
TAIL_CASES = [('GRU', 1, True), ('LSTM', 1, True), ('RNN', 2, True), ('LSTM', 2, False),
              (None, 1, True), ('GRU', 2, False)]


@pytest.mark.parametrize('rnn,cells,head_pose', TAIL_CASES)
def test_eyenet_tail_sequences_with_state_and_gradients(cfg, rnn, cells, head_pose):
    """eye_net.py:109-140 for every RNN variant, with non-zero initial states and gradients
    w.r.t. features, initial states and every weight."""
    from eve_b200 import synth
    from eve_b200.models import EyeNet
    cfg.override('eye_net_use_rnn', rnn is not None)
    if rnn:
        cfg.override('eye_net_rnn_type', rnn)
        cfg.override('eye_net_rnn_num_cells', cells)
    cfg.override('eye_net_use_head_pose_input', head_pose)
    sd = synth.make_state_dict(synth.eye_net_param_shapes(cfg), 77)
    net = _load(EyeNet(), sd)
    S, T, nf = 3, 4, 128
    g = torch.Generator().manual_seed(8)
    feat = torch.randn(S, T, nf, generator=g)
    hp = torch.rand(S, T, 2, generator=g) - 0.5
    h0 = torch.randn(cells, S, nf, generator=g) * 0.5 if rnn else None
    c0 = torch.randn(cells, S, nf, generator=g) * 0.5 if rnn == 'LSTM' else None
    wg = torch.randn(S, T, 2, generator=g)
    wp = torch.randn(S, T, generator=g)
    wh = torch.randn(cells, S, nf, generator=g) if rnn else None

    # oracle (fp64, CPU)
    osd = {'eye_net.' + k: v.double().requires_grad_(True) for k, v in sd.items()}
    f64 = feat.double().requires_grad_(True)
    h64 = h0.double().requires_grad_(True) if rnn else None
    c64 = c0.double().requires_grad_(True) if c0 is not None else None
    states = None
    if rnn:
        states = [(h64[i], c64[i]) if rnn == 'LSTM' else h64[i] for i in range(cells)]
    gs, ps = [], []
    for t in range(T):
        go, po, states = O.eye_net_tail_step(osd, cfg, f64[:, t], hp[:, t].double(), states or None)
        gs.append(go)
        ps.append(po)
    og, op = torch.stack(gs, 1), torch.stack(ps, 1)
    loss = (og * wg.double()).sum() + (op * wp.double()).sum()
    if rnn:
        fin = torch.stack([s[0] if isinstance(s, tuple) else s for s in states], 0)
        loss = loss + (fin * wh.double()).sum()
    loss.backward()

    fc = feat.cuda().requires_grad_(True)
    hc = h0.cuda().requires_grad_(True) if rnn else None
    cc = c0.cuda().requires_grad_(True) if c0 is not None else None
    gg, gp, hT, cT = net.tail_sequence(fc, hp.cuda(), hc, cc)
    assert G.rel(gg, og) < 2e-5
    assert G.rel(gp, op) < 2e-5
    closs = (gg * wg.cuda()).sum() + (gp * wp.cuda()).sum()
    if rnn:
        assert G.rel(hT, fin) < 2e-5
        closs = closs + (hT * wh.cuda()).sum()
    closs.backward()
    assert G.rel(fc.grad, f64.grad) < 1e-4
    if rnn:
        assert G.rel(hc.grad, h64.grad) < 1e-4
    if cc is not None:
        assert G.rel(cc.grad, c64.grad) < 1e-4
    for name, p in net.named_parameters():
        if name.startswith('cnn_layers.'):
            continue
        want = osd['eye_net.' + name].grad
        assert p.grad is not None, name
        assert G.rel(p.grad, want) < 2e-4, name


def test_eyenet_cnn_gradients_match_oracle(cfg, conv_mode):
    tolx = conv_mode
    """ResNet-18/InstanceNorm forward + every conv weight gradient against the fp64 oracle."""
    from eve_b200 import synth
    from eve_b200.models import EyeNet
    sd = synth.make_state_dict(synth.eye_net_param_shapes(cfg), 78)
    net = _load(EyeNet(), sd)
    g = torch.Generator().manual_seed(9)
    x = torch.rand(3, 3, 128, 128, generator=g) * 2 - 1
    wf = torch.randn(3, 128, generator=g)
    def oracle(dtype):
        osd = {'eye_net.' + k: v.detach().clone().to(dtype).requires_grad_(k.startswith('cnn_layers.'))
               for k, v in sd.items()}
        want = O.resnet18_in_features(osd, 'eye_net.cnn_layers.', x.to(dtype))
        (want * wf.to(dtype)).sum().backward()
        return want.detach(), {k: v.grad for k, v in osd.items()}

    want, g64 = oracle(torch.float64)
    _, g32 = oracle(torch.float32)
    got = net.cnn_features(x.cuda())
    assert G.rel(got, want) < 2e-5 * tolx
    (got * wf.cuda()).sum().backward()

    def l2(a, b):
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        return float((a - b).norm() / (b.norm() + 1e-30))

    # yardstick: the fp32 noise of the reference arithmetic itself (the stem gradient sits
    # at the end of a 20-conv / 20-InstanceNorm backward chain and is the noisiest)
    for name, p in net.named_parameters():
        if not name.startswith('cnn_layers.'):
            continue
        ref = g64['eye_net.' + name]
        e, e32 = l2(p.grad, ref), l2(g32['eye_net.' + name], ref)
        # + 1e-2: one ReLU mask flip (a pre-activation within rounding distance of zero) moves
        # every weight of one stem filter -- observed 4.7e-3 from a single flipped element
        assert e < 4.0 * tolx * e32 + 1e-2, (name, e, e32)


REFINE_CASES = [('CGRU', 1, True, True), ('CRNN', 2, True, True), ('CGRU', 2, False, False),
                ('CLSTM', 1, True, True), (None, 1, True, False)]


@pytest.mark.parametrize('rnn,cells,skip,screen', REFINE_CASES)
def test_refinenet_sequences_with_state_and_gradients(cfg, rnn, cells, skip, screen, conv_mode):
    """refine_net.py:237-255 over B x T with non-zero initial states: heatmaps, final states
    and gradients w.r.t. the input heatmap, the initial state and every weight."""
    from eve_b200 import synth
    from eve_b200.models import RefineNet
    tolx = conv_mode
    cfg.override('refine_net_enabled', True)
    cfg.override('load_screen_content', screen)
    cfg.override('refine_net_use_skip_connections', skip)
    cfg.override('refine_net_use_rnn', rnn is not None)
    if rnn:
        cfg.override('refine_net_rnn_type', rnn)
        cfg.override('refine_net_rnn_num_cells', cells)
    sd = synth.make_state_dict(synth.refine_net_param_shapes(cfg), 79)
    net = _load(RefineNet(), sd)
    B, T = 2, 3
    g = torch.Generator().manual_seed(10)
    px = torch.stack([torch.rand(B, T, generator=g) * 1920, torch.rand(B, T, generator=g) * 1080], -1)
    hm = O.make_heatmaps(px, 10.0)
    scr = torch.rand(B, T, 3, 72, 128, generator=g) if screen else None
    h0 = torch.randn(cells, B, 64, 5, 8, generator=g) * 0.5 if rnn else None
    c0 = torch.randn(cells, B, 64, 5, 8, generator=g) * 0.5 if rnn == 'CLSTM' else None
    wo = torch.randn(B, T, 1, 72, 128, generator=g)
    wh = torch.randn(cells, B, 64, 5, 8, generator=g) if rnn in ('CGRU', 'CRNN') else None
    def oracle(dtype):
        osd = {'refine_net.' + k: v.detach().clone().to(dtype).requires_grad_(True)
               for k, v in sd.items()}
        hmo = hm.detach().clone().to(dtype).requires_grad_(True)
        ho = h0.detach().clone().to(dtype).requires_grad_(True) if rnn else None
        states = None
        if rnn:
            states = [(ho[i], c0[i].to(dtype)) if rnn == 'CLSTM' else ho[i] for i in range(cells)]
        outs = []
        for t in range(T):
            o, states = O.refine_net_step(osd, cfg, scr[:, t].to(dtype) if screen else None,
                                          hmo[:, t], states or None)
            outs.append(o)
        want = torch.stack(outs, 1)
        loss = (want * wo.to(dtype)).sum()
        fin = None
        if rnn:
            fin = [torch.stack([s[0] if isinstance(s, tuple) else s for s in states], 0)]
            if rnn == 'CLSTM':
                fin.append(torch.stack([s[1] for s in states], 0))
            if wh is not None:
                loss = loss + (fin[0] * wh.to(dtype)).sum()
        loss.backward()
        grads = {k: v.grad for k, v in osd.items()}
        return want.detach(), fin, hmo.grad, (ho.grad if rnn else None), grads

    want, fin, dhm64, dh64, g64 = oracle(torch.float64)
    _, _, dhm32, dh32, g32 = oracle(torch.float32)

    def l2(a, b):
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        return float((a - b).norm() / (b.norm() + 1e-30))

    hmc = hm.detach().cuda().requires_grad_(True)
    hc = h0.detach().cuda().requires_grad_(True) if rnn else None
    got, hT, cT = net.sequence(scr.cuda() if screen else None, hmc, hc,
                               c0.cuda() if c0 is not None else None)
    assert G.rel(got, want) < 1e-4 * tolx
    closs = (got * wo.cuda()).sum()
    if rnn:
        assert G.rel(hT, fin[0]) < 1e-4 * tolx
        if rnn == 'CLSTM':
            assert G.rel(cT, fin[1]) < 1e-4 * tolx
    if wh is not None:
        closs = closs + (hT * wh.cuda()).sum()
    closs.backward()

    # Gradients: RefineNet at random weights is ill-conditioned (InstanceNorm over nearly
    # constant maps), so the yardstick is the fp32 noise of the reference arithmetic itself:
    # our fp32 result must be as close to the fp64 truth as the oracle's own fp32 run,
    # within a factor 3 (+1e-4 absolute floor on the relative L2 error).
    def bar(ref32, ref64):
        return 3.0 * tolx * l2(ref32, ref64) + 1e-4

    assert l2(hmc.grad, dhm64) < bar(dhm32, dhm64), (l2(hmc.grad, dhm64), l2(dhm32, dhm64))
    if wh is not None:
        assert l2(hc.grad, dh64) < bar(dh32, dh64), (l2(hc.grad, dh64), l2(dh32, dh64))
    scale = max(float(v.norm()) for v in g64.values() if v is not None)
    rows = []
    for name, p in net.named_parameters():
        ref = g64['refine_net.' + name]
        if ref is None or float(ref.norm()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) < 1e-6, name
            continue
        assert p.grad is not None, name
        if float(ref.norm()) < 1e-6 * scale:
            # a bias feeding an InstanceNorm: the true gradient is zero, both sides are noise
            assert float(p.grad.norm()) < 1e-3 * scale, name
            continue
        rows.append((name, l2(p.grad, ref), l2(g32['refine_net.' + name], ref)))
    assert len(rows) > 40
    # the fp32 noise level of this network / input (single tensors can be lucky)
    noise = float(np.median([e32 for _, _, e32 in rows]))
    for name, e, e32 in rows:
        assert e < 4.0 * tolx * max(e32, noise) + 1e-4, (name, e, e32, noise)


def test_per_step_and_time_batched_paths_agree(cfg):
    """EyeNet.forward / RefineNet.forward called once per time step with
    previous_output_dict (the reference's own calling pattern, eve.py:91-147) give what the
    time-batched EVE.forward computes."""
    from eve_b200 import synth
    from eve_b200.models import EVE
    from eve_b200.models.common import batch_make_heatmaps
    cfg.override('refine_net_enabled', True)
    cfg.override('load_screen_content', True)
    B, T = 2, 3
    sd = synth.make_state_dict(synth.eye_net_param_shapes(cfg), 5, 'eye_net.')
    sd.update(synth.make_state_dict(synth.refine_net_param_shapes(cfg), 1005, 'refine_net.'))
    model = _load(EVE(output_predictions=True), sd).eval()
    inputs = _cuda(synth.make_clip_batch(B, T, seed=5))
    with torch.no_grad():
        out = model(dict(inputs))
        mid = model.last_intermediates
        prev = None
        for t in range(T):
            sub = {k: v[:, t] for k, v in inputs.items()}
            o = {}
            model.eye_net(sub, o, side='left', previous_output_dict=prev)
            model.eye_net(sub, o, side='right', previous_output_dict=prev)
            assert G.rel(o['left_g_initial'], mid['left_g_initial'][:, t]) < 1e-5
            assert G.rel(o['right_pupil_size'], mid['right_pupil_size'][:, t]) < 1e-5
            o['heatmap_initial'] = mid['heatmap_initial'][:, t]
            model.refine_net(sub, o, previous_output_dict=prev)
            assert G.rel(o['heatmap_final'], mid['heatmap_final'][:, t]) < 1e-4
            prev = o
    assert out['PoG_px_final'].shape == (B, T, 2)


def test_empty_and_single_frame_inputs(cfg):
    from eve_b200 import synth
    from eve_b200.models import EyeNet
    net = _load(EyeNet(), synth.make_state_dict(synth.eye_net_param_shapes(cfg), 3))
    with torch.no_grad():
        assert net.cnn_features(torch.zeros(0, 3, 128, 128, device='cuda')).shape == (0, 128)
        f = net.cnn_features(torch.zeros(1, 3, 128, 128, device='cuda'))
        assert f.shape == (1, 128) and bool(torch.isfinite(f).all())


@pytest.mark.parametrize('kind', ['CRNN', 'CLSTM', 'CGRU'])
def test_standalone_conv_rnn_cells(cfg, kind, conv_mode):
    """models.common.{CRNNCell, CLSTMCell, CGRUCell}: same parameter names as the reference
    modules and the same step arithmetic (oracle: common.py:331-415), with gradients."""
    from eve_b200.models import common as MC
    tolx = conv_mode
    cls = {'CRNN': MC.CRNNCell, 'CLSTM': MC.CLSTMCell, 'CGRU': MC.CGRUCell}[kind]
    cell = cls(input_size=64, hidden_size=64).cuda()
    names = {'CRNN': {'cell.weight', 'cell.bias'}, 'CLSTM': {'gates.weight', 'gates.bias'},
             'CGRU': {'gates_1.weight', 'gates_1.bias', 'gate_2.weight', 'gate_2.bias'}}[kind]
    assert set(cell.state_dict().keys()) == names
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 64, 5, 8, generator=g)
    h = torch.randn(2, 64, 5, 8, generator=g) * 0.5
    c = torch.randn(2, 64, 5, 8, generator=g) * 0.5
    sd = {'p.' + k: v.detach().cpu().double().requires_grad_(True) for k, v in cell.state_dict().items()}
    x64 = x.double().requires_grad_(True)
    state = (h.double(), c.double()) if kind == 'CLSTM' else h.double()
    want = O.conv_rnn_cell(kind, sd, 'p.', x64, state)
    want_h = want[0] if kind == 'CLSTM' else want
    want_h.sum().backward()
    xc = x.cuda().requires_grad_(True)
    got = cell(xc, (h.cuda(), c.cuda()) if kind == 'CLSTM' else h.cuda())
    got_h = got[0] if kind == 'CLSTM' else got
    assert G.rel(got_h, want_h) < 2e-5 * tolx
    if kind == 'CLSTM':
        assert G.rel(got[1], want[1]) < 2e-5 * tolx
    got_h.sum().backward()
    assert G.rel(xc.grad, x64.grad) < 1e-4 * tolx
    for k, p in cell.named_parameters():
        assert G.rel(p.grad, sd['p.' + k].grad) < 1e-4 * tolx, k
    # zero initial state when no previous state is given (common.py:344-346)
    with torch.no_grad():
        first = cell(xc)
    assert (first[0] if kind == 'CLSTM' else first).shape == (2, 64, 5, 8)


def test_inference_stream_900_frames_in_chunks(cfg):
    """BASELINE config 5 at full size: RefineNet over a 900-frame stream (B=1, no grad).
    Size-independent property: processing the stream in three 300-frame chunks with the
    ConvGRU state carried across chunk boundaries reproduces the one-shot result, and frames are
    independent of the batch they were processed in (per-sample norms)."""
    from eve_b200 import synth
    from eve_b200.models import RefineNet
    from eve_b200.models.common import batch_make_heatmaps
    cfg.override('refine_net_enabled', True)
    cfg.override('load_screen_content', True)
    net = _load(RefineNet(), synth.make_state_dict(synth.refine_net_param_shapes(cfg), 11)).eval()
    T = 900
    g = torch.Generator().manual_seed(12)
    px = torch.stack([torch.rand(1, T, generator=g) * 1920, torch.rand(1, T, generator=g) * 1080], -1)
    screen = torch.rand(1, T, 3, 72, 128, generator=g).cuda()
    with torch.no_grad():
        hm = batch_make_heatmaps(px.cuda(), cfg.gaze_heatmap_sigma_initial)
        full, hT, _ = net.sequence(screen, hm)
        outs, state = [], None
        for t0 in range(0, T, 300):
            o, state, _ = net.sequence(screen[:, t0:t0 + 300], hm[:, t0:t0 + 300], state)
            outs.append(o)
        chunked = torch.cat(outs, 1)
    assert full.shape == (1, T, 1, 72, 128)
    assert bool(torch.isfinite(full).all())
    assert torch.equal(chunked, full)
    assert torch.equal(state, hT)
    assert float(full.min()) > 0.0 and float(full.max()) < 1.0


def test_long_clip_training_step_t60(cfg):
    """BASELINE config 4 per-GPU shape (B=8, T=60, full EVE): one optimisation step runs,
    the loss is finite, and the gradient norm equals the norm of the packed flat buffer."""
    from eve_b200 import synth
    from eve_b200.models import EVE
    from eve_b200.parallel import FlatAdamTrainer
    cfg.override('refine_net_enabled', True)
    cfg.override('load_screen_content', True)
    sd = synth.make_state_dict(synth.eye_net_param_shapes(cfg), 21, 'eye_net.')
    sd.update(synth.make_state_dict(synth.refine_net_param_shapes(cfg), 1021, 'refine_net.'))
    model = _load(EVE(), sd).train()
    tr = FlatAdamTrainer(model, lr=1e-5)
    inputs = _cuda(synth.make_clip_batch(8, 60, seed=3))
    np.random.seed(0)
    out = model({'x': inputs}, current_epoch=0.0)
    assert out['left_pupil_size'].shape == (8, 60)
    out['full_loss'].backward()
    tr.gather_grads()
    want = float(tr.grad.double().norm())
    tr.apply()
    assert np.isfinite(float(out['full_loss']))
    assert abs(float(tr.last_grad_norm) - want) < 1e-4 * want
    del out, model, tr
    torch.cuda.empty_cache()
