"""bench.py -- eye-frames/sec (forward + backward + gradient step) of the EVE hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload eve_refine|eyenet_static]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU arm (oracle port of the reference)

One "step" = one optimisation step over one synthetic batch of B=8 clips x T=30 frames per
GPU (BASELINE.json): time-batched EVE.forward (EyeNet for both eyes [+ GazeRefineNet]),
full_loss.backward(), one gradient allreduce when N > 1, fused clip + Adam.  An eye-frame is
one 128x128 eye patch pushed through EyeNet: 2*B*T = 480 per GPU per step.

Rank 0 prints ONE JSON line (see the keys at the bottom).  `value` is measured with inputs
resident in HBM; `e2e` repeats the measurement through the same public call with the batch
in pinned host memory (H2D copy + D2H read of the loss inside the timed region).  The
`roofline` object times the convolution kernels (the dense contractions that dominate the
step) live with CUDA events recorded on their launching stream inside the timed region.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = 'eye-frames/sec (fwd+bwd) at seq_len=30, 128x128 patches'
UNIT = 'eye-frames/s'

WORKLOADS = {
    # BASELINE.json configs[2]: the north_star's target workload (eye patches + screen frames)
    'eve_refine': dict(desc='EyeNet(GRU) x2 eyes + GazeRefineNet(CGRU), 128x72 screen frames, '
                            'seq_len=30, batch=8 clips/GPU (BASELINE configs[2])',
                       overrides=dict(refine_net_enabled=True, load_screen_content=True)),
    # BASELINE.json configs[1]
    'eyenet_static': dict(desc='EyeNet static (eye_net_use_rnn=0, refine_net_enabled=0), '
                               'seq_len=30, batch=8 clips/GPU (BASELINE configs[1])',
                          overrides=dict(refine_net_enabled=False, load_screen_content=False,
                                         eye_net_use_rnn=False)),
}
# algorithmic conv-stack FLOPs per unit, fwd+bwd
FLOP_EYE_FRAME = 3477409536.0
FLOP_SCREEN_FRAME = 9629761536.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=80)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='eve_refine', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=8, help='clips per GPU')
    ap.add_argument('--seq-len', type=int, default=30)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the secondary workload line')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-input leg (profiling runs)')
    ap.add_argument('--leg', default=None, choices=['t60', 'stream900', 'stock_torch_b200'],
                    help='run ONE secondary measurement in this process and print its JSON '
                         '(bench.py runs each of them in a child process of its own)')
    return ap.parse_args()


def configure(workload):
    from eve_b200.config import DefaultConfig
    cfg = DefaultConfig()
    cfg.reset()
    for k, v in WORKLOADS[workload]['overrides'].items():
        cfg.override(k, v)
    return cfg


def build_state_dict(cfg, seed=0):
    from eve_b200 import synth
    sd = synth.make_state_dict(synth.eye_net_param_shapes(cfg), seed, 'eye_net.')
    if cfg.refine_net_enabled:
        sd.update(synth.make_state_dict(synth.refine_net_param_shapes(cfg), seed + 1000,
                                        'refine_net.'))
    return sd


# --------------------------------------------------------------------------- clocks --
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                power.append(float(r[3]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                  'sw_power_cap'), r[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None,
                'sm_max_mhz': max(smax) if smax else None,
                'power_w_max': max(power) if power else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------ reference --
def run_reference_step(cfg, sd, B, T, seed, threads):
    """One fwd+bwd of the CPU oracle (the restatement of the reference's PyTorch path) on a
    B x T sample; returns (eye_frames, seconds)."""
    import numpy as np
    import torch
    from eve_b200 import synth
    from oracle import eve_oracle as O
    torch.set_num_threads(threads)
    inputs = synth.make_clip_batch(B, T, seed=seed, with_screen=bool(cfg.load_screen_content))
    np.random.seed(seed)
    std = np.radians(cfg.refine_net_offset_augmentation_sigma)
    kap = {'left': torch.from_numpy(np.random.normal(size=(B, 2), scale=std).astype(np.float32)),
           'right': torch.from_numpy(np.random.normal(size=(B, 2), scale=std).astype(np.float32))}
    osd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    t0 = time.perf_counter()
    out, _ = O.eve_forward(osd, cfg, inputs, True, kap)
    out['full_loss'].backward()
    return 2 * B * T, time.perf_counter() - t0


REF_BUDGET_S = 150.0     # timed CPU work of the reference arm (seconds)


def _timed_reference_steps(cfg, sd, B, T, threads, want_steps, budget_s, min_steps):
    """Warm-up (one tiny step for the oneDNN primitive caches, one full-size step that also sizes
    the budget), then full-size fwd+bwd steps of the SAME B x T workload the B200 arm runs."""
    run_reference_step(cfg, sd, 1, 2, 10, threads)
    _, s0 = run_reference_step(cfg, sd, B, T, 11, threads)
    k = max(min_steps, min(want_steps, int(budget_s / max(s0, 1e-3))))
    frames, sec = 0, 0.0
    for i in range(k):
        f, s = run_reference_step(cfg, sd, B, T, 100 + i, threads)
        frames += f
        sec += s
    return k, frames, sec


def _reference_sample_text(B, T, threads, k, frames, sec):
    import torch
    return ('oracle/eve_oracle.py = CPU restatement of the reference PyTorch path (fp32, torch %s, '
            '%d threads; the reference itself is pure Python without a build system and is not '
            'present on the GPU box): warm-up (B=1,T=2 then one B=%d,T=%d step), then %d x fwd+bwd of '
            'the full B=%d clips x T=%d frames workload = %d eye-frames in %.1f s'
            % (torch.__version__, threads, B, T, k, B, T, frames, sec))


def cpu_baseline(cfg, sd, B, T, budget_s=30.0):
    threads = os.cpu_count() or 1
    k, frames, sec = _timed_reference_steps(cfg, sd, B, T, threads, 2, budget_s, 1)
    return {'value': frames / sec, 'unit': UNIT, 'cores': threads, 'kind': 'port',
            'sample': _reference_sample_text(B, T, threads, k, frames, sec)}


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference
    is pure Python/PyTorch and is not present on the GPU box, so this is the oracle port
    (kind "port"), on all host threads, on the SAME config as the B200 arm (B x T per step);
    the number of timed steps is bounded by REF_BUDGET_S of CPU work (at least 3)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = configure(args.workload)
    sd = build_state_dict(cfg)
    threads = os.cpu_count() or 1
    B, T = args.batch, args.seq_len
    k, frames, sec = _timed_reference_steps(cfg, sd, B, T, threads, args.steps, REF_BUDGET_S, 3)
    value = frames / sec
    sample = _reference_sample_text(B, T, threads, k, frames, sec)
    _emit({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': k, 'steps_requested': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * sec / k,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': WORKLOADS[args.workload]['desc'], 'global_batch': args.batch,
                   'seq_len': args.seq_len, 'eye_frames_per_step': 2 * B * T,
                   'sample_per_step': 'B=%d,T=%d (the full workload)' % (B, T)},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    })


# ------------------------------------------------------------------------- B200 arm --
def make_batch(B, T, cfg, seed, pinned):
    from eve_b200 import synth
    d = synth.make_clip_batch(B, T, seed=seed, with_screen=bool(cfg.load_screen_content))
    if pinned:
        d = {k: v.pin_memory() for k, v in d.items()}
    return d


def _barrier_sync(world):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed_graph_run(step_fn, inputs_fn, steps, extra_warmup, world, device, read_loss,
                    prefetch=False):
    """K timed replays of the captured step (L2 flushed between them); returns max-over-ranks
    total ms and the last loss.  ``read_loss``: D2H read of the loss inside every timed step.
    ``prefetch``: host inputs are double-buffered -- while step i computes, the H2D copy of step
    i+1's batch runs on a second stream (one full H2D copy inside every timed step; only the very
    first batch is staged before the clock starts, and the last timed step stages a batch that is
    never consumed)."""
    import torch
    import torch.distributed as dist
    flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device=device)  # > 126 MB L2
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()
    for i in range(extra_warmup):
        step_fn(inputs_fn(i))
    if prefetch:
        step_fn.prefetch(inputs_fn(extra_warmup))
    _barrier_sync(world)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    last = None
    for i in range(steps):
        flush.zero_()                       # evict L2 between timed iterations
        ev0[i].record()
        if prefetch:
            loss = step_fn(None)                                # staged batch -> static buffers
            step_fn.prefetch(inputs_fn(extra_warmup + i + 1))   # next batch, overlapped H2D
        else:
            loss = step_fn(inputs_fn(extra_warmup + i))
        if read_loss:
            host_loss.copy_(loss, non_blocking=False)      # D2H + sync, like training.py:506
            last = float(host_loss)
        ev1[i].record()
    _barrier_sync(world)
    if last is None:
        last = float(loss)
    ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), last


def profiled_eager_run(model, trainer, inputs_fn, steps, world, device):
    """The same step launched eagerly with the library's conv profiler on: CUDA events around
    every convolution kernel on its launching stream.  Returns (total ms, per-kind dict)."""
    import torch
    from eve_b200 import lib as L
    lib = L.load()
    flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device=device)

    def one(i):
        out = model({'bench': dict(inputs_fn(i))}, current_epoch=0.0)
        trainer.step(out['full_loss'])

    one(0)
    _barrier_sync(world)
    # `ncu --profile-from-start off ... python bench.py` captures exactly these eagerly launched
    # steps (the launch list / DRAM traffic under profiles/); a no-op without a profiler attached
    ncu_range = os.environ.get('EVE_BENCH_NCU_RANGE') == '1'
    if ncu_range:
        torch.cuda.cudart().cudaProfilerStart()
    lib.eve_profile_reset()
    lib.eve_profile_enable(1)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for i in range(steps):
        flush.zero_()
        ev0[i].record()
        one(1 + i)
        ev1[i].record()
    _barrier_sync(world)
    lib.eve_profile_enable(0)
    if ncu_range:
        torch.cuda.cudart().cudaProfilerStop()
    ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    prof = {}
    for kind, name in ((0, 'conv_fwd'), (1, 'conv_dgrad'), (2, 'conv_wgrad')):
        v = [C.c_double(), C.c_double(), C.c_double(), C.c_longlong()]
        L.check(lib.eve_profile_read(kind, C.byref(v[0]), C.byref(v[1]), C.byref(v[2]),
                                     C.byref(v[3])), 'eve_profile_read')
        prof[name] = {'ms': v[0].value, 'flops': v[1].value, 'bytes': v[2].value,
                      'launches': v[3].value}
    lib.eve_profile_reset()
    return ms, prof


# ---------------------------------------------------------------- secondary measurements --
def _event_ms(fn, reps, device, flush=None):
    import torch
    ms = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize(device)
        ms += a.elapsed_time(b)
    return ms / reps


def leg_t60(args, device):
    """BASELINE configs[3] per-GPU shape: full EVE, seq_len=60, batch=8 (graph-replayed step)."""
    import numpy as np
    from eve_b200.graph import GraphedTrainStep
    from eve_b200.models import EVE
    from eve_b200.parallel import FlatAdamTrainer
    cfg = configure('eve_refine')
    np.random.seed(77)
    model = EVE()
    model.load_state_dict(build_state_dict(cfg), strict=True)
    model = model.to(device).train()
    trainer = FlatAdamTrainer(model)
    B, T = args.batch, 60
    dev = [{k: v.to(device) for k, v in make_batch(B, T, cfg, seed=500 + i, pinned=False).items()}
           for i in range(2)]
    step_fn = GraphedTrainStep(model, trainer, dev[0], warmup=3, tag='bench')
    steps = 10
    ms, _ = timed_graph_run(step_fn, lambda i: dev[i % 2], steps, 1, 1, device, False)
    step_fn.close()
    return {'workload': 'full EVE (EyeNet x2 + GazeRefineNet CGRU), seq_len=60, batch=%d clips '
                        '(BASELINE configs[3], per-GPU shape)' % B,
            'value': 2 * B * T * steps / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms / steps,
            'steps': steps}


def leg_stream900(args, device):
    """BASELINE configs[4]: inference-only RefineNet stream, seq_len=900, batch=1 (heatmap raster
    + RefineNet with the ConvGRU state carried across all 900 frames + soft-argmax), screen
    frames/s, next to the CPU oracle on a bounded sample of the same stream."""
    import torch
    from eve_b200 import synth
    from eve_b200.models import RefineNet
    from eve_b200.models.common import batch_make_heatmaps, soft_argmax
    from oracle import eve_oracle as O      # CPU baseline of this leg only
    cfg = configure('eve_refine')
    T = 900
    sd = synth.make_state_dict(synth.refine_net_param_shapes(cfg), 1000)
    net = RefineNet()
    net.load_state_dict(sd)
    net = net.to(device).eval()
    g = torch.Generator().manual_seed(0)
    px = torch.stack([torch.rand(1, T, generator=g) * 1920, torch.rand(1, T, generator=g) * 1080], -1)
    screen = torch.rand(1, T, 3, 72, 128, generator=g)
    pxd, scd = px.to(device), screen.to(device)

    def run():
        with torch.no_grad():
            hm = batch_make_heatmaps(pxd, cfg.gaze_heatmap_sigma_initial)
            out, _, _ = net.sequence(scd, hm)
            return soft_argmax(out.reshape(T, 1, 72, 128))

    for _ in range(3):
        run()
    torch.cuda.synchronize(device)
    ms = _event_ms(run, 5, device)
    n = 45
    osd = {'refine_net.' + k: v for k, v in sd.items()}
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    with torch.no_grad():
        hm = O.make_heatmaps(px[:, :n], cfg.gaze_heatmap_sigma_initial)
        O.refine_net_sequence(osd, cfg, screen[:, :4], hm[:, :4])
        t0 = time.perf_counter()
        O.soft_argmax(O.refine_net_sequence(osd, cfg, screen[:, :n], hm).reshape(n, 1, 72, 128))
        dt = time.perf_counter() - t0
    return {'workload': 'inference-only RefineNet stream, seq_len=900, batch=1 (BASELINE configs[4])',
            'value': T / (ms * 1e-3), 'unit': 'screen frames/s', 'ms_per_stream': ms,
            'cpu_baseline': {'value': n / dt, 'unit': 'screen frames/s', 'cores': threads,
                             'kind': 'port', 'sample': 'first %d frames of the stream, %.2f s' % (n, dt)}}


def leg_stock_torch(args, device):
    """The "existing Blackwell library path" (BASELINE.md section 3): the same training step written
    in plain PyTorch ops (the oracle's time-batched restatement of the reference modules) running
    on the B200 through stock cuDNN / cuBLAS, fp32 and TF32.  A baseline beside the product."""
    import numpy as np
    import torch
    from eve_b200 import synth
    from oracle import eve_oracle as O      # baseline leg only
    cfg = configure(args.workload)
    B, T = args.batch, args.seq_len
    sd = {k: v.to(device).requires_grad_(True) for k, v in build_state_dict(cfg).items()}
    # lr = 0: the same work per step with the parameters staying at their seeded values (sixteen
    # steps at the reference's learning rate from random weights can leave the sigmoid output
    # non-finite, which trips a device-side assert inside F.binary_cross_entropy)
    opt = torch.optim.Adam(list(sd.values()), lr=0.0, weight_decay=0.0)
    inputs = {k: v.to(device) for k, v in synth.make_clip_batch(
        B, T, seed=7, with_screen=bool(cfg.load_screen_content)).items()}
    std = np.radians(cfg.refine_net_offset_augmentation_sigma)
    kap = {s_: torch.from_numpy(np.random.normal(size=(B, 2), scale=std).astype(np.float32)).to(device)
           for s_ in ('left', 'right')}

    def step():
        opt.zero_grad(set_to_none=True)
        out, _ = O.eve_forward(sd, cfg, inputs, True, kap)
        out['full_loss'].backward()
        torch.nn.utils.clip_grad_norm_(list(sd.values()), cfg.gradient_clip_amount)
        opt.step()

    res = {'workload': WORKLOADS[args.workload]['desc'],
           'what': 'plain PyTorch ops (torch %s, cuDNN %s) on the same B200, same B=%d x T=%d step '
                   '(fwd + bwd + clip + Adam), time-batched like the product'
                   % (torch.__version__, torch.backends.cudnn.version(), B, T), 'unit': UNIT}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32,
           torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device=device)
    try:
        for name, tf32 in (('fp32', False), ('tf32', True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            for _ in range(3):
                step()
            torch.cuda.synchronize(device)
            ms = _event_ms(step, 5, device, flush)
            res[name] = {'value': 2 * B * T / (ms * 1e-3), 'ms_per_step': ms}
    finally:
        (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32,
         torch.backends.cudnn.benchmark) = old
    return res


LEGS = {'t60': leg_t60, 'stream900': leg_stream900, 'stock_torch_b200': leg_stock_torch}


def run_leg_in_child(name, args, timeout=900):
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), '--leg', name, '--workload', args.workload,
           '--batch', str(args.batch), '--seq-len', str(args.seq_len)]
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout,
                           cwd=REPO, text=True)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
        if r.returncode == 0 and lines:
            return json.loads(lines[-1])
        return {'error': 'rc=%d: %s' % (r.returncode, r.stderr.strip().splitlines()[-1][:300]
                                        if r.stderr.strip() else 'no output')}
    except Exception as e:
        return {'error': '%s: %s' % (type(e).__name__, e)}


def leg_main(args):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the B200 path has no CPU fallback')
    torch.cuda.set_device(0)
    from eve_b200 import lib as L
    L.load()
    _emit(LEGS[args.leg](args, torch.device('cuda', 0)))


def b200_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world != args.gpus and world > 1:
        raise SystemExit('bench.py: --gpus %d but WORLD_SIZE=%d' % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit('bench.py: launch N>1 with torch.distributed.run (one rank per GPU)')
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the B200 path has no CPU fallback')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)

    from eve_b200 import lib as L
    from eve_b200.models import EVE
    from eve_b200.parallel import FlatAdamTrainer
    import numpy as np
    L.load()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except Exception:
        pass

    def measure(workload, steps, warmup, with_e2e, profile):
        from eve_b200.graph import GraphedTrainStep
        lib = L.load()
        cfg = configure(workload)
        np.random.seed(1234 + rank)        # kappa augmentation draws (eve.py:468-469)
        model = EVE()
        model.load_state_dict(build_state_dict(cfg), strict=True)
        model = model.to(device).train()
        trainer = FlatAdamTrainer(model)
        B, T = args.batch, args.seq_len
        nb = 2                              # two distinct synthetic batches, alternated
        host = [make_batch(B, T, cfg, seed=1000 * rank + i, pinned=True) for i in range(nb)]
        dev = [{k: v.to(device) for k, v in h.items()} for h in host]
        res = {}
        # W eager warm-up steps, then the step is captured once as a CUDA graph
        n0 = lib.eve_launch_count()
        step_fn = GraphedTrainStep(model, trainer, dev[0], warmup=max(warmup, 3), tag='bench')
        launches_per_step = (lib.eve_launch_count() - n0) // (max(warmup, 3) + 1)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ms, loss = timed_graph_run(step_fn, lambda i: dev[i % nb], steps, 1, world, device, False)
        clocks = sampler.stop() if rank == 0 else None
        frames = 2 * B * T * world * steps
        res.update(ms=ms, launches=launches_per_step * steps, loss=loss, clocks=clocks,
                   value=frames / (ms * 1e-3), frames_per_step=2 * B * T * world)
        if with_e2e:
            # same call, HOST (pinned) inputs: H2D of the batch + D2H of the loss every step
            h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
            ms2, _ = timed_graph_run(step_fn, lambda i: host[i % nb], steps, 1, world, device, True,
                                     prefetch=True)
            ms3, _ = timed_graph_run(step_fn, lambda i: host[i % nb], steps, 1, world, device, True)
            # the same step fed with the frames as the decoder leaves them (uint8, N x H x W x C;
            # eve_b200/input_pipeline.py converts them on the device): 4x fewer bytes over PCIe
            raw = []
            for i in range(nb):
                g = torch.Generator().manual_seed(77 + 1000 * rank + i)
                r = {k: v for k, v in host[i].items()
                     if k not in ('left_eye_patch', 'right_eye_patch', 'screen_frame')}
                r['eyes_frames'] = torch.randint(0, 256, (B, T, 128, 256, 3), generator=g,
                                                 dtype=torch.uint8).pin_memory()
                if 'screen_frame' in host[i]:
                    r['screen_frames'] = torch.randint(0, 256, (B, T, 72, 128, 3), generator=g,
                                                       dtype=torch.uint8).pin_memory()
                raw.append(r)
            raw_bytes = sum(v.numel() * v.element_size() for v in raw[0].values())
            ms4, _ = timed_graph_run(step_fn, lambda i: raw[i % nb], steps, 1, world, device, True,
                                     prefetch=True)
            res['e2e'] = {'value': frames / (ms2 * 1e-3), 'unit': UNIT,
                          'h2d_bytes_per_step': int(h2d_bytes), 'd2h_bytes_per_step': 4,
                          'ms_per_step': ms2 / steps,
                          'input_pipeline': 'pinned host batch of step i+1 copied H2D on a second '
                                            'stream while step i computes (GraphedTrainStep.prefetch); '
                                            'loss read back D2H every step',
                          'unpipelined': {'value': frames / (ms3 * 1e-3), 'ms_per_step': ms3 / steps,
                                          'note': 'H2D copy serialised in front of every step'},
                          'uint8_frames': {'value': frames / (ms4 * 1e-3), 'ms_per_step': ms4 / steps,
                                           'h2d_bytes_per_step': int(raw_bytes),
                                           'note': 'eye / screen frames handed over as uint8 in the '
                                                   'decoder layout and converted on the device '
                                                   '(eve_preprocess_frames), same prefetch pipeline'}}
        step_fn.close()
        if profile:
            pms, prof = profiled_eager_run(model, trainer, lambda i: dev[i % nb], steps, world,
                                           device)
            res['prof'] = prof
            res['prof_ms'] = pms
        res['cfg'] = cfg
        return res

    main = measure(args.workload, args.steps, args.warmup, not args.no_e2e, True)
    cfg = main['cfg']
    extra = None
    if not args.no_extra:
        other = 'eyenet_static' if args.workload == 'eve_refine' else 'eve_refine'
        torch.cuda.empty_cache()
        ex_steps = min(args.steps, 20)
        ex = measure(other, ex_steps, args.warmup, False, False)
        extra = {'other_workload': {'workload': WORKLOADS[other]['desc'], 'value': ex['value'],
                                    'unit': UNIT, 'ms_per_step': ex['ms'] / ex_steps}}
        torch.cuda.empty_cache()
        if world == 1 and rank == 0:
            # each secondary leg in a child process: a failure there (even a device-side assert,
            # which poisons the CUDA context) must never take the headline down
            torch.cuda.synchronize(device)
            for name in LEGS:
                extra[name] = run_leg_in_child(name, args)
    cfg = configure(args.workload)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (implicit-GEMM convolutions)
    prof = main['prof']
    prof_steps = args.steps
    conv_ms = sum(p['ms'] for p in prof.values())
    conv_flops = sum(p['flops'] for p in prof.values())
    conv_launches = sum(p['launches'] for p in prof.values())
    peak = peaks.get('bf16_tflops_sustained')
    peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)'
    if peak is None:
        peak, peak_src = 1400.0, 'fallback (B200_PROFILING.md sustained figure)'
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    # DRAM bytes per conv launch, from the committed ncu pass over this same command
    # (dram__bytes_read.sum + dram__bytes_write.sum of every conv kernel of one step / launches)
    traffic, traffic_src = None, None
    tpath = os.path.join(REPO, 'profiles', 'conv_traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get('workload') == args.workload:
            traffic = tj.get('dram_bytes_per_conv_launch')
            traffic_src = tj.get('source')
    roofline = {
        'bound': 'tensor', 'kernel': 'conv_tc_kernel / conv_tc_strip_kernel / conv_tc_row_kernel / '
                                     'conv_tc_wgrad_kernel / conv_tc_wgrad_strip_kernel / '
                                     'conv_tc_wgrad_row_kernel / conv_tc_wgrad_row3_kernel (tcgen05 implicit GEMM, split '
                                     'fp16/bf16 operands = 2-3 MMAs per product) + the few CUDA-core '
                                     'convs left; all conv fwd + dgrad + wgrad launches',
        'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
        'peak_source': peak_src, 'traffic': traffic, 'traffic_source': traffic_src,
        'launches': conv_launches, 'avg_launch_ms': conv_ms / max(conv_launches, 1),
        # conv kernel time (events around each launch) over the graph-replayed step it is part of; the
        # eager step the events were taken in is host-bound (launch gaps), so its share reads lower
        'share_of_step': conv_ms / main['ms'],
        'share_of_eager_step': conv_ms / main['prof_ms'],
        'timed_in': 'the same K steps launched eagerly (kernel-by-kernel, events on the launching '
                    'stream) right after the graph-replayed timed region; kernels are identical',
        'eager_ms_per_step': main['prof_ms'] / prof_steps,
        'per_kind': {k: {'ms_per_step': v['ms'] / args.steps,
                         'tflops': (v['flops'] / (v['ms'] * 1e-3) / 1e12) if v['ms'] > 0 else 0.0,
                         'launches_per_step': v['launches'] / args.steps}
                     for k, v in prof.items()},
        'step_algorithmic_tflop': (FLOP_EYE_FRAME * 2 + (FLOP_SCREEN_FRAME if cfg.refine_net_enabled
                                                         else 0.0)) * args.batch * args.seq_len / 1e12,
    }
    out = {
        'metric': METRIC, 'value': main['value'], 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': main['ms'] / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': WORKLOADS[args.workload]['desc'],
                   'global_batch': args.batch * world, 'seq_len': args.seq_len,
                   'eye_frames_per_step': main['frames_per_step'],
                   'parallelism': 'dp%d (clips sharded on the batch axis, one NCCL allreduce of '
                                  'the flat fp32 gradient buffer)' % world,
                   'step': 'EVE.forward + full_loss.backward + clip_grad_norm + Adam, captured once '
                           'as a CUDA graph and replayed',
                   'conv_mode': 'tcgen05 split operands (fp16 hi+lo forward, bf16 hi+lo gradients)',
                   'cache': 'L2 flushed (160 MB write) between timed iterations; activations '
                            '(>8 GB/step) exceed L2',
                   'weights': 'random init (seeded), reference architecture'},
        'e2e': main.get('e2e'), 'gpu_launches': main['launches'], 'clocks': main['clocks'],
        'roofline': roofline, 'final_loss': main['loss'],
    }
    if extra is not None:
        out['also'] = extra
    if world == 1 and not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_baseline(cfg, build_state_dict(cfg), args.batch, args.seq_len)
    _emit(out)
    if world > 1:
        dist.destroy_process_group()


_RECORD_OUT = None


def _emit(record):
    """The one JSON line of a run, on the process's original stdout."""
    out = _RECORD_OUT if _RECORD_OUT is not None else sys.stdout
    out.write(json.dumps(record) + '\n')
    out.flush()


def main():
    args = parse()
    # stdout carries exactly ONE line, the JSON record: everything else a library writes to file
    # descriptor 1 (the NCCL version banner under torchrun, for one) is sent to stderr instead
    global _RECORD_OUT
    sys.stdout.flush()
    _RECORD_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        reference_arm(args)
    elif args.leg:
        leg_main(args)
    else:
        b200_arm(args)


if __name__ == '__main__':
    main()
