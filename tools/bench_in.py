"""Micro-benchmark of the InstanceNorm kernels at the bench geometry (B=8, T=30): one-pass cluster
kernels (in_fused.cu) against the separate stats / apply / reduce passes, in achieved GB/s of the
algorithmic bytes each moves.  Usage: python tools/bench_in.py [fwd|bwd|all]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, '.')
from eve_b200 import lib as L  # noqa: E402

lib = L.load()
SHAPES = [  # (n, hw, c) RefineNet (240 frames) and EyeNet (480 patches)
    (240, 9216, 16), (240, 9216, 32), (240, 9216, 64), (240, 2304, 32), (240, 2304, 64),
    (240, 2304, 128), (240, 576, 64), (240, 576, 128), (240, 576, 256), (240, 144, 128),
    (240, 144, 256), (240, 144, 512), (240, 40, 256), (240, 40, 64), (480, 1024, 64),
    (480, 256, 128), (480, 64, 256), (480, 16, 512)]


def timeit(fn, iters=5):
    flush = torch.empty(160 * 1024 * 1024 // 4, device='cuda')
    fn()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    return ms / iters


def u16(n):
    return torch.empty(n, dtype=torch.int16, device='cuda')


import os
if os.environ.get('BENCH_IN_SHAPES'):
    SHAPES = [tuple(int(v) for v in t.split('x')) for t in os.environ['BENCH_IN_SHAPES'].split(',')]


def main(which):
    s = L.stream_ptr()
    if os.environ.get('BENCH_IN_STREAM'):
        L.set_option('in_stream', int(os.environ['BENCH_IN_STREAM']))
    chunk_mb = float(os.environ.get('BENCH_IN_CHUNK_MB', '0'))
    print('%-18s %10s %10s %10s %10s %12s %8s' % ('shape', 'fused ms', 'GB/s', 'legacy ms', 'GB/s',
                                                   'chunked ms', 'GB/s'))
    tot_f = tot_l = 0.0
    for n, hw, c in SHAPES:
        e = n * hw * c
        x = torch.randn(e, device='cuda')
        dy = torch.randn(e, device='cuda')
        y = torch.empty(e, device='cuda')
        hi, lo, hi2, lo2 = u16(e), u16(e), u16(e), u16(e)
        mean, rstd = torch.zeros(n * c, device='cuda'), torch.ones(n * c, device='cuda')
        g, b = torch.ones(c, device='cuda'), torch.zeros(c, device='cuda')
        dg, db, dbias = (torch.empty(c, device='cuda') for _ in range(3))
        ws = torch.empty(lib.eve_instnorm_fused_workspace_bytes(n, hw, c) + 2 * n * c * 4 + 256,
                         dtype=torch.uint8, device='cuda')
        if which in ('fwd', 'all'):
            def fused():
                L.check(lib.eve_instnorm_fused_fwd(L.ptr(x), None, 0, n, hw, c, L.ptr(g), L.ptr(b),
                                                   None, None, 1, 0, L.ptr(mean), L.ptr(rstd), None,
                                                   None, None, L.ptr(hi), L.ptr(lo), None, None, s), 'f')

            def legacy():
                L.check(lib.eve_instnorm_act_fwd(L.ptr(x), n, hw, c, L.ptr(g), L.ptr(b), 1, L.ptr(y),
                                                 L.ptr(mean), L.ptr(rstd), s), 'l')
            # the separate stats / apply passes over chunks of images small enough that the second
            # pass finds its input in L2 (BENCH_IN_CHUNK_MB of x per chunk)
            per = max(1, int(chunk_mb * 1e6 / (hw * c * 4))) if chunk_mb > 0 else n

            def chunked():
                for i0 in range(0, n, per):
                    m = min(per, n - i0)
                    o, so = i0 * hw * c, i0 * c
                    L.check(lib.eve_instnorm_act_fwd(x[o:].data_ptr(), m, hw, c, L.ptr(g), L.ptr(b), 1,
                                                     y[o:].data_ptr(), mean[so:].data_ptr(),
                                                     rstd[so:].data_ptr(), s), 'c')
            tf, tl = timeit(fused), timeit(legacy)
            tc = timeit(chunked) if chunk_mb > 0 else float('nan')
            bytes_ = 8.0 * e
            print('fwd %-14s %10.3f %10.0f %10.3f %10.0f %12.3f %8.0f' % ('%dx%dx%d' % (n, hw, c), tf,
                  bytes_ / tf / 1e6, tl, bytes_ / tl / 1e6, tc, bytes_ / tc / 1e6))
            tot_f += tf
            tot_l += tl
        if which in ('bwd', 'all'):
            def fusedb():
                L.check(lib.eve_instnorm_fused_bwd(L.ptr(dy), None, None, L.ptr(x), n, hw, c,
                                                   L.ptr(mean), L.ptr(rstd), L.ptr(g), L.ptr(b), None,
                                                   None, 1, None, None, L.ptr(hi), L.ptr(lo), None,
                                                   L.ptr(dg), L.ptr(db), None, None, L.ptr(dbias),
                                                   L.ptr(ws), ws.numel(), s), 'fb')

            def legacyb():
                L.check(lib.eve_instnorm_act_bwd(L.ptr(dy), L.ptr(x), L.ptr(x), n, hw, c, L.ptr(mean),
                                                 L.ptr(rstd), L.ptr(g), 1, L.ptr(y), L.ptr(dg),
                                                 L.ptr(db), L.ptr(ws), ws.numel(), s), 'lb')
            per = max(1, int(chunk_mb * 1e6 / (hw * c * 8))) if chunk_mb > 0 else n

            def chunkedb():
                for i0 in range(0, n, per):
                    m = min(per, n - i0)
                    o, so = i0 * hw * c, i0 * c
                    L.check(lib.eve_instnorm_act_bwd(dy[o:].data_ptr(), x[o:].data_ptr(), x[o:].data_ptr(),
                                                     m, hw, c, mean[so:].data_ptr(), rstd[so:].data_ptr(),
                                                     L.ptr(g), 1, y[o:].data_ptr(), L.ptr(dg), L.ptr(db),
                                                     L.ptr(ws), ws.numel(), s), 'cb')
            tf, tl = timeit(fusedb), timeit(legacyb)
            tc = timeit(chunkedb) if chunk_mb > 0 else float('nan')
            bytes_ = 12.0 * e
            print('bwd %-14s %10.3f %10.0f %10.3f %10.0f %12.3f %8.0f' % ('%dx%dx%d' % (n, hw, c), tf,
                  bytes_ / tf / 1e6, tl, bytes_ / tl / 1e6, tc, bytes_ / tc / 1e6))
            tot_f += tf
            tot_l += tl
    print('total fused %.3f ms, legacy %.3f ms' % (tot_f, tot_l))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'all')
