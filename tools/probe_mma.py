"""tcgen05.mma issue rate from shared-memory operands (M=128, K=16, bf16): cycles per MMA as a
function of N, with one CTA per SM (all 148 busy) and with a single CTA; aligned and pixel-shifted
A start addresses.  Usage: python tools/probe_mma.py"""
import sys

import torch

sys.path.insert(0, '.')
from eve_b200 import lib as L   # noqa: E402

lib = L.load()
out = torch.zeros(148, dtype=torch.int64, device='cuda')
print('%4s %5s %6s %8s %10s %10s' % ('N', 'grid', 'shift', 'distinct', 'cyc/mma', 'floor N/2'))
for grid in (148, 1):
    for n in (16, 32, 64, 128, 256):
        for shift, distinct in ((0, 0), (0, 1), (128 * 35, 1)):
            nmma, reps = 96, 50
            L.check(lib.eve_probe_mma_rate(n, nmma, reps, shift, distinct, grid, L.ptr(out), L.stream_ptr()),
                    'probe')
            torch.cuda.synchronize()
            cyc = out[:grid].double().mean().item() / (nmma * reps)
            print('%4d %5d %6d %8d %10.2f %10.1f' % (n, grid, shift, distinct, cyc, n / 2))
print()
print('operand row width (swizzle mode), 148 CTAs, aligned start')
print('%4s %10s %10s' % ('N', 'row bytes', 'cyc/mma'))
for n in (16, 32, 64, 128):
    for rb in (128, 64, 32):
        nmma, reps = 96, 50
        L.check(lib.eve_probe_mma_rate_swizzle(n, nmma, reps, 0, 1, rb, 148, L.ptr(out), L.stream_ptr()), 'probe')
        torch.cuda.synchronize()
        print('%4d %10d %10.2f' % (n, rb, out.double().mean().item() / (nmma * reps)))
print()
print('concurrent issuing warps (independent chains), 148 CTAs: cycles per MMA of one chain / aggregate')
print('%4s %8s %10s %10s' % ('N', 'issuers', 'per chain', 'aggregate'))
for n in (16, 32, 64, 128):
    for iss in (1, 2, 4):
        nmma, reps = 96, 50
        L.check(lib.eve_probe_mma_rate_issuers(n, nmma, reps, 0, 1, 128, iss, 148, L.ptr(out), L.stream_ptr()), 'probe')
        torch.cuda.synchronize()
        c = out.double().mean().item() / (nmma * reps)
        print('%4d %8d %10.2f %10.2f' % (n, iss, c, c / iss))

print()
print('CTA pairs (cta_group::2, M = 256 over two SMs, each CTA holds 128 rows of A and N/2 rows of B):')
print('cycles per instruction = cycles each SM spends per M=128 x N x K=16 of work')
print('%4s %6s %10s %14s' % ('N', 'pairs', 'cyc/mma', '1-CTA M=128'))
single = {32: 43.7, 64: 51.4, 128: 68.3, 256: 132.2}
for pairs in (74, 1):
    for n in (32, 64, 128, 256):
        nmma, reps = 96, 50
        L.check(lib.eve_probe_mma_rate_pair(n, nmma, reps, pairs, L.ptr(out), L.stream_ptr()), 'probe')
        torch.cuda.synchronize()
        print('%4d %6d %10.2f %14.1f' % (n, pairs, out[:pairs].double().mean().item() / (nmma * reps), single[n]))
