"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name
(and optionally per grid size) launch count, total time and share."""
import csv, collections, sys, re
path = sys.argv[1]
by_grid = len(sys.argv) > 2 and sys.argv[2] == '--grid'
rows = list(csv.reader(open(path, errors='replace')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
ik, ig, iv, iu = hdr.index('Kernel Name'), hdr.index('Grid Size'), hdr.index('Metric Value'), hdr.index('Metric Unit')
imn = hdr.index('Metric Name')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= iv or r[imn] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', r[ik])
    name = name.split('(')[0][:90]
    key = (name, r[ig]) if by_grid else (name,)
    v = float(r[iv].replace(',', ''))
    ms = v / 1e6 if r[iu] in ('ns', 'nsecond') else (v / 1e3 if r[iu] in ('us', 'usecond') else v)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
n = sum(a[0] for a in agg.values())
print('%-100s %8s %10s %6s %9s' % ('kernel' + (' [grid]' if by_grid else ''), 'launches', 'total_ms', 'share', 'avg_us'))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 60]:
    print('%-100s %8d %10.3f %5.1f%% %9.1f' % (' '.join(k), a[0], a[1], 100 * a[1] / tot, 1e3 * a[1] / a[0]))
print('TOTAL %d launches %.1f ms' % (n, tot))
