"""From an ncu CSV launch list with gpu__time_duration.sum, dram__bytes_read.sum and
dram__bytes_write.sum per launch: total DRAM bytes of the convolution kernels and the per-launch
average that bench.py reports as roofline.traffic.  Usage: ncu_conv_traffic.py list.csv steps out.json"""
import collections, csv, json, re, sys
path, steps, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(open(path, errors='replace')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
iid, ik, im, iu, iv = (hdr.index(k) for k in ('ID', 'Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value'))
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= iv:
        continue
    d = per.setdefault(r[iid], {'name': r[ik]})
    v = float(r[iv].replace(',', ''))
    u = r[iu].lower()
    if 'byte' in u:
        v *= {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
    if u in ('us', 'usecond'):
        v *= 1e3
    if u in ('ms', 'msecond'):
        v *= 1e6
    d[r[im]] = v
conv = re.compile(r'conv_tc|igemm_gather|igemm_wgrad')
n = b = t = 0
tot_t = 0.0
for d in per.values():
    tot_t += d.get('gpu__time_duration.sum', 0.0)
    if conv.search(d['name']):
        n += 1
        b += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
        t += d.get('gpu__time_duration.sum', 0.0)
res = {'workload': 'eve_refine', 'steps_captured': steps, 'conv_launches': n,
       'conv_dram_bytes': b, 'dram_bytes_per_conv_launch': b / max(n, 1),
       'conv_time_ms': t / 1e6, 'all_kernels_time_ms': tot_t / 1e6,
       'conv_share_of_kernel_time': t / max(tot_t, 1.0),
       'source': 'profiles/%s: ncu --profile-from-start off --metrics gpu__time_duration.sum,'
                 'dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over the eager '
                 'profiled step of `bench.py --steps 1 --warmup 3 --no-extra --no-e2e '
                 '--no-cpu-baseline` (EVE_BENCH_NCU_RANGE=1)' % out.split('/')[-1].replace('.json', '_launches.csv.gz')}
json.dump(res, open(out, 'w'), indent=1)
print(json.dumps(res, indent=1))
