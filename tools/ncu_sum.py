import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
hdr = r[0]
want = ['Kernel Name', 'launch__grid_size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_subpipe_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum']
idx = [(w, hdr.index(w)) for w in want if w in hdr]
tens = [(h, i) for i, h in enumerate(hdr) if 'tensor' in h.lower()]
for row in r[2:]:
    print('---')
    for w, i in idx:
        print('  %-75s %s %s' % (w, row[i], r[1][i]))
    for h, i in tens[:12]:
        if (h, i) not in idx:
            print('  %-75s %s %s' % (h, row[i], r[1][i]))
