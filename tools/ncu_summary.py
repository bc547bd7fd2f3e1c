"""Key metrics of every kernel in an .ncu-rep (run here, no GPU needed): python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
    print(d['Kernel Name'][0][:110])
    for k in KEYS:
        if k in d:
            print('   %-66s %s %s' % (k, d[k][0], d[k][1]))
    print()
