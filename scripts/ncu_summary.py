"""Key metrics of an `ncu --set full` report as markdown.  usage: python scripts/ncu_summary.py rep.ncu-rep "title" """
import csv, io, subprocess, sys
rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
print('# %s\n' % title)
print('`ncu --set full --clock-control none` (replayed, cold caches: bytes and utilisation are per launch; '
      'never a bench number)\n')
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print('## `%s`\n' % r[hdr.index('Kernel Name')][:100])
    print('| metric | value | unit |\n|---|---:|---|')
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print('| %s | %s | %s |' % (k, r[i], units[i]))
    try:
        U = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
        ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        tot = float(r[ir]) * U[units[ir]] + float(r[iw]) * U[units[iw]]
        print('| **traffic = dram read + write** | %.4f | Gbyte |' % (tot / 1e9))
    except Exception:
        pass
    print()
