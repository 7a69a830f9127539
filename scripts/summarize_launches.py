"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
time and share.  usage: python scripts/summarize_launches.py launches.csv [title] > profiles/x.md"""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1], newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
r = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
order = []
for row in r:
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = row['Kernel Name']
    name = re.sub(r'\(.*$', '', name)
    name = re.sub(r'^void ', '', name)
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    us = v / 1000.0 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1000.0)
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print('# %s' % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print()
print('Per-launch times are cold-cache and serialised (ncu): compare SHARES, not absolutes.')
print()
print('total %.1f us over %d launches' % (tot, sum(v[0] for v in agg.values())))
print()
print('| kernel | launches | total us | share |')
print('|---|---:|---:|---:|')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('| `%s` | %d | %.1f | %.1f%% |' % (k[:110], n, t, 100 * t / tot))
