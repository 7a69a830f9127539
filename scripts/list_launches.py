"""Ordered per-launch dump (name, grid, block, us) from an ncu gpu__time_duration csv."""
import csv, re, sys
with open(sys.argv[1], newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'^void ', '', re.sub(r'\(.*$', '', row['Kernel Name']))
    v = float(row['Metric Value'].replace(',', ''))
    us = v / 1000.0 if row['Metric Unit'] in ('ns', 'nsecond') else v
    print('%4s %-60s grid=%-18s block=%-12s %9.1f us' % (row['ID'], name[:60], row['Grid Size'], row['Block Size'], us))
