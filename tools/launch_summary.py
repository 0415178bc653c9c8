"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/launch_summary.py gpurun_out/launches.csv
"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
for i, r in enumerate(rows):
    if r[0] == 'ID':
        hdr, rows = r, rows[i + 1:]
        break
agg = collections.OrderedDict()
for r in rows:
    d = dict(zip(hdr, r))
    if d['Metric Name'] != 'gpu__time_duration.sum':
        continue
    agg.setdefault(d['Kernel Name'][:44], []).append(float(d['Metric Value'].replace(',', '')) / 1e3)
for k, v in agg.items():
    print(f"{k:46s} n={len(v):3d} last_us={v[-1]:9.1f}  first: {[round(x, 1) for x in v[:6]]}")
