#!/usr/bin/env python3
"""Per-kernel totals and shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_shares.py launches.csv"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; idx = {h: i for i, h in enumerate(H)}
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    name = r[idx['Kernel Name']].replace('void ', '').replace('<unnamed>::', '')
    name = re.sub(r'[<(].*', '', name)
    v = float(r[idx['Metric Value']].replace(',', '')); u = r[idx['Metric Unit']]
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} launches {c:4d}  total {t:9.1f} us  avg {t / c:8.1f} us  share {t / tot:.3f}")
