#!/usr/bin/env python
"""Regions of equal execution count in the SASS page of an ncu report, with their share of the stall samples.

    python profiles/tools/sass_regions.py sass.csv units_per_launch
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) / 32
hdr = rows[1]
data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
KEYS = ['stall_wait', 'stall_math', 'stall_not_selected', 'stall_short_sb', 'stall_long_sb', 'stall_barrier',
        'stall_branch_resolving', 'stall_selected', 'stall_dispatch', 'stall_no_inst', 'stall_lg', 'stall_mio']
reg = []
cur = None
for n, r in enumerate(data):
    e = int(r[ix['Instructions Executed']])
    if cur is None or abs(cur['e'] - e) > 0.02 * max(cur['e'], 1):
        cur = {'e': e, 'n0': n, 'n1': n, 's': 0, 'st': {}, 'fp64': 0}
        reg.append(cur)
    cur['n1'] = n
    cur['s'] += int(r[ix['# Samples']])
    op = [o for o in r[ix['Source']].split() if not o.startswith('@')][0]
    if op.startswith(('DFMA', 'DMUL', 'DADD', 'DSETP')):
        cur['fp64'] += 1
    for k in KEYS:
        cur['st'][k] = cur['st'].get(k, 0) + int(r[ix[k]])
for r in reg:
    if r['s'] > 0.008 * tot:
        top = sorted(r['st'].items(), key=lambda kv: -kv[1])[:4]
        n = r['n1'] - r['n0'] + 1
        print(f"inst {r['n0']:5d}-{r['n1']:5d} exec/unit {r['e'] / units:7.4f} n={n:4d} fp64={r['fp64']:4d} "
              f"samples {100 * r['s'] / tot:5.1f}%  " + ", ".join(f"{k[6:]} {100 * v / max(r['s'], 1):.0f}%" for k, v in top))
