#!/usr/bin/env python
"""Opcode histogram (per unit) and stall-sample totals from `ncu -i X.ncu-rep --page source --csv --print-source sass`.

    python profiles/sass_hist.py sass.csv units_per_launch
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2])
hdr = rows[1]
data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot_inst = sum(int(r[ix['Instructions Executed']]) for r in data)
tot_samp = sum(int(r[ix['# Samples']]) for r in data)
print('total warp inst', tot_inst, 'samples', tot_samp, 'warp-inst per 32 units', tot_inst / (units / 32))
c = Counter()
s = Counter()
for r in data:
    op = [o for o in r[ix['Source']].split() if not o.startswith('@')][0].split('.')[0]
    c[op] += int(r[ix['Instructions Executed']])
    s[op] += int(r[ix['# Samples']])
for op, n in c.most_common(32):
    print(f"{op:10s} {n / (units / 32):8.2f} inst/unit   samples {100 * s[op] / tot_samp:5.1f}%")
for k in ['stall_barrier', 'stall_branch_resolving', 'stall_dispatch', 'stall_long_sb', 'stall_math', 'stall_mio', 'stall_no_inst',
          'stall_not_selected', 'stall_selected', 'stall_short_sb', 'stall_wait', 'stall_lg', 'stall_misc']:
    print(k, sum(int(r[ix[k]]) for r in data))
