#!/bin/bash
# per-kernel durations of the prediction step (ncu launch list) -> stdout summary
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_pred_tmp.csv python scratch/pred_only.py > gpurun_out/ncu_list.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_pred_tmp.csv')))
seen={}
for r in rows:
    if len(r)>5 and r[-3]=='gpu__time_duration.sum':
        seen.setdefault(r[4][:60],[]).append(float(r[-1])/1e6)
for k,v in seen.items(): print(f"{k:62s} n={len(v)} last={v[-1]:.3f} ms")
PY
