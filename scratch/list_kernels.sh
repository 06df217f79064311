#!/bin/bash
# usage: scratch/list_kernels.sh <count> <python script and args...> : per-kernel duration summary from an ncu launch list
cnt=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -c $cnt --csv --log-file gpurun_out/launches_tmp.csv python "$@" > gpurun_out/ncu_list.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_tmp.csv')))
seen={}
for r in rows:
    if len(r)>5 and r[-3]=='gpu__time_duration.sum':
        seen.setdefault(r[4][:70],[]).append(float(r[-1])/1e6)
tot=sum(sum(v) for v in seen.values())
for k,v in sorted(seen.items(), key=lambda kv:-sum(kv[1])): print(f"{k:72s} n={len(v):4d} sum={sum(v):9.3f} ms  mean={sum(v)/len(v):.4f}")
print("total", tot)
PY
