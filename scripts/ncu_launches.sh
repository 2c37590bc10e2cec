#!/bin/bash
# per-launch durations (ncu --metrics gpu__time_duration.sum) of a python snippet; prints a per-kernel summary
# usage: ncu_launches.sh TAG COUNT 'python code'
TAG=$1; CNT=$2; shift 2
ncu --metrics gpu__time_duration.sum --clock-control none -c $CNT --csv --log-file gpurun_out/${TAG}_launches.csv python -c "$1" > gpurun_out/${TAG}_ncu.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/${TAG}_launches.csv')) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
H = rows[hdr]
kn, mv = H.index('Kernel Name'), H.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    try:
        agg[r[kn][:70]].append(float(r[mv].replace(',', '')))
    except Exception:
        pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:70s} n={len(v):4d} mean={sum(v)/len(v)/1e3:9.2f} us total={sum(v)/1e6:9.3f} ms")
PY
