#!/bin/bash
# per-stage durations of one keyswitch call (N=16384, D/K=7/8, batch 1024) from an ncu launch list
#   usage: bash tools/ks_stage_times.sh tag [ENVVAR=..]
TAG=$1
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ks_ --csv --log-file gpurun_out/r2_ks_launches_$TAG.csv python tools/prof_target.py keyswitch 1 1024 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_ks_launches_$TAG.csv")) if len(r)>10 and r[0].isdigit()]
rows=rows[len(rows)//2:]
tot={}
for r in rows:
    name=r[4].split('(')[0].replace('void ','')
    tot.setdefault(name,[]).append(int(r[-1]))
s=0
for k,v in tot.items():
    print(f"{sum(v)/1000:9.1f} us  {k[:70]}  {v}")
    s+=sum(v)
print("sum us", s/1000, "per item us", s/1000/1024)
PY
