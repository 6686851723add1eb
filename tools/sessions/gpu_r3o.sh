#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
(cd build/r1snap && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file ../../$OUT/o_launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file $OUT/o_launches_now.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
for f in ("o_launches_r1","o_launches_now"):
    rows=[r for r in csv.reader(open("gpurun_out/%s.csv"%f)) if len(r)>10 and r[0].isdigit()]
    print("==",f,len(rows))
    for r in rows[:14]:
        name=r[4][:60]; print("  %-60s grid %-14s %s us"%(name, r[7] if len(r)>7 else "", r[-1]))
PY
