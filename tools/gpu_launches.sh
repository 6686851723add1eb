#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02c}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/${TAG}_launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size") if "Grid Size" in hdr else None
for r in rows[1:40]:
    print(r[ki][:60], r[gi] if gi is not None else "", r[vi])
PY
ITB_FORCE_CFG=0 timeout 300 python tools/tile_calib.py 2>&1 | grep -v edge
