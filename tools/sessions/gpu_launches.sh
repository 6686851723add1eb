#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02c}
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$OUT/${TAG}_launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size"); mi=hdr.index("Metric Name"); ii=hdr.index("ID")
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ii],[r[ki][:70],r[gi]]).append(r[mi].split("__")[1][:12]+"="+r[vi])
for k,v in list(d.items())[:24]: print(*v)
PY
