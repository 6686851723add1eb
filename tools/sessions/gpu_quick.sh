#!/bin/bash
# quick GPU check: parity tests + bench + ncu launch list (no full capture)
TAG=${1:-q}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/${TAG}_pytest.txt
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > $OUT/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
echo done
