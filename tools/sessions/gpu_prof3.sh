#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02d}
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 1500 $OUT/${TAG}_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsc_gemm -s 6 -c 2 -o $OUT/${TAG}_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ITB_FORCE_CFG=0 timeout 300 python tools/tile_calib.py 2>&1 | grep -v edge | grep "=>"
