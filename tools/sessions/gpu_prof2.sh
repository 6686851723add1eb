#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsc_skinny_smallk -s 8 -c 1 -o $OUT/r01m_skinny python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/r01m_skinny.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:perm_tile -s 2 -c 1 -o $OUT/r01m_perm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/r01m_perm.log 2>&1
timeout 900 python tools/synth_sweep.py --ms 512,1024,2048,4096 --out $OUT/synth_sweep.jsonl 2>&1 | tail -30
