#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== round-1 snapshot"; (cd build/r1snap && timeout 200 python ../../tools/cta_stats.py) 2>&1 | tail -3
echo "== current, guided"; timeout 200 python tools/cta_stats.py 2>&1 | tail -3
echo "== current, streamk"; ITB_SCHED=streamk timeout 200 python tools/cta_stats.py 2>&1 | tail -3
