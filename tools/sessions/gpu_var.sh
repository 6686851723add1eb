#!/bin/bash
for v in 0 1 2; do
  ITB_SKINNY_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('variant $v value',round(d['value'],2),'ms',round(d['ms_per_step'],3), d['roofline']['ms_per_step'], 'perm', round(d['permute']['achieved_gbs']))"
done
