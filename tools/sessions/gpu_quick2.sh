#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02f}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/${TAG}_pytest.txt
timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench.json')); print('value',d['value'], 'ms',d['ms_per_step'], 'frac',d['roofline']['frac'], d['roofline']['ms_per_step'], 'stream GB/s', d['roofline']['streaming_class_gbs'], 'e2e', d['e2e']['value'], 'perm', d['permute']['achieved_gbs'])"
