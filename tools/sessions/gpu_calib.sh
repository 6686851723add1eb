#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02b}
for f in 0 1 2; do ITB_FORCE_CFG=$f timeout 300 python tools/tile_calib.py 2>&1 | tee -a $OUT/${TAG}_calib.txt; done
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 3000 $OUT/${TAG}_bench.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/${TAG}_pytest.txt
