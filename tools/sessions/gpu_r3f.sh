#!/bin/bash
# round-2 session F: hybrid static+dynamic schedule variants
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/f_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests/test_contract_gpu.py tests/test_golden.py -m gpu -x -q > $OUT/f_pytest1.log 2>&1; echo "pytest1 rc=$?"; tail -3 $OUT/f_pytest1.log
for v in "0.85 1" "0.85 2" "0.92 1" "0 2" "0.7 1"; do set -- $v
  ITB_STATIC_FRAC=$1 ITB_GUIDED_FACTOR=$2 timeout 200 python tools/tile_probe.py | sed "s/^/static=$1 /" >> $OUT/f_probe.txt 2>> $OUT/f_probe.err
done
grep -E "step [14]:|TOTAL" $OUT/f_probe.txt | cut -c1-260
