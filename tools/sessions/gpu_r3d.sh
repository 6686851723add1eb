#!/bin/bash
# round-2 session D: non-blocking look-ahead; config-2 diagnostics; eigh batch test
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/d_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests/test_contract_gpu.py tests/test_golden.py tests/test_solver_gpu.py -m gpu -x -q > $OUT/d_pytest1.log 2>&1; echo "pytest1 rc=$?"; tail -3 $OUT/d_pytest1.log
for f in 1 2 3; do
  ITB_GUIDED_FACTOR=$f ITB_MIN_PIECE=8 timeout 200 python tools/tile_probe.py >> $OUT/d_probe.txt 2>> $OUT/d_probe.err
done
cat $OUT/d_probe.txt
timeout 1200 python tools/config2_diag.py 800 > $OUT/d_diag.txt 2>&1; echo "diag rc=$?"; cat $OUT/d_diag.txt | tail -8
