#!/bin/bash
# round-2 session E: in-kernel split-K fix-up; producer-only / consumer-only rates; config-2 diagnostics; full bench
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/e_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests/test_contract_gpu.py tests/test_golden.py -m gpu -x -q > $OUT/e_pytest1.log 2>&1; echo "pytest1 rc=$?"; tail -3 $OUT/e_pytest1.log
ITB_GUIDED_FACTOR=2 timeout 200 python tools/tile_probe.py > $OUT/e_probe.txt 2>> $OUT/e_probe.err
ITB_GUIDED_FACTOR=2 ITB_DEBUG_NOCOMPUTE=1 timeout 200 python tools/tile_probe.py | sed 's/^/NOCOMPUTE(producer only) /' >> $OUT/e_probe.txt 2>> $OUT/e_probe.err
ITB_GUIDED_FACTOR=2 ITB_DEBUG_NOCOMPUTE=2 timeout 200 python tools/tile_probe.py | sed 's/^/NOLOAD(consumer only) /' >> $OUT/e_probe.txt 2>> $OUT/e_probe.err
cat $OUT/e_probe.txt
ITB_GUIDED_FACTOR=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/e_bench.json 2> $OUT/e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/e_bench.json").read().strip().split("\n")[-1]); r=d["roofline"]; p=d["permute"]
print("value %.2f ms %.3f frac %.3f tile_ms %.3f stream_ms %.3f e2e %.2f perm %.0f GB/s (%.3f) acc %.3f launches %d"%(d["value"],d["ms_per_step"],r["frac"],r["ms_per_step"]["tile_kernel"],r["ms_per_step"]["streaming_kernel"],d["e2e"]["value"],p["achieved_gbs"],p["frac"],p["accumulate"]["frac"],d["gpu_launches"]))
PY
timeout 1200 python tools/config2_diag.py 800 > $OUT/e_diag.txt 2>&1; echo "diag rc=$?"; cat $OUT/e_diag.txt | tail -8
