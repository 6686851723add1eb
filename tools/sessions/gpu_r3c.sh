#!/bin/bash
# round-2 session C: fixed epilogue; per-item anatomy; seeded config-2 parity; 2-GPU bench is a separate call
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/c_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests/test_contract_gpu.py tests/test_golden.py tests/test_permute_blas1_gpu.py -m gpu -x -q > $OUT/c_pytest1.log 2>&1; echo "pytest1 rc=$?"; tail -3 $OUT/c_pytest1.log
for f in 1 2; do
  ITB_GUIDED_FACTOR=$f ITB_MIN_PIECE=16 timeout 200 python tools/tile_probe.py >> $OUT/c_probe.txt 2>> $OUT/c_probe.err
done
cat $OUT/c_probe.txt
timeout 900 python -m pytest tests/test_plugin_dmrg.py -m gpu -x -q -k "config2 or writedim" -s > $OUT/c_pytest2.log 2>&1; echo "pytest2 rc=$?"; grep -E "config 2|passed|failed|assert" $OUT/c_pytest2.log | head
ITB_GUIDED_FACTOR=2 ITB_MIN_PIECE=16 timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/c_bench.json 2> $OUT/c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c_bench.json").read().strip().split("\n")[-1]); r=d["roofline"]; p=d["permute"]
print("value %.2f ms %.3f frac %.3f tile_ms %.3f stream_ms %.3f e2e %.2f perm %.0f GB/s (%.3f) acc %.3f"%(d["value"],d["ms_per_step"],r["frac"],r["ms_per_step"]["tile_kernel"],r["ms_per_step"]["streaming_kernel"],d["e2e"]["value"],p["achieved_gbs"],p["frac"],p["accumulate"]["frac"]))
print("cpu", d["cpu_baseline"]); print("parity", d["parity_vs_reference"]); print("plugin", d["e2e_plugin"])
PY
