#!/bin/bash
# round-2 session G: clean bench numbers of the current build + the whole GPU test suite
OUT=gpurun_out; mkdir -p $OUT
for f in 2 1.5; do
ITB_GUIDED_FACTOR=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/g_bench_f$f.json 2> $OUT/g_bench_f$f.err; echo "bench f=$f rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/g_bench_*.json")):
    d=json.loads(open(f).read().strip().split("\n")[-1]); r=d["roofline"]; p=d["permute"]
    print(f,"value %.2f ms %.3f frac %.3f tile_ms %.3f stream_ms %.3f e2e %.2f perm %.0f GB/s (%.3f) acc %.3f launches %d"%(d["value"],d["ms_per_step"],r["frac"],r["ms_per_step"]["tile_kernel"],r["ms_per_step"]["streaming_kernel"],d["e2e"]["value"],p["achieved_gbs"],p["frac"],p["accumulate"]["frac"],d["gpu_launches"]))
PY
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/g_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -5 $OUT/g_pytest_all.log
