#!/bin/bash
# round-2 session I (1 GPU): round-1 snapshot vs current build on the SAME box; whole GPU test suite
OUT=gpurun_out; mkdir -p $OUT
(cd build/r1snap && timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ../../$OUT/i_bench_r1.json 2> ../../$OUT/i_bench_r1.err); echo "r1 bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/i_bench_now.json 2> $OUT/i_bench_now.err; echo "bench rc=$?"
(cd build/r1snap && timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ../../$OUT/i_bench_r1b.json 2> ../../$OUT/i_bench_r1b.err)
python - <<'PY'
import json
for f in ("i_bench_r1","i_bench_now","i_bench_r1b"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().split("\n")[-1]); r=d["roofline"]; p=d["permute"]
    print(f,"value %.2f ms %.3f frac %.3f tile_ms %.3f stream_ms %.3f e2e %.2f perm %.0f GB/s (%.3f) clocks %s"%(d["value"],d["ms_per_step"],r["frac"],r["ms_per_step"]["tile_kernel"],r["ms_per_step"]["streaming_kernel"],d["e2e"]["value"],p["achieved_gbs"],p["frac"],d["clocks"]["sm_mhz"]))
d=json.loads(open("gpurun_out/i_bench_now.json").read().strip().split("\n")[-1]); print("plugin", d["e2e_plugin"]); print("cpu", {k:v["tflops"] for k,v in d["cpu_baseline"]["modes"].items()}); print("parity ok", d["parity_vs_reference"]["ok"])
PY
timeout 1800 python -m pytest tests -m gpu -x -q -s > $OUT/i_pytest_all.log 2>&1; echo "pytest all rc=$?"; grep -E "config 2|passed|failed" $OUT/i_pytest_all.log | tail -5
