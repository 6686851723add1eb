#!/bin/bash
# round-2 session B: metadata ring + fast epilogue; planner knob sweep; ncu capture; seeded config-2 parity
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/b_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests/test_contract_gpu.py tests/test_golden.py -m gpu -x -q > $OUT/b_pytest1.log 2>&1; echo "pytest1 rc=$?"; tail -3 $OUT/b_pytest1.log
for f in 1 2 3 4; do for mp in 8 16; do
  ITB_GUIDED_FACTOR=$f ITB_MIN_PIECE=$mp timeout 200 python tools/tile_probe.py >> $OUT/b_probe.txt 2>> $OUT/b_probe.err
done; done
grep -E "TOTAL|step" $OUT/b_probe.txt
timeout 900 python -m pytest tests/test_plugin_dmrg.py -m gpu -x -q -k "config2 or writedim" -s > $OUT/b_pytest2.log 2>&1; echo "pytest2 rc=$?"; grep -E "config 2|passed|failed|assert" $OUT/b_pytest2.log | head
ITB_GUIDED_FACTOR=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsc_gemm_kernel -s 4 -c 2 -o $OUT/b_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/b_ncu.log 2>&1; echo "ncu rc=$?"
ITB_GUIDED_FACTOR=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/b_bench.json 2> $OUT/b_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b_bench.json").read().strip().split("\n")[-1]); r=d["roofline"]; p=d["permute"]
print("value %.2f ms %.3f frac %.3f tile_ms %.3f stream_ms %.3f e2e %.2f perm %.0f GB/s (%.3f)"%(d["value"],d["ms_per_step"],r["frac"],r["ms_per_step"]["tile_kernel"],r["ms_per_step"]["streaming_kernel"],d["e2e"]["value"],p["achieved_gbs"],p["frac"]))
PY
