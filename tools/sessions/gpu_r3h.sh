#!/bin/bash
# round-2 session H: guided queue vs restored static stream-K partition on the same kernel (bench, non-profile numbers)
OUT=gpurun_out; mkdir -p $OUT
for v in "guided" "streamk"; do
ITB_SCHED=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/h_bench_$v.json 2> $OUT/h_bench_$v.err; echo "bench $v rc=$?"
ITB_SCHED=$v timeout 200 python tools/tile_probe.py | grep -E "step [14]:|TOTAL|full      whole" | sed "s/^/$v /" >> $OUT/h_probe.txt
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/h_bench_*.json")):
    d=json.loads(open(f).read().strip().split("\n")[-1]); r=d["roofline"]; p=d["permute"]
    print(f,"value %.2f ms %.3f frac %.3f tile_ms %.3f stream_ms %.3f e2e %.2f perm %.0f GB/s (%.3f) acc %.3f launches %d"%(d["value"],d["ms_per_step"],r["frac"],r["ms_per_step"]["tile_kernel"],r["ms_per_step"]["streaming_kernel"],d["e2e"]["value"],p["achieved_gbs"],p["frac"],p["accumulate"]["frac"],d["gpu_launches"]))
PY
cat $OUT/h_probe.txt | cut -c1-230
ITB_SCHED=streamk timeout 600 python -m pytest tests/test_contract_gpu.py -m gpu -x -q 2>&1 | tail -2
