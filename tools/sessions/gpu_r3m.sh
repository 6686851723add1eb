#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== r1 kernel inside the current library, streamk partition"; ITB_SCHED=streamk ITB_TILE_KERNEL=static timeout 200 python tools/cta_stats.py 2>&1 | tail -2
echo "== ring kernel, streamk partition"; ITB_SCHED=streamk timeout 200 python tools/cta_stats.py 2>&1 | tail -2
echo "== ring kernel, guided"; timeout 200 python tools/cta_stats.py 2>&1 | tail -2
ITB_SCHED=streamk ITB_TILE_KERNEL=static timeout 600 python -m pytest tests/test_contract_gpu.py -m gpu -x -q 2>&1 | tail -1
ITB_SCHED=streamk ITB_TILE_KERNEL=static timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/m_bench_static.json 2> $OUT/m_bench_static.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/m_bench_static.json").read().strip().split("\n")[-1]); r=d["roofline"]
print("static kernel + streamk: value %.2f ms %.3f frac %.3f tile_ms %.3f"%(d["value"],d["ms_per_step"],r["frac"],r["ms_per_step"]["tile_kernel"]))
PY
