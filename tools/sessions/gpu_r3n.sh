#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_contract_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -1
echo "== ring kernel (role loops), guided"; timeout 200 python tools/cta_stats.py 2>&1 | tail -2
echo "== ring kernel (role loops), streamk partition"; ITB_SCHED=streamk timeout 200 python tools/cta_stats.py 2>&1 | tail -2
echo "== r1 kernel, streamk partition"; ITB_SCHED=streamk ITB_TILE_KERNEL=static timeout 200 python tools/cta_stats.py 2>&1 | tail -2
timeout 200 python tools/tile_probe.py | grep -E "step [14]:|TOTAL|full      whole|edge>=96  whole" | cut -c1-200
for f in 2 3; do
ITB_GUIDED_FACTOR=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/n_bench_f$f.json 2> $OUT/n_bench_f$f.err
done
python - <<'PY'
import json
for f in (2,3):
    d=json.loads(open("gpurun_out/n_bench_f%d.json"%f).read().strip().split("\n")[-1]); r=d["roofline"]
    print("guided f=%d: value %.2f ms %.3f frac %.3f tile_ms %.3f"%(f,d["value"],d["ms_per_step"],r["frac"],r["ms_per_step"]["tile_kernel"]))
PY
