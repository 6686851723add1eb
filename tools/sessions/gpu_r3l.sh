#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_contract_gpu.py -m gpu -x -q 2>&1 | tail -2
echo "== current (row4 predicated), guided"; timeout 200 python tools/cta_stats.py 2>&1 | tail -3
timeout 200 python tools/tile_probe.py | grep -E "step [14]:|TOTAL|full      whole|edge>=96  whole" | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/l_bench.json 2> $OUT/l_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/l_bench.json").read().strip().split("\n")[-1]); r=d["roofline"]; p=d["permute"]
print("value %.2f ms %.3f frac %.3f tile_ms %.3f"%(d["value"],d["ms_per_step"],r["frac"],r["ms_per_step"]["tile_kernel"]))
PY
