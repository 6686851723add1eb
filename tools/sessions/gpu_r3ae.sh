#!/bin/bash
# round-2 session AE (1 GPU): static kernel on compact tile records; both tile kernels through the parity tests; bench
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_contract_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -1
ITB_SCHED=guided timeout 600 python -m pytest tests/test_contract_gpu.py -m gpu -x -q 2>&1 | tail -1
ITB_TILE_KERNEL=ring timeout 600 python -m pytest tests/test_contract_gpu.py -m gpu -x -q -k "bench_workload or large_blocks or random_qn_pairs_larger" 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/ae_bench.json 2> $OUT/ae_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/ae_bench.json").read().strip().split("\n")[-1]); r=d["roofline"]
print("value %.2f ms %.3f frac %.3f e2e %.2f plugin %s"%(d["value"],d["ms_per_step"],r["frac"],d["e2e"]["value"],{k:round(v,3) for k,v in d["e2e_plugin"].items() if k.endswith("tflops")}))
PY
