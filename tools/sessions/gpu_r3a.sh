#!/bin/bash
# round-2 session A: validate the dynamic tile queue + pipelined permute, then measure variants
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/a_gpu.txt; nproc >> $OUT/a_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/a_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/a_smoke.log
timeout 900 python -m pytest tests/test_contract_gpu.py tests/test_permute_blas1_gpu.py tests/test_golden.py -m gpu -x -q > $OUT/a_pytest1.log 2>&1; echo "pytest1 rc=$?" | tee -a $OUT/a_pytest1.log
tail -5 $OUT/a_pytest1.log
for f in 1 2 3; do
  ITB_GUIDED_FACTOR=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/a_bench_f$f.json 2> $OUT/a_bench_f$f.err; echo "bench f=$f rc=$?"
done
for v in "0 3" "1 2" "1 3" "1 4" "1 6"; do set -- $v
  ITB_PERM_PIPE=$1 ITB_PERM_CTAS=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/a_bench_p$1_$2.json 2> $OUT/a_bench_p$1_$2.err; echo "bench perm $v rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/a_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        r=d["roofline"]; p=d["permute"]
        print(f, "value %.2f ms %.3f frac %.3f tile_ms %.3f stream_ms %.3f e2e %.2f perm %.0f GB/s (%.3f)"%(d["value"],d["ms_per_step"],r["frac"],r["ms_per_step"]["tile_kernel"],r["ms_per_step"]["streaming_kernel"],d["e2e"]["value"],p["achieved_gbs"],p["frac"]))
    except Exception as e: print(f,"ERR",e)
PY
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/a_pytest_all.log 2>&1; echo "pytest all rc=$?" | tee -a $OUT/a_pytest_all.log
tail -15 $OUT/a_pytest_all.log
