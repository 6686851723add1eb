#!/bin/bash
# round-2 session S (1 GPU): measured re-partition of the static tile schedule (itb_contract_plan_refine) A/B, its parity test, TRG spectra test
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/s_bench_refine.json 2> $OUT/s_bench_refine.err; echo "bench refine rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-refine > $OUT/s_bench_norefine.json 2> $OUT/s_bench_norefine.err; echo "bench norefine rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --refine-rounds 8 > $OUT/s_bench_refine8.json 2> $OUT/s_bench_refine8.err; echo "bench refine8 rc=$?"
python - <<'PY'
import json
for f in ("s_bench_refine","s_bench_norefine","s_bench_refine8"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().split("\n")[-1]); r=d["roofline"]
        print(f,"value %.2f ms %.3f tile %.2f TF/s frac %.3f partition: %s"%(d["value"],d["ms_per_step"],r["achieved"],r["frac"],d["config"]["tile_partition"][-60:]))
    except Exception as e: print(f, "failed", e)
PY
tail -3 $OUT/s_bench_refine.err
timeout 900 python -m pytest tests/test_contract_gpu.py -m gpu -x -q -k "refined or bench_workload" > $OUT/s_pytest_refine.log 2>&1; echo "refine test rc=$?"; tail -3 $OUT/s_pytest_refine.log
timeout 1500 python -m pytest tests/test_plugin_dmrg.py -m gpu -x -q -s -k "trg_per_scale" > $OUT/s_pytest_trg.log 2>&1; echo "trg test rc=$?"; grep -E "TRG maxdim|passed|failed|assert" $OUT/s_pytest_trg.log | head -5
# per-CTA balance before / after
timeout 300 python tools/cta_stats.py > $OUT/s_cta_stats.txt 2>&1; tail -12 $OUT/s_cta_stats.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsc_gemm_static -s 40 -c 2 -o $OUT/s_gemm python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
ls -la $OUT/s_*.ncu-rep
