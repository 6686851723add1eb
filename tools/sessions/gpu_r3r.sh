#!/bin/bash
# round-2 session R (1 GPU): graph-replay bench, TRG spectra test, synthetic sweep to m=8192, ncu captures for profiles/
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r_bench_graph.json 2> $OUT/r_bench_graph.err; echo "bench graph rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph > $OUT/r_bench_eager.json 2> $OUT/r_bench_eager.err; echo "bench eager rc=$?"
python - <<'PY'
import json
for f in ("r_bench_graph","r_bench_eager"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().split("\n")[-1]); r=d["roofline"]
    print(f,"value %.2f ms %.3f frac %.3f launch: %s launches %d"%(d["value"],d["ms_per_step"],r["frac"],d["config"]["launch"],d["gpu_launches"]))
PY
timeout 1500 python -m pytest tests/test_plugin_dmrg.py -m gpu -x -q -s -k "trg_per_scale" > $OUT/r_pytest_trg.log 2>&1; echo "trg test rc=$?"; grep -E "TRG maxdim|passed|failed|assert" $OUT/r_pytest_trg.log | head -5
timeout 900 python tools/synth_sweep.py --ms 2048,4096,8192 --sectors 4,8,16 --out $OUT/r_synth_sweep.jsonl > $OUT/r_synth.log 2>&1; echo "synth rc=$?"; tail -3 $OUT/r_synth.log | cut -c1-200
# ncu: launch list + full captures of the three dominant kernels of the default build
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 40 --csv --log-file $OUT/r_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsc_gemm_static -s 4 -c 2 -o $OUT/r_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:perm_tile_pipe -s 2 -c 2 -o $OUT/r_perm python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
ls -la $OUT/r_*.ncu-rep
