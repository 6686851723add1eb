#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list + one full capture of the top kernel.
# Usage (from repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/${TAG}_pytest.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.txt
echo "== bench" ; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee $OUT/${TAG}_bench.json
if [ "$2" != "noncu" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
echo "== ncu full (gemm kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bsc_gemm_kernel -s 8 -c 2 -o $OUT/${TAG}_gemm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
fi
echo "== done"
