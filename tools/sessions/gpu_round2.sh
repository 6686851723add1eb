#!/bin/bash
# end-of-round evidence: tests, bench line, ncu launch list, full captures of the three dominant kernels
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02z}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/${TAG}_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $OUT/${TAG}_smoke.txt
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 2500 $OUT/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_reference.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsc_gemm -s 6 -c 2 -o $OUT/${TAG}_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsc_rowgroup -s 8 -c 1 -o $OUT/${TAG}_rowgroup python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_rowgroup.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:perm_tile -s 2 -c 1 -o $OUT/${TAG}_perm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_perm.log 2>&1
ls -la $OUT/${TAG}_*
