#!/bin/bash
# round-2 session J (N GPUs, default 2): bench under torchrun, row-sharded DMRG through the plugin
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$N" = "2" ]; then
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/j_bench_n1.json 2> $OUT/j_bench_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/j_bench_n1.json").read().strip().split("\n")[-1]); r=d["roofline"]; p=d["permute"]
print("N=1 default: value %.2f ms %.3f frac %.3f tile_ms %.3f e2e %.2f perm %.3f acc %.3f"%(d["value"],d["ms_per_step"],r["frac"],r["ms_per_step"]["tile_kernel"],d["e2e"]["value"],p["frac"],p["accumulate"]["frac"]))
PY
ITB_SCHED=guided timeout 600 python -m pytest tests/test_contract_gpu.py -m gpu -x -q 2>&1 | tail -1
timeout 600 python -m pytest tests/test_contract_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -1
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/j_bench_n$N.json 2> $OUT/j_bench_n$N.err; echo "bench N=$N rc=$?"
tail -3 $OUT/j_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/j_bench_n$N.json").read().strip().split("\n")[-1])
    print("N=%d value %.2f TFLOP/s ms %.3f e2e %.2f sharding: %s"%(d["n_gpus"],d["value"],d["ms_per_step"],d["e2e"]["value"],d["config"]["sharding"]))
except Exception as e: print("bench parse failed", e)
PY
if [ "$N" = "2" ]; then
timeout 900 python -m pytest tests/test_plugin_dmrg.py -m gpu -x -q -k "two_gpus" > $OUT/j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/j_pytest.log
fi
D=./build/plugin/dmrg_driver
cat /usr/local/cuda/lib64/libcusolver.so.11 /usr/local/cuda/lib64/libcublas.so.12 /usr/local/cuda/lib64/libcublasLt.so.12 > /dev/null
# S=1/2 Heisenberg N=100 ramp (SVD path from sweep 4): 1 GPU vs N GPUs
SCH="10,20,100,200,400,800,1200"
if [ ! -f $OUT/j_heis_1gpu.json ]; then
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 $D heis_half 100 qn gpu $SCH 0 2 1e-7,1e-8,1e-10,0 $OUT/j_heis_1gpu.json > /dev/null 2> $OUT/j_heis_1gpu.err
fi
RANK_LOG_DIR=$OUT ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 tools/run_ranks.sh $N $D heis_half 100 qn gpu $SCH 0 2 1e-7,1e-8,1e-10,0 $OUT/j_heis_${N}gpu.json > /dev/null 2> $OUT/j_heis_${N}gpu.err
# Hubbard 16x4 ramp to 800 (configs[2] model): 1 GPU vs N GPUs
SH="20,60,100,200,400,800"
if [ ! -f $OUT/j_hub_1gpu.json ]; then
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 $D hubbard 16x4 qn gpu $SH 1e-6 2 1e-7,1e-8,1e-10,0 $OUT/j_hub_1gpu.json > /dev/null 2> $OUT/j_hub_1gpu.err
fi
RANK_LOG_DIR=$OUT ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 tools/run_ranks.sh $N $D hubbard 16x4 qn gpu $SH 1e-6 2 1e-7,1e-8,1e-10,0 $OUT/j_hub_${N}gpu.json > /dev/null 2> $OUT/j_hub_${N}gpu.err
python - <<PY
import json
for t in ("j_heis_1gpu","j_heis_${N}gpu","j_hub_1gpu","j_hub_${N}gpu"):
    try:
        d=json.load(open("gpurun_out/%s.json"%t))
        print(t, "E=%.12f total %.1fs"%(d["energy"],d["total_seconds"]), [(s["maxlink"], round(s["seconds"],1)) for s in d["sweeps"]])
    except Exception as e: print(t, "no result", e)
PY
grep -E "all-gather|Contract QDenseGPU|svdOrd2 wait|diagH wait" $OUT/j_heis_${N}gpu.err $OUT/j_hub_${N}gpu.err $OUT/j_hub_1gpu.err | head -20
