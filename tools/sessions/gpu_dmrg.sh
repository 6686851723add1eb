#!/bin/bash
# DMRG through the reference's unmodified dmrg() on CPU storage vs HBM-resident storage, same binary.
OUT=gpurun_out; mkdir -p $OUT
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
D=./build/plugin/dmrg_driver
run() { # tag model N qn maxdims cutoffs niters noises
  for st in gpu cpu; do
    OPENBLAS_NUM_THREADS=${BLAS_THREADS:-1} timeout ${TMO:-900} $D $2 $3 $4 $st $5 $6 $7 $8 $OUT/dmrg_$1_$st.json > /dev/null 2> $OUT/dmrg_$1_$st.err || echo "$1 $st FAILED: $(tail -3 $OUT/dmrg_$1_$st.err)"
  done
  python - <<PY
import json
try:
    g=json.load(open("$OUT/dmrg_$1_gpu.json")); c=json.load(open("$OUT/dmrg_$1_cpu.json"))
    print("$1: E_gpu=%.12f E_cpu=%.12f dE=%.2e  t_gpu=%.2fs t_cpu=%.2fs  maxlink %d/%d launches %d" % (g["energy"],c["energy"],abs(g["energy"]-c["energy"]),g["total_seconds"],c["total_seconds"],g["sweeps"][-1]["maxlink"],c["sweeps"][-1]["maxlink"],g["gpu_launches"]))
    print("   sweep secs gpu", [round(s["seconds"],2) for s in g["sweeps"]], "cpu", [round(s["seconds"],2) for s in c["sweeps"]])
except Exception as e: print("$1: no result", e)
PY
}
run half20 heis_half 20 qn 10,20,100,100,200 1e-10 2 1e-7,1e-8,0
run one20d heis_one 12 dense 10,20,40 1e-10 2 1e-7,1e-8,0
run sample heis_one 100 qn 10,20,100,100,200 1e-10 2 1e-7,1e-8,0
if [ "$1" = "big" ]; then
run half400 heis_half 100 qn 10,20,100,400,400 0 2 1e-7,1e-8,0
fi
