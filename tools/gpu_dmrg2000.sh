#!/bin/bash
# BASELINE configs[1]: S=1/2 Heisenberg N=100, Sz QDense blocks, ramp to maxdim 2000 on HBM storage;
# CPU reference (same binary, host storage) on the same box for the sweeps it can finish in minutes.
OUT=gpurun_out; mkdir -p $OUT
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
D=./build/plugin/dmrg_driver
SCHED_G="10,20,100,200,400,800,1200,1600,2000,2000"
SCHED_C="10,20,100,200,400,800"
OPENBLAS_NUM_THREADS=8 timeout 1500 $D heis_half 100 qn cpu $SCHED_C 0 2 1e-7,1e-8,1e-10 $OUT/dmrg2000_cpu.json > /dev/null 2> $OUT/dmrg2000_cpu.err &
CPU_PID=$!
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 1500 $D heis_half 100 qn gpu $SCHED_G 0 2 1e-7,1e-8,1e-10 $OUT/dmrg2000_gpu.json > /dev/null 2> $OUT/dmrg2000_gpu.err
wait $CPU_PID
python - <<PY
import json
for t in ("gpu","cpu"):
    try:
        d=json.load(open("$OUT/dmrg2000_%s.json"%t))
        print(t, "E=%.10f total %.1fs"%(d["energy"],d["total_seconds"]), "sweeps (maxlink, s):", [(s["maxlink"], round(s["seconds"],1)) for s in d["sweeps"]])
    except Exception as e: print(t, "no result", e)
PY
tail -12 $OUT/dmrg2000_gpu.err
