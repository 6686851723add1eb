#!/bin/bash
# BASELINE configs[1]: S=1/2 Heisenberg N=100, Sz QDense blocks, ramp to maxdim 2000 on HBM storage (noise -> 0 and
# cutoff 0 from sweep 4 on: the SVD path of svdBond); CPU reference (same binary, host storage) on the same box for
# the sweeps it can finish in minutes.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02}
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
D=./build/plugin/dmrg_driver
SCHED_G="10,20,100,200,400,800,1200,1600,2000,2000"
SCHED_C="10,20,100,200,400,800"
# page the solver libraries in before anything is timed (a fresh box reads them at ~30 MB/s on first touch)
cat /usr/local/cuda/lib64/libcusolver.so.11 /usr/local/cuda/lib64/libcublas.so.12 /usr/local/cuda/lib64/libcublasLt.so.12 > /dev/null
nproc > $OUT/${TAG}_nproc.txt
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 1500 $D heis_half 100 qn gpu $SCHED_G 0 2 1e-7,1e-8,1e-10,0 $OUT/${TAG}_dmrg2000_gpu.json > /dev/null 2> $OUT/${TAG}_dmrg2000_gpu.err
OPENBLAS_NUM_THREADS=$(nproc) timeout 900 $D heis_half 100 qn cpu $SCHED_C 0 2 1e-7,1e-8,1e-10,0 $OUT/${TAG}_dmrg2000_cpu.json > /dev/null 2> $OUT/${TAG}_dmrg2000_cpu.err
python - <<PY
import json
for t in ("gpu","cpu"):
    try:
        d=json.load(open("$OUT/${TAG}_dmrg2000_%s.json"%t))
        print(t, "E=%.12f total %.1fs"%(d["energy"],d["total_seconds"]), "sweeps (maxlink, s, E):", [(s["maxlink"], round(s["seconds"],1), round(s["energy"],10)) for s in d["sweeps"]])
    except Exception as e: print(t, "no result", e)
PY
tail -22 $OUT/${TAG}_dmrg2000_gpu.err
