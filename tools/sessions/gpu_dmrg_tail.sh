#!/bin/bash
# GPU-only: ramp quickly to large maxdim and time the big sweeps
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02t}
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
cat /usr/local/cuda/lib64/libcusolver.so.11 /usr/local/cuda/lib64/libcublas.so.12 /usr/local/cuda/lib64/libcublasLt.so.12 > /dev/null
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 1500 ./build/plugin/dmrg_driver heis_half 100 qn gpu ${2:-10,20,100,200,400,800,1200,1600,2000,2000} 0 2 1e-7,1e-8,1e-10,0 $OUT/${TAG}_gpu.json > /dev/null 2> $OUT/${TAG}_gpu.err
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_gpu.json"))
print("E=%.12f total %.1fs"%(d["energy"],d["total_seconds"]), [(s["maxlink"], round(s["seconds"],1), round(s["energy"],10)) for s in d["sweeps"]])
PY
tail -22 $OUT/${TAG}_gpu.err
