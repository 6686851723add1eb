#!/bin/bash
# round-2 session Y (1 GPU): Gram-route SVD — solver tests, speed, DMRG parity and the Heisenberg ramp
OUT=gpurun_out; mkdir -p $OUT
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
timeout 600 python -m pytest tests/test_solver_gpu.py -m gpu -x -q > $OUT/y_pytest_solver.log 2>&1; echo "solver tests rc=$?"; tail -12 $OUT/y_pytest_solver.log | cut -c1-250
D=./build/plugin/dmrg_driver
cat /usr/local/cuda/lib64/libcusolver.so.11 /usr/local/cuda/lib64/libcublas.so.12 /usr/local/cuda/lib64/libcublasLt.so.12 > /dev/null
SCH="10,20,100,200,400,800,1200,2000"
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 $D heis_half 100 qn gpu $SCH 0 2 1e-7,1e-8,1e-10,0 $OUT/y_heis_1gpu.json > /dev/null 2> $OUT/y_heis_1gpu.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/y_heis_1gpu.json"))
    print("heis ramp to 2000: E=%.12f total %.1fs"%(d["energy"],d["total_seconds"]), [(s["maxlink"], round(s["seconds"],2)) for s in d["sweeps"]])
except Exception as e: print("no result", e)
PY
grep -E "svdOrd2|svd device|svd host" $OUT/y_heis_1gpu.err | head
ITB_SVD_GRAM_MIN_N=-1 ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 $D heis_half 100 qn gpu $SCH 0 2 1e-7,1e-8,1e-10,0 $OUT/y_heis_1gpu_polar.json > /dev/null 2> $OUT/y_heis_1gpu_polar.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/y_heis_1gpu_polar.json"))
    print("heis ramp to 2000 (polar only): E=%.12f total %.1fs"%(d["energy"],d["total_seconds"]), [(s["maxlink"], round(s["seconds"],2)) for s in d["sweeps"]])
except Exception as e: print("no result", e)
PY
grep -E "svdOrd2 wait|svd device" $OUT/y_heis_1gpu_polar.err | head -3
