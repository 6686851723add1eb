#!/bin/bash
# BASELINE configs[2] model (sample/hubbard_2d.cc, 16x4 cylinder, U=8, Nf+Sz conserved): GPU vs CPU sweep times
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02}
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
cat /usr/local/cuda/lib64/libcusolver.so.11 /usr/local/cuda/lib64/libcublas.so.12 /usr/local/cuda/lib64/libcublasLt.so.12 > /dev/null
D=./build/plugin/dmrg_driver
SG=${2:-20,60,100,200,400,800}
SC=${3:-20,60,100,200,400}
OPENBLAS_NUM_THREADS=8 timeout 1200 $D hubbard 16x4 qn cpu $SC 1e-6 2 1e-7,1e-8,1e-10,0 $OUT/${TAG}_hubbard_cpu.json > /dev/null 2> $OUT/${TAG}_hubbard_cpu.err &
CPID=$!
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 1200 $D hubbard 16x4 qn gpu $SG 1e-6 2 1e-7,1e-8,1e-10,0 $OUT/${TAG}_hubbard_gpu.json > /dev/null 2> $OUT/${TAG}_hubbard_gpu.err
wait $CPID
python - <<PY
import json
for t in ("gpu","cpu"):
    try:
        d=json.load(open("$OUT/${TAG}_hubbard_%s.json"%t))
        print(t, "E=%.10f total %.1fs"%(d["energy"],d["total_seconds"]), [(s["maxlink"], round(s["seconds"],1), round(s["energy"],8)) for s in d["sweeps"]])
    except Exception as e: print(t, "no result", e)
PY
tail -24 $OUT/${TAG}_hubbard_gpu.err
