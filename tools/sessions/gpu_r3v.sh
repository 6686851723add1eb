#!/bin/bash
# round-2 session V (1 GPU): full GPU test suite on the current build + DMRG ramps with the stream-ordered allocator / helper-thread eigh
OUT=gpurun_out; mkdir -p $OUT
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/v_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 $OUT/v_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/v_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/v_smoke.log
D=./build/plugin/dmrg_driver
cat /usr/local/cuda/lib64/libcusolver.so.11 /usr/local/cuda/lib64/libcublas.so.12 /usr/local/cuda/lib64/libcublasLt.so.12 > /dev/null
SH="20,60,100,200,400,800"
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 $D hubbard 16x4 qn gpu $SH 1e-6 2 1e-7,1e-8,1e-10,0 $OUT/v_hub_1gpu.json > /dev/null 2> $OUT/v_hub_1gpu.err
SCH="10,20,100,200,400,800,1200"
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 $D heis_half 100 qn gpu $SCH 0 2 1e-7,1e-8,1e-10,0 $OUT/v_heis_1gpu.json > /dev/null 2> $OUT/v_heis_1gpu.err
python - <<PY
import json
for t in ("v_hub_1gpu","v_heis_1gpu"):
    try:
        d=json.load(open("gpurun_out/%s.json"%t))
        print(t, "E=%.12f total %.1fs"%(d["energy"],d["total_seconds"]), [(s["maxlink"], round(s["seconds"],2)) for s in d["sweeps"]])
    except Exception as e: print(t, "no result", e)
PY
grep -E "Contract|PlusEQ|combine|svdOrd2 wait|diagH" $OUT/v_hub_1gpu.err | head -24
grep -E "Contract QDenseGPU|Contract: d|svdOrd2 wait|svdOrd2 host" $OUT/v_heis_1gpu.err | head
