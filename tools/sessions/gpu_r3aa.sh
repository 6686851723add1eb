#!/bin/bash
# round-2 session AA (1 GPU): per-lane threads of the eigh batch — solver tests, DMRG tests on the eigh path, Hubbard ramps
OUT=gpurun_out; mkdir -p $OUT
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_contract_gpu.py -m gpu -x -q > $OUT/aa_pytest_a.log 2>&1; echo "solver+contract tests rc=$?"; tail -2 $OUT/aa_pytest_a.log
timeout 1500 python -m pytest tests/test_plugin_dmrg.py -m gpu -x -q -k "not trg_per_scale and not two_gpus and not config2" > $OUT/aa_pytest_b.log 2>&1; echo "plugin tests rc=$?"; tail -2 $OUT/aa_pytest_b.log
D=./build/plugin/dmrg_driver
cat /usr/local/cuda/lib64/libcusolver.so.11 /usr/local/cuda/lib64/libcublas.so.12 /usr/local/cuda/lib64/libcublasLt.so.12 > /dev/null
SH="20,60,100,200,400,800,1200,2000"
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 $D hubbard 16x4 qn gpu $SH 1e-6 2 1e-7,1e-8,1e-10,1e-10,0 $OUT/aa_hub2000_1gpu.json > /dev/null 2> $OUT/aa_hub2000_1gpu.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/aa_hub2000_1gpu.json"))
    print("hubbard 16x4 ramp to 2000: E=%.12f total %.1fs"%(d["energy"],d["total_seconds"]), [(s["maxlink"], round(s["seconds"],2)) for s in d["sweeps"]])
except Exception as e: print("no result", e)
PY
grep -E "Contract QDenseGPU|diagH launch|diagH host|diagH wait" $OUT/aa_hub2000_1gpu.err | head
