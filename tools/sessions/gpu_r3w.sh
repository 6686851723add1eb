#!/bin/bash
# round-2 session W (2 GPUs): 2-GPU tests (plugin row-sharded DMRG, peer-memory exchange) + Hubbard ramp on 2 GPUs through the plugin
OUT=gpurun_out; mkdir -p $OUT
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
timeout 900 python -m pytest tests/test_plugin_dmrg.py tests/test_p2p_gpu.py -m gpu -x -q -k "two_gpus or p2p" > $OUT/w_pytest_2gpu.log 2>&1; echo "2-GPU tests rc=$?"; tail -3 $OUT/w_pytest_2gpu.log
D=./build/plugin/dmrg_driver
cat /usr/local/cuda/lib64/libcusolver.so.11 /usr/local/cuda/lib64/libcublas.so.12 /usr/local/cuda/lib64/libcublasLt.so.12 > /dev/null
SH="20,60,100,200,400,800"
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 600 $D hubbard 16x4 qn gpu $SH 1e-6 2 1e-7,1e-8,1e-10,0 $OUT/w_hub_1gpu.json > /dev/null 2> $OUT/w_hub_1gpu.err
RANK_LOG_DIR=$OUT ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 600 tools/run_ranks.sh 2 $D hubbard 16x4 qn gpu $SH 1e-6 2 1e-7,1e-8,1e-10,0 $OUT/w_hub_2gpu.json > /dev/null 2> $OUT/w_hub_2gpu.err
python - <<PY
import json
for t in ("w_hub_1gpu","w_hub_2gpu"):
    try:
        d=json.load(open("gpurun_out/%s.json"%t))
        print(t, "E=%.12f total %.1fs"%(d["energy"],d["total_seconds"]), [(s["maxlink"], round(s["seconds"],2)) for s in d["sweeps"]])
    except Exception as e: print(t, "no result", e)
PY
grep -E "diagH host|diagH launch|Contract QDenseGPU|all-gather" $OUT/w_hub_1gpu.err $OUT/w_hub_2gpu.err | head
