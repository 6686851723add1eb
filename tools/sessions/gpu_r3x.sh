#!/bin/bash
# round-2 session X (2 GPUs): 2-GPU tests again, Heisenberg ramp to 2000 on 1 and 2 GPUs through the plugin (the contractions that are worth sharding)
OUT=gpurun_out; mkdir -p $OUT
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
timeout 900 python -m pytest tests/test_plugin_dmrg.py tests/test_p2p_gpu.py -m gpu -x -q -k "two_gpus or p2p" > $OUT/x_pytest_2gpu.log 2>&1; echo "2-GPU tests rc=$?"; tail -3 $OUT/x_pytest_2gpu.log
D=./build/plugin/dmrg_driver
cat /usr/local/cuda/lib64/libcusolver.so.11 /usr/local/cuda/lib64/libcublas.so.12 /usr/local/cuda/lib64/libcublasLt.so.12 > /dev/null
SCH="10,20,100,200,400,800,1200,2000"
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 $D heis_half 100 qn gpu $SCH 0 2 1e-7,1e-8,1e-10,0 $OUT/x_heis_1gpu.json > /dev/null 2> $OUT/x_heis_1gpu.err &
wait
RANK_LOG_DIR=$OUT ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 tools/run_ranks.sh 2 $D heis_half 100 qn gpu $SCH 0 2 1e-7,1e-8,1e-10,0 $OUT/x_heis_2gpu.json > /dev/null 2> $OUT/x_heis_2gpu.err
python - <<PY
import json
for t in ("x_heis_1gpu","x_heis_2gpu"):
    try:
        d=json.load(open("gpurun_out/%s.json"%t))
        print(t, "E=%.12f total %.1fs"%(d["energy"],d["total_seconds"]), [(s["maxlink"], round(s["seconds"],2)) for s in d["sweeps"]])
    except Exception as e: print(t, "no result", e)
PY
grep -E "svdOrd2 wait|svdOrd2 host|Contract QDenseGPU|all-gather" $OUT/x_heis_1gpu.err $OUT/x_heis_2gpu.err | head
