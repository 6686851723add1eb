#!/bin/bash
# round-2 session AG (2 GPUs): final check of the multi-GPU paths (tests + bench)
OUT=gpurun_out; mkdir -p $OUT
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
timeout 900 python -m pytest tests/test_plugin_dmrg.py tests/test_p2p_gpu.py -m gpu -x -q -k "two_gpus or p2p" 2>&1 | tail -1
bash tools/gpu_bench_n.sh 2
