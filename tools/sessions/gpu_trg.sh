#!/bin/bash
# BASELINE configs[4]: sample/trg.cc at maxdim 32 / 64 (/ 96), GPU vs CPU (all host cores)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02}
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
cat /usr/local/cuda/lib64/libcusolver.so.11 /usr/local/cuda/lib64/libcublas.so.12 /usr/local/cuda/lib64/libcublasLt.so.12 > /dev/null
T=./build/plugin/trg_driver
for chi in ${2:-20 32 48 64}; do
  ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 $T $chi 12 gpu $OUT/${TAG}_trg${chi}_gpu.json 2> $OUT/${TAG}_trg${chi}_gpu.err | tail -1
  if [ $chi -le ${3:-48} ]; then OPENBLAS_NUM_THREADS=$(nproc) timeout 900 $T $chi 12 cpu $OUT/${TAG}_trg${chi}_cpu.json | tail -1; fi
done
tail -20 $OUT/${TAG}_trg64_gpu.err
