#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02i}
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
SCHED=${2:-10,20,100,200,400,800,1200,1200}
ITB_PROFILE=${4:-1} OPENBLAS_NUM_THREADS=${3:-4} timeout 1200 ./build/plugin/dmrg_driver_times heis_half 100 qn gpu $SCHED 0 2 1e-7,1e-8,1e-10,0 $OUT/${TAG}_times.json > $OUT/${TAG}_times.out 2> $OUT/${TAG}_times.err
cut -c1-160 $OUT/${TAG}_times.out | grep -v "^{" | grep "Sweep\|Section" | tail -14
tail -12 $OUT/${TAG}_times.err
python -c "
import json; d=json.load(open('$OUT/${TAG}_times.json')); print([(s['maxlink'], round(s['seconds'],1)) for s in d['sweeps']], d['energy'])"
