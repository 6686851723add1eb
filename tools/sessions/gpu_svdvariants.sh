#!/bin/bash
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
for v in "ITB_SOLVER_MIN_N=100000" "ITB_SOLVER_MIN_N=128" "ITB_SOLVER_MIN_N=256"; do
  echo "== eigh path $v"
  env $v OPENBLAS_NUM_THREADS=4 timeout 600 ./build/plugin/dmrg_driver_times heis_half 100 qn gpu 10,20,100,400,400 1e-14 2 1e-7,1e-8,1e-12 2>&1 | grep -E "Sweep 5/5|Section  [134]," | tail -4
done
