#!/bin/bash
# round-2 session T (1 GPU): planner speed-ups + staging ring in the Hubbard / Heisenberg ramps, e2e upload order, parity subset
OUT=gpurun_out; mkdir -p $OUT
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/t_bench.json 2> $OUT/t_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/t_bench.json").read().strip().split("\n")[-1]); r=d["roofline"]
print("value %.2f ms %.3f tile %.2f TF/s frac %.3f e2e %.2f (%.3f ms) plugin %s"%(d["value"],d["ms_per_step"],r["achieved"],r["frac"],d["e2e"]["value"],d["e2e"]["ms_per_step"],{k:round(v,3) for k,v in d["e2e_plugin"].items() if k.endswith("tflops") or k.endswith("ms_per_step")}))
PY
timeout 900 python -m pytest tests/test_contract_gpu.py tests/test_golden.py tests/test_permute_blas1_gpu.py -m gpu -x -q 2>&1 | tail -2
D=./build/plugin/dmrg_driver
cat /usr/local/cuda/lib64/libcusolver.so.11 /usr/local/cuda/lib64/libcublas.so.12 /usr/local/cuda/lib64/libcublasLt.so.12 > /dev/null
SH="20,60,100,200,400,800"
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 $D hubbard 16x4 qn gpu $SH 1e-6 2 1e-7,1e-8,1e-10,0 $OUT/t_hub_1gpu.json > /dev/null 2> $OUT/t_hub_1gpu.err
SCH="10,20,100,200,400,800,1200"
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 $D heis_half 100 qn gpu $SCH 0 2 1e-7,1e-8,1e-10,0 $OUT/t_heis_1gpu.json > /dev/null 2> $OUT/t_heis_1gpu.err
python - <<PY
import json
for t in ("t_hub_1gpu","t_heis_1gpu"):
    try:
        d=json.load(open("gpurun_out/%s.json"%t))
        print(t, "E=%.12f total %.1fs"%(d["energy"],d["total_seconds"]), [(s["maxlink"], round(s["seconds"],2)) for s in d["sweeps"]])
    except Exception as e: print(t, "no result", e)
PY
grep -E "Contract|PlusEQ|combine|svdOrd2 wait|diagH" $OUT/t_hub_1gpu.err | head -24
timeout 900 python -m pytest tests/test_plugin_dmrg.py -m gpu -x -q -k "not trg_per_scale and not two_gpus and not config2" 2>&1 | tail -2
