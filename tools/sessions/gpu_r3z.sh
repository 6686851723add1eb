#!/bin/bash
# round-2 session Z (1 GPU): final state — full GPU suite, smoke, both bench arms, Hubbard 16x4 ramp to maxdim 2000
OUT=gpurun_out; mkdir -p $OUT
export LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/z_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 $OUT/z_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/z_smoke.log
timeout 600 python bench.py --impl reference > $OUT/z_bench_reference.json 2> $OUT/z_bench_reference.err; echo "reference arm rc=$?"
timeout 600 python bench.py > $OUT/z_bench.json 2> $OUT/z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
r=json.loads(open("gpurun_out/z_bench_reference.json").read().strip().split("\n")[-1])
d=json.loads(open("gpurun_out/z_bench.json").read().strip().split("\n")[-1]); ro=d["roofline"]
print("reference arm: %.4f TFLOP/s (%s)"%(r["value"], r["cpu_baseline"].get("sample","")[:120]))
print("bench: value %.2f ms %.3f e2e %.2f frac %.3f traffic %.3g launches %s cpu %s parity %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],ro["frac"],ro["traffic"] or 0,d["gpu_launches"],{k:(round(v["tflops"],4) if isinstance(v,dict) and "tflops" in v else v) for k,v in (d["cpu_baseline"].get("modes") or {}).items()} if d.get("cpu_baseline") else None, d["parity_vs_reference"]["ok"] if d.get("parity_vs_reference") else None))
print("permute", round(d["permute"]["frac"],3), "accumulate", round(d["permute"]["accumulate"]["frac"],3), "plugin", {k:round(v,3) for k,v in d["e2e_plugin"].items() if k.endswith("tflops")})
PY
D=./build/plugin/dmrg_driver
cat /usr/local/cuda/lib64/libcusolver.so.11 /usr/local/cuda/lib64/libcublas.so.12 /usr/local/cuda/lib64/libcublasLt.so.12 > /dev/null
SH="20,60,100,200,400,800,1200,2000"
ITB_PROFILE=1 OPENBLAS_NUM_THREADS=4 timeout 900 $D hubbard 16x4 qn gpu $SH 1e-6 2 1e-7,1e-8,1e-10,1e-10,0 $OUT/z_hub2000_1gpu.json > /dev/null 2> $OUT/z_hub2000_1gpu.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/z_hub2000_1gpu.json"))
    print("hubbard 16x4 ramp to 2000: E=%.12f total %.1fs"%(d["energy"],d["total_seconds"]), [(s["maxlink"], round(s["seconds"],2)) for s in d["sweeps"]])
except Exception as e: print("no result", e)
PY
grep -E "Contract QDenseGPU|diagH launch|diagH host|diagH wait|PlusEQ|combine " $OUT/z_hub2000_1gpu.err | head
