#!/usr/bin/env python
"""Where does the GPU-vs-CPU difference of one config-2 sweep at maxdim 800 come from? Same saved state, one CPU sweep, GPU
sweeps with the decompositions routed differently (device polar+Jacobi SVD / Jacobi only / host LAPACK)."""
import json, os, subprocess, sys, sysconfig, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "build", "plugin", "dmrg_driver")
ENV = dict(os.environ, LD_LIBRARY_PATH=os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs") + ":" + os.environ.get("LD_LIBRARY_PATH", ""),
           ITB_WARM_LIBS="0")
tmp = tempfile.mkdtemp()
m = sys.argv[1] if len(sys.argv) > 1 else "800"
ramp = {"400": "10,20,100,200,400", "800": "10,20,100,200,400,800"}[m]
def run(args, env):
    out = subprocess.run([D] + args, env=dict(ENV, **env), capture_output=True, text=True, timeout=3000)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().split("\n")[-1])
st = os.path.join(tmp, "st")
run(["heis_half", "100", "qn", "gpu", ramp, "0", "2", "1e-7,1e-8,1e-10,0", "--save", st], {"OPENBLAS_NUM_THREADS": "4"})
ncpu = str(len(os.sched_getaffinity(0)))
c = run(["heis_half", "100", "qn", "cpu", m, "0", "2", "0", "--load", st, "--bonds", os.path.join(tmp, "c.json")], {"OPENBLAS_NUM_THREADS": ncpu})
cb = json.load(open(os.path.join(tmp, "c.json")))
variants = {
    "device svd: polar >= 96, jacobi below, host < 24": {"ITB_SVD_DEVICE_MIN_N": "24", "ITB_SVD_MIN_N": "24", "ITB_EIGH_MIN_N": "24"},
    "device svd: default thresholds (host < 160)": {},
    "device svd: jacobi only": {"ITB_SVD_DEVICE_MIN_N": "24", "ITB_SVD_POLAR_MIN_N": "1000000"},
    "host LAPACK svd (ITB_SVD_DEVICE=0, LAPACK boundary on host)": {"ITB_SVD_DEVICE": "0", "ITB_SVD_MIN_N": "1000000000", "ITB_EIGH_MIN_N": "1000000000"},
}
res = {}
for name, env in variants.items():
    g = run(["heis_half", "100", "qn", "gpu", m, "0", "2", "0", "--load", st, "--bonds", os.path.join(tmp, "g.json")], dict(env, OPENBLAS_NUM_THREADS="4"))
    gb = json.load(open(os.path.join(tmp, "g.json")))
    de = [abs(a["energy"] - b["energy"]) for a, b in zip(gb, cb)]
    ds = [float(np.abs(np.array(a["spectrum"]) - np.array(b["spectrum"])).max()) for a, b in zip(gb, cb)]
    first = next((i for i, v in enumerate(ds) if v > 1e-12), -1)
    res[name] = {"max_dE": max(de), "max_dspec": max(ds), "first_bond_with_dspec_gt_1e-12": first, "dspec_first_10_bonds": ds[:10], "seconds": g["sweeps"][-1]["seconds"]}
    print(f"{name}: max |dE| {max(de):.2e}  max |dspec| {max(ds):.2e}  first bond with |dspec|>1e-12: {first}  first bonds {['%.1e' % v for v in ds[:8]]}  {g['sweeps'][-1]['seconds']:.1f} s", flush=True)
json.dump({"maxdim": int(m), "cpu_seconds": c["sweeps"][-1]["seconds"], "cpu_threads": ncpu, "variants": res}, open(os.path.join(ROOT, "gpurun_out", f"config2_diag_m{m}.json"), "w"), indent=1)
