#!/usr/bin/env python
"""BASELINE configs[3]: synthetic block-sparse contraction sweep — Sz sectors S in {4,8,16}, equal-size and
binomial-weighted (--dist binomial), bond dimension m in {512..8192}, real and complex: TFLOP/s of phi*L (the
tensor-pipe step) and of the whole H_eff*phi chain, device-resident, CUDA events, best of 3. One JSON line per case."""
import argparse, ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itensor_b200 as itb
from itensor_b200 import ITB_C64, ITB_F64, synth
from itensor_b200._lib import check, lib

ap = argparse.ArgumentParser()
ap.add_argument("--ms", default="512,1024,2048,4096,8192")
ap.add_argument("--sectors", default="4,8,16")
ap.add_argument("--out", default="gpurun_out/synth_sweep.jsonl")
ap.add_argument("--dist", default="equal,binomial")
ap.add_argument("--dtypes", default="real,complex")
args = ap.parse_args()
ctx = itb.Context(0)
dev = ctx.device
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = open(args.out, "w")
def binomial_sectors(m, S):
    """sector sizes proportional to binomial(S-1, q), at least 1, summing to ~m"""
    from math import comb
    w = np.array([comb(S - 1, q) for q in range(S)], float)
    return [int(v) for v in np.maximum(1, np.rint(m * w / w.sum()))]

for dtype, dname in [(ITB_F64, "real"), (ITB_C64, "complex")]:
  if dname not in args.dtypes.split(","):
    continue
  for dist in args.dist.split(","):
    for S in [int(x) for x in args.sectors.split(",")]:
        for m in [int(x) for x in args.ms.split(",")]:
            sizes = synth.equal_sectors(m, S) if dist == "equal" else binomial_sectors(m, S)
            structs = synth.heff_chain(sizes, dtype=dtype)
            plans, s = [], structs[0]
            for t in structs[1:]:
                p = itb.ContractPlan(s, t); plans.append(p); s = p.C
            need = sum(st.nreal for st in structs) + sum(p.C.nreal for p in plans)
            if need * 8 > 120e9:
                continue
            gen = torch.Generator(device=dev); gen.manual_seed(m + S)
            dts = [itb.QTensor(ctx, st, torch.rand(max(st.nreal, 1), dtype=torch.float64, device=dev, generator=gen)[:st.nreal] * 2 - 1) for st in structs]
            outs = [itb.QTensor(ctx, p.C, ctx.empty(p.C.nreal)) for p in plans]
            def run(k0, k1):
                cur = dts[0] if k0 == 0 else outs[k0 - 1]
                for k in range(k0, k1):
                    check(lib().itb_contract_run(ctx.handle, plans[k]._h, cur.ptr, dts[k + 1].ptr, outs[k].ptr))
                    cur = outs[k]
            def timed(k0, k1):
                run(k0, k1); best = 1e9
                for _ in range(3):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); run(k0, k1); e1.record(); e1.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                return best
            t_all = timed(0, 4)
            t_1 = timed(0, 1)
            fl_all = sum(p.flops for p in plans)
            rec = {"dtype": dname, "dist": dist, "sectors": S, "m": m, "sector_sizes": sizes, "pairs": [int(p.npairs) for p in plans], "flops_heff": fl_all,
                   "heff_ms": t_all, "heff_tflops": fl_all / t_all / 1e9, "phiL_ms": t_1, "phiL_tflops": plans[0].flops / t_1 / 1e9}
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n"); out.flush()
            del dts, outs, plans
            torch.cuda.empty_cache()
