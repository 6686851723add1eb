#!/usr/bin/env python
"""Per-CTA clock64 spans of the tile kernel on the bench workload (works on any build that has itb_contract_last_cta_cycles)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import itensor_b200 as itb
from itensor_b200 import synth
from itensor_b200._lib import lib, check
ctx = itb.Context(0)
structs = synth.heff_chain(synth.gaussian_sectors(2000, 9))
hosts = [synth.random_values(s, 10 + i) for i, s in enumerate(structs)]
dts = [itb.QTensor.from_host(ctx, s, h) for s, h in zip(structs, hosts)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.device)
lib().itb_ctx_set_profile(ctx.handle, 1)
cur = dts[0]
for k in range(4):
    p = itb.ContractPlan(cur.struct, structs[k + 1])
    out = itb.QTensor(ctx, p.C, ctx.empty(p.C.nreal))
    best = None
    for _ in range(8):
        flush.zero_()
        check(lib().itb_contract_run(ctx.handle, p._h, cur.ptr, dts[k + 1].ptr, out.ptr))
        ms = (C.c_float * 5)(); lib().itb_contract_last_ms(ctx.handle, ms)
        n = lib().itb_contract_last_cta_cycles(ctx.handle, None, 0)
        cyc = np.zeros(max(n, 1), np.int64); lib().itb_contract_last_cta_cycles(ctx.handle, cyc.ctypes.data_as(C.POINTER(C.c_int64)), n)
        if p.info.n_gemm_tiles and (best is None or ms[0] < best[0]): best = (ms[0], cyc[:n].copy())
    for phase in ("modelled", "refined"):
      if phase == "refined":
        if not p.info.n_gemm_tiles or not hasattr(p, "refine"): break
        lib().itb_ctx_set_profile(ctx.handle, 0)
        gain = p.refine(ctx, cur.ptr, dts[k + 1].ptr, out.ptr, rounds=4)
        lib().itb_ctx_set_profile(ctx.handle, 1)
        best = None
        for _ in range(8):
            flush.zero_()
            check(lib().itb_contract_run(ctx.handle, p._h, cur.ptr, dts[k + 1].ptr, out.ptr))
            ms = (C.c_float * 5)(); lib().itb_contract_last_ms(ctx.handle, ms)
            n = lib().itb_contract_last_cta_cycles(ctx.handle, None, 0)
            cyc = np.zeros(max(n, 1), np.int64); lib().itb_contract_last_cta_cycles(ctx.handle, cyc.ctypes.data_as(C.POINTER(C.c_int64)), n)
            if best is None or ms[0] < best[0]: best = (ms[0], cyc[:n].copy())
        print(f"   refine gain reported {gain:.3f}")
      if best is not None:
        c = best[1]
        print(f"[{phase}] ", end="")
        print(f"step {k+1}: tile class {best[0]*1e3:.1f} us, items {p.info.n_gemm_tiles}; CTA cycles max {c.max()} mean {c.mean():.0f} min {c.min()} max/mean {c.max()/c.mean():.3f} sum {c.sum()/1e6:.1f}M")
    cur = out
