#!/usr/bin/env python
"""Fit the planner's cycle model to measured per-CTA cycles (profile mode) of the DMMA tile kernel on the bench
workload and report the balance actually achieved. GPU only.  python tools/sched_fit.py [--m 2000]"""
import argparse, ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.dirname(__file__))
import torch
import itensor_b200 as itb
from itensor_b200 import synth, ITB_F64, ITB_C64
from itensor_b200._lib import lib, check
import sched_sim as S

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=2000); ap.add_argument("--nsect", type=int, default=9); ap.add_argument("--complex", action="store_true")
ap.add_argument("--flush", type=int, default=1)
a = ap.parse_args()
ctx = itb.Context(0)
sizes = synth.gaussian_sectors(a.m, a.nsect)
structs = synth.heff_chain(sizes, dtype=ITB_C64 if a.complex else ITB_F64)
hosts = [synth.random_values(s, 10 + i) for i, s in enumerate(structs)]
dts = [itb.QTensor.from_host(ctx, s, h) for s, h in zip(structs, hosts)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.device)
lib().itb_ctx_set_profile(ctx.handle, 1)
cur = dts[0]
X, Y = [], []
for k in range(4):
    p = itb.ContractPlan(cur.struct, structs[k + 1])
    out = itb.QTensor(ctx, p.C, ctx.empty(p.C.nreal))
    t, c, g = S.tiles_of(p)
    best = None
    for rep in range(4):
        if a.flush: flush.zero_()
        check(lib().itb_contract_run(ctx.handle, p._h, cur.ptr, dts[k + 1].ptr, out.ptr))
        n = lib().itb_contract_last_cta_cycles(ctx.handle, None, 0)
        cyc = np.zeros(n, dtype=np.int64)
        lib().itb_contract_last_cta_cycles(ctx.handle, cyc.ctypes.data_as(C.POINTER(C.c_int64)), n)
        best = cyc if best is None else np.minimum(best, cyc)
    cur = out
    if len(t) == 0:
        continue
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez(f"gpurun_out/sched_fit_step{k+1}.npz", tiles=t, cblks=c, cta_begin=g, cycles=best)
    M, N = c[t[:, 0], 0], c[t[:, 0], 1]
    vm = np.minimum(t[:, 3], M - t[:, 1]); vn = np.minimum(t[:, 4], N - t[:, 2])
    ch = (t[:, 6] - t[:, 5]).astype(np.float64)
    npairs = c[t[:, 0], 3]
    model = np.array([chs * S.chunk_cycles(tm, x, y) + S.OVERHEAD[tm] for chs, tm, x, y in zip(ch, t[:, 3], vm, vn)])
    mload = np.array([model[g[b]:g[b + 1]].sum() for b in range(len(g) - 1)])
    print(f"step {k+1}: measured cycles/CTA min {best.min()} mean {best.mean():.0f} max {best.max()}  (max/mean {best.max()/best.mean():.3f});"
          f" model mean {mload.mean():.0f} max {mload.max():.0f}; corr {np.corrcoef(mload, best)[0,1]:.3f}")
    # features per CTA: per cfg: items, chunks, chunks*dmma_cycles(ideal 64*worst), split pieces, pair switches
    for b in range(len(g) - 1):
        f = np.zeros(12)
        for i in range(g[b], g[b + 1]):
            ci = {128: 0, 64: 1, 32: 2}[int(t[i, 3])]
            W = int(t[i, 3]) // 4; F = W // 8
            fm = [max(0, min(F, (int(vm[i]) - q * W + 7) // 8)) for q in range(4)]
            fn = [max(0, min(F, (int(vn[i]) - q * W + 7) // 8)) for q in range(4)]
            worst = max(sum(fm[s ^ j] * fn[j] for j in range(4)) for s in range(4))
            f[ci] += 1; f[3 + ci] += ch[i]; f[6 + ci] += ch[i] * 64.0 * worst
            f[9] += 1 if t[i, 7] >= 0 else 0
            f[10] += npairs[i] if (t[i, 6] - t[i, 5]) > 0 else 0
        f[11] = 1
        X.append(f); Y.append(best[b])
X = np.array(X); Y = np.array(Y, dtype=np.float64)
names = ["item128", "item64", "item32", "chunk128", "chunk64", "chunk32", "dmma128", "dmma64", "dmma32", "split_piece", "pairs", "const"]
keep = [i for i in range(12) if X[:, i].std() > 0 or names[i] == "const"]
coef, res, rk, sv = np.linalg.lstsq(X[:, keep], Y, rcond=None)
pred = X[:, keep] @ coef
print("least-squares fit of measured cycles per CTA:")
for i, cf in zip(keep, coef):
    print(f"   {names[i]:12s} {cf:12.2f}")
print(f"   rms residual {np.sqrt(np.mean((pred - Y) ** 2)):.0f} cycles of mean {Y.mean():.0f}")
