#!/usr/bin/env python
"""Per-step timing and per-CTA balance of the DMMA tile kernel on the bench workload (H_eff*phi at maxdim m): kernel +
split-K reduce time from profile-mode CUDA events, clock64 span of every CTA (max / mean / min), item and piece counts.
Planner knobs are read once per process: run as  ITB_GUIDED_FACTOR=2 ITB_MIN_PIECE=8 python tools/tile_probe.py"""
import argparse, ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import itensor_b200 as itb
from itensor_b200 import synth, ITB_F64, ITB_C64
from itensor_b200._lib import lib, check

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=2000); ap.add_argument("--nsect", type=int, default=9); ap.add_argument("--complex", action="store_true")
ap.add_argument("--reps", type=int, default=8)
a = ap.parse_args()
ctx = itb.Context(0)
sizes = synth.gaussian_sectors(a.m, a.nsect)
structs = synth.heff_chain(sizes, dtype=ITB_C64 if a.complex else ITB_F64)
hosts = [synth.random_values(s, 10 + i) for i, s in enumerate(structs)]
dts = [itb.QTensor.from_host(ctx, s, h) for s, h in zip(structs, hosts)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.device)
lib().itb_ctx_set_profile(ctx.handle, 1)
cur = dts[0]
tag = f"factor={os.environ.get('ITB_GUIDED_FACTOR','1')} minpiece={os.environ.get('ITB_MIN_PIECE','8')}"
tot = 0.0
for k in range(4):
    p = itb.ContractPlan(cur.struct, structs[k + 1])
    out = itb.QTensor(ctx, p.C, ctx.empty(p.C.nreal))
    ms_best, cyc_best, items_best = 1e9, None, None
    for _ in range(a.reps):
        flush.zero_()
        check(lib().itb_contract_run(ctx.handle, p._h, cur.ptr, dts[k + 1].ptr, out.ptr))
        ms = (C.c_float * 5)(); lib().itb_contract_last_ms(ctx.handle, ms)
        if p.info.n_gemm_tiles and ms[0] < ms_best:
            ms_best = ms[0]
            n = lib().itb_contract_last_cta_cycles(ctx.handle, None, 0)
            cyc = np.zeros(n, np.int64); lib().itb_contract_last_cta_cycles(ctx.handle, cyc.ctypes.data_as(C.POINTER(C.c_int64)), n)
            cyc_best = cyc
            ni = lib().itb_contract_last_item_cycles(ctx.handle, None, 0)
            it = np.zeros((ni, 4), np.int64); lib().itb_contract_last_item_cycles(ctx.handle, it.ctypes.data_as(C.POINTER(C.c_int64)), 4 * ni)
            items_best = it
    if p.info.n_gemm_tiles:
        nt = lib().itb_contract_plan_tiles(p._h, None, 0)
        t = np.zeros((nt, 8), np.int32); lib().itb_contract_plan_tiles(p._h, t.ctypes.data_as(C.POINTER(C.c_int32)), nt)
        fl = sum(p.info.class_flops[:3])
        tot += ms_best
        print(f"{tag} step {k+1}: tile kernel+reduce {ms_best*1e3:.1f} us = {fl/ms_best/1e9:.2f} TFLOP/s; items {nt} (pieces {int((t[:,7]>=0).sum())}); "
              f"CTA cycles max {cyc_best.max()} mean {cyc_best.mean():.0f} min {cyc_best.min()} (max/mean {cyc_best.max()/cyc_best.mean():.3f}; kernel alone ~{cyc_best.max()/1.965e3:.1f} us)")
        # per-item anatomy (profile build): K-loop cycles per chunk, epilogue, gap to the next item of the same CTA, by item kind
        nc = lib().itb_contract_plan_cblks(p._h, None, 0)
        cb = np.zeros((nc, 4), np.int64); lib().itb_contract_plan_cblks(p._h, cb.ctypes.data_as(C.POINTER(C.c_int64)), nc)
        it = items_best
        vm = np.minimum(t[:, 3], cb[t[:, 0], 0] - t[:, 1]); vn = np.minimum(t[:, 4], cb[t[:, 0], 1] - t[:, 2])
        nch = (t[:, 6] - t[:, 5]).astype(float)
        kind = np.where(t[:, 3] != 128, "small", np.where((vm == 128) & (vn == 128), "full", np.where((vm >= 96) & (vn >= 96), "edge>=96", "edge<96")))
        piece = np.where(t[:, 7] >= 0, "piece", "whole")
        loop = (it[:, 2] - it[:, 1]).astype(float); epi = (it[:, 3] - it[:, 2]).astype(float)
        gap = np.full(len(it), np.nan)
        for c in np.unique(it[:, 0]):
            idx = np.where(it[:, 0] == c)[0]
            idx = idx[np.argsort(it[idx, 1])]
            gap[idx[:-1]] = it[idx[1:], 1] - it[idx[:-1], 3]
        for kd in ("full", "edge>=96", "edge<96", "small"):
            for pc in ("whole", "piece"):
                sel = (kind == kd) & (piece == pc)
                if sel.sum() == 0:
                    continue
                print(f"      {kd:9s} {pc:5s} n={int(sel.sum()):5d} chunks/item {nch[sel].mean():6.1f}  K-loop cycles/chunk {np.sum(loop[sel])/np.sum(nch[sel]):7.0f}  "
                      f"epilogue {np.median(epi[sel]):6.0f}  gap to next item {np.nanmedian(gap[sel]):6.0f}  share of busy time {np.sum(loop[sel]+epi[sel])/np.sum(loop+epi):.3f}")
        busy = np.array([np.sum((loop + epi)[it[:, 0] == c]) for c in range(len(cyc_best))])
        print(f"      per CTA: busy (K loops + epilogues) mean {busy.mean():.0f} of span {cyc_best.mean():.0f} cycles -> between-item gaps {1 - busy.mean()/cyc_best.mean():.3f}")
    cur = out
print(f"{tag} TOTAL tile class {tot*1e3:.1f} us per H_eff*phi")
