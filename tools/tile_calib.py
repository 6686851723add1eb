#!/usr/bin/env python
"""Calibrate the planner's cycle model of the DMMA tile kernel (plan.cc: g_tile_floor / kTileOverhead) on a GPU.
For each forced tile configuration a dense A(M,K)*B(K,N) with exactly 4 tiles per SM is timed at two K values:
cycles/chunk = slope, per-item overhead = intercept. Edge cases (M = 16T + r) check the fragment-skipping model.
Run one configuration per process: ITB_FORCE_CFG=f python tools/tile_calib.py"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import itensor_b200 as itb
from itensor_b200 import Index, BlockStruct
from itensor_b200._lib import lib, check

f = int(os.environ.get("ITB_FORCE_CFG", "0"))
T = [128, 64, 32][f]
ctx = itb.Context(0)
clk = 1.965e9
lib().itb_ctx_set_profile(ctx.handle, 1)

def run(M, N, K, layout="nn"):
    im, ik, in_ = Index(1, (M,)), Index(2, (K,)), Index(3, (N,))
    A = BlockStruct.dense([im, ik] if layout[0] == "n" else [ik, im])
    B = BlockStruct.dense([ik, in_] if layout[1] == "n" else [in_, ik])
    p = itb.ContractPlan(A, B)
    a = torch.randn(A.nreal, dtype=torch.float64, device=ctx.device)
    b = torch.randn(B.nreal, dtype=torch.float64, device=ctx.device)
    c = ctx.empty(p.C.nreal)
    best = 1e9
    for _ in range(6):
        check(lib().itb_contract_run(ctx.handle, p._h, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr())))
        ms = (C.c_float * 5)()
        lib().itb_contract_last_ms(ctx.handle, ms)
        best = min(best, ms[0])
    ref = (a.view(K, M).T if layout[0] == "n" else a.view(M, K)) @ (b.view(N, K).T if layout[1] == "n" else b.view(K, N))
    # ("n": the operand's first (fastest) index is its row index m / k; "t": transposed storage)
    err = float((c.view(N, M).T - ref).abs().max() / ref.abs().max())
    return best, p.info.n_gemm_tiles, err

for layout in ("nn", "tn", "nt", "tt"):
    res = {}
    for per_cta in (4, 8):
        for K in (256, 1024):
            ms, nt, err = run(4 * per_cta * T, 37 * T, K, layout)
            res[(per_cta, K)] = ms * 1e-3 * clk  # cycles per CTA
            print(f"cfg {f} T={T} layout {layout} {per_cta} full tiles/CTA K={K}: {ms*1e3:.1f} us, items {nt}, {2*4*per_cta*T*37*T*K/ms/1e9:.2f} TFLOP/s, err {err:.1e}")
    # cycles(per_cta, K) = launch + per_cta * (item + chunks * w)
    w = (res[(8, 1024)] - res[(8, 256)]) / (8 * (1024 - 256) / 16)
    item = (res[(8, 256)] - res[(4, 256)]) / 4 - w * 16
    launch = res[(4, 256)] - 4 * (item + 16 * w)
    print(f"   => cycles/chunk {w:.0f}, overhead/item {item:.0f}, per launch {launch:.0f}")
for r in (8, 40, 72, 104):
    if r >= T:
        continue
    ms, nt, err = run(16 * T + r, 37 * T, 1024)
    ms0, _, _ = run(16 * T, 37 * T, 1024)
    print(f"cfg {f} edge +{r} rows: {ms*1e3:.1f} us vs {ms0*1e3:.1f} us full-only ({nt} items), err {err:.1e}")
