#!/usr/bin/env python
"""Schedule analysis of the DMMA tile queue (host only): padded vs algorithmic work, and a simulation of the kernel's
dynamic in-order queue (each free CTA takes the next item) under a deliberately WRONG cost model (per-item multiplicative
noise), i.e. how much balance depends on the model. Usage: python tools/sched_sim.py [--m 2000] [--nsect 9] [--complex]"""
import argparse, ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import itensor_b200 as itb
from itensor_b200 import synth, ITB_F64, ITB_C64
from itensor_b200._lib import lib

def tiles_of(plan):
    n = lib().itb_contract_plan_tiles(plan._h, None, 0)
    t = np.zeros((n, 8), dtype=np.int32)
    lib().itb_contract_plan_tiles(plan._h, t.ctypes.data_as(C.POINTER(C.c_int32)), n)
    nc = lib().itb_contract_plan_cblks(plan._h, None, 0)
    c = np.zeros((nc, 4), dtype=np.int64)
    lib().itb_contract_plan_cblks(plan._h, c.ctypes.data_as(C.POINTER(C.c_int64)), nc)
    ng = lib().itb_contract_plan_cta_begin(plan._h, None, 0)
    g = np.zeros(ng, dtype=np.int32)
    lib().itb_contract_plan_cta_begin(plan._h, g.ctypes.data_as(C.POINTER(C.c_int32)), ng)
    return t, c, g

FLOOR = {128: 2950.0, 64: 1700.0, 32: 1050.0}
OVERHEAD = {128: 3800.0, 64: 8000.0, 32: 5300.0}
PAIR_OVERHEAD = 1000.0

def chunk_cycles(T, vm, vn):
    W = T // 4; F = W // 8
    fm = [max(0, min(F, (vm - i * W + 7) // 8)) for i in range(4)]
    fn = [max(0, min(F, (vn - i * W + 7) // 8)) for i in range(4)]
    worst = max(sum(fm[s ^ j] * fn[j] for j in range(4)) for s in range(4))
    if T == 128:
        return 2466.0 + 0.54 * 64.0 * worst
    return max(1.11 * 64.0 * worst, FLOOR[T])

def analyse(plan, G=148, label=""):
    t, c, g = tiles_of(plan)
    if len(t) == 0:
        return
    M, N = c[t[:, 0], 0], c[t[:, 0], 1]
    vm = np.minimum(t[:, 3], M - t[:, 1]); vn = np.minimum(t[:, 4], N - t[:, 2])
    ch = (t[:, 6] - t[:, 5]).astype(np.float64)
    useful = (vm * vn * ch * 16).sum() * 2
    padded = (t[:, 3] * t[:, 4] * ch * 16).sum() * 2
    # cost model: time of an item ~ chunks * tile area / eff + fixed overhead per item (epilogue/prologue ~ 2 chunks)
    cost = np.array([chs * chunk_cycles(tm, a, b) + OVERHEAD[tm] for chs, tm, a, b in zip(ch, t[:, 3], vm, vn)])
    ideal = (vm * vn * ch).sum() * 16 * 2 / 128.0 / G  # cycles at the bare DMMA rate (128 flop/cycle/SM)
    print(f"{label}: items {len(t)} (pieces of cut tiles {int((t[:,7]>=0).sum())}), cfg hist {dict(zip(*np.unique(t[:,3]*1000+t[:,4], return_counts=True)))}")
    import heapq
    rng = np.random.default_rng(0)
    for noise in (0.0, 0.3):
        real = cost * np.exp(rng.normal(0, noise, len(cost))) if noise else cost
        # hybrid schedule: every CTA first runs its static range, then the CTA that frees up first takes the next queue item
        free = [(float(real[g[b]:g[b + 1]].sum()), b) for b in range(G)]
        heapq.heapify(free)
        for v in real[g[G]:g[G + 1]]:
            tfree, b = heapq.heappop(free)
            heapq.heappush(free, (tfree + v, b))
        ends = np.array([x for x, _ in free])
        print(f"   static ranges + dynamic queue, model noise {noise:.1f}: makespan / mean CTA busy time {ends.max() / (real.sum() / G):.3f}; "
              f"useful/padded flops {useful/padded:.3f}; modelled fraction of the DMMA peak {ideal / ends.max():.3f}")

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=2000); ap.add_argument("--nsect", type=int, default=9); ap.add_argument("--complex", action="store_true")
    a = ap.parse_args()
    sizes = synth.gaussian_sectors(a.m, a.nsect)
    structs = synth.heff_chain(sizes, dtype=ITB_C64 if a.complex else ITB_F64)
    print("sectors", sizes)
    s = structs[0]
    for k, tt in enumerate(structs[1:]):
        p = itb.ContractPlan(s, tt)
        analyse(p, label=f"step {k+1} ({p.npairs} pairs, {p.flops/1e9:.2f} GFLOP)")
        s = p.C
