#!/usr/bin/env python
"""Host planning cost at Hubbard scale (no GPU needed): a 2-component (Nf,Sz) QN structure shaped like sample/hubbard_2d.cc's
centre bond (SURVEY 8a census: 26 link sectors, MPO bond 18 in 5 sectors, 4 one-dimensional site sectors), all four steps of
LocalOp::product. Prints blocks / pairs and milliseconds per itb_contract_plan_create (+ table build)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import itensor_b200 as itb
from itensor_b200.tensor import BlockStruct, Index, flux_blocks
from itensor_b200._lib import lib, check, ContractInfo

def link(idn, sizes, qns, dirn): return Index(idn, tuple(sizes), tuple(qns), dirn, (1, 1))
def structs(scale=1):
    # (Nf,Sz) sectors of a 64-site half-filled Hubbard MPS bond: Nf around 32, Sz in -3..3
    rng = np.random.default_rng(0)
    qn, sz = [], []
    for nf in range(29, 36):
        for s in range(-3, 4):
            if (nf + s) % 2 == 0:
                qn.append((nf, s)); sz.append(int(max(1, scale * rng.integers(1, 23) * np.exp(-((nf - 32) ** 2 + s * s) / 6.0))))
    l = link(1, sz, qn, -1); r = link(2, sz, [(a + 2, b) for a, b in qn], +1)
    site = lambda i, d: Index(i, (1, 1, 1, 1), ((0, 0), (1, 1), (1, -1), (2, 0)), d, (1, 1))
    s1, s2 = site(3, +1), site(4, +1)
    kq = ((0, 0), (1, 1), (1, -1), (-1, -1), (-1, 1)); ks = (2, 4, 4, 4, 4)
    k0, k1, k2 = (Index(i, ks, kq, +1, (1, 1)) for i in (5, 6, 7))
    full = lambda inds, flux=(0, 0): BlockStruct(inds, flux_blocks(inds, flux))
    phi = full([l, s1, s2, r])
    return phi, full([l.dag(), k0, l.prime()]), full([k0.dag(), s1.dag(), s1.prime(), k1]), full([k1.dag(), s2.dag(), s2.prime(), k2]), full([r.dag(), k2.dag(), r.prime()])

for scale in (1, 20):
    st = structs(scale)
    cur = st[0]
    print(f"scale {scale}: link dim {sum(st[0].inds[0].sizes)}, sectors {st[0].inds[0].nsect}")
    for k, t in enumerate(st[1:]):
        t0 = time.perf_counter(); reps = 20
        for _ in range(reps):
            p = itb.ContractPlan(cur, t)
        dt_py = (time.perf_counter() - t0) / reps
        # C calls only (what the plugin pays): create + info (table build)
        la = np.arange(cur.order, dtype=np.int32); key = {(i.id, i.plev): j for j, i in enumerate(cur.inds)}
        lb = np.array([key.get((i.id, i.plev), 100 + j) for j, i in enumerate(t.inds)], np.int32)
        da, db = cur.desc(), t.desc()
        t0 = time.perf_counter()
        for _ in range(reps):
            h = C.c_void_p(); check(lib().itb_contract_plan_create(C.byref(da), la.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(db), lb.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(h)))
            info = ContractInfo(); check(lib().itb_contract_plan_info(h, C.byref(info)))
            t1 = time.perf_counter()
            lib().itb_contract_plan_destroy(h)
        dt_c = (time.perf_counter() - t0) / reps
        print(f"   step {k+1}: A blocks {cur.nblocks} B blocks {t.nblocks} pairs {p.npairs} C blocks {p.C.nblocks} tiles {p.info.n_gemm_tiles} skinny {p.info.n_skinny}: "
              f"C planner {dt_c*1e3:.3f} ms ({dt_c/max(p.npairs,1)*1e6:.2f} us/pair), python mirror {dt_py*1e3:.2f} ms")
        cur = p.C
