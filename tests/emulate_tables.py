"""Run in a subprocess by tests/test_tables_emulation_cpu.py with ITB200_LIB_PATH=<mock>, ITB_MOCK_TABLES=1: every
contraction is executed by walking the planner's DEVICE tables on the host (oracle/mock_itb200.cc emu_contract) and
compared with the CPU oracle (definition-level loops). Prints one summary line; exits non-zero on the first mismatch."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensor_b200 as itb
from itensor_b200 import ITB_C64, ITB_F64, synth
from itensor_b200._lib import check, lib
from itensor_b200.tensor import BlockStruct, Index
from oracle import orc

assert "mock" in os.environ.get("ITB200_LIB_PATH", "") and os.environ.get("ITB_MOCK_TABLES") == "1"
ctx = C.c_void_p()
check(lib().itb_ctx_create(0, C.byref(ctx)))
stats = {"cases": 0, "tiles": 0, "split_pieces": 0, "rowgroups": 0, "skinny": 0, "dots": 0}


def classes(P):
    n = lib().itb_contract_plan_tiles(P._h, None, 0)
    t = np.zeros((max(n, 1), 8), np.int32)
    lib().itb_contract_plan_tiles(P._h, t.ctypes.data_as(C.POINTER(C.c_int32)), n)
    stats["tiles"] += n
    stats["split_pieces"] += int((t[:n, 7] >= 0).sum())
    stats["rowgroups"] += lib().itb_contract_plan_rowgroups(P._h, None, 0)
    stats["skinny"] += P.info.n_skinny
    stats["dots"] += P.info.n_dot


def run(A, Av, B, Bv, what):
    P = itb.ContractPlan(A, B)
    Cs, tr, ref = orc.contract(A, Av, B, Bv)
    assert np.array_equal(Cs.blocks, P.C.blocks) and np.array_equal(Cs.offsets, P.C.offsets)
    out = np.full(max(P.C.nreal, 1), np.nan)
    a = np.ascontiguousarray(Av).view(np.float64).reshape(-1)
    b = np.ascontiguousarray(Bv).view(np.float64).reshape(-1)
    want = np.ascontiguousarray(ref).view(np.float64).reshape(-1)
    scale = max(np.abs(want).max(), 1e-300) if P.C.nreal else 1.0
    # three executions: the planner is tiered (plan.cc) — a small streaming class starts on the C-stationary kernels and is
    # re-planned with row groups at the plan's third execution; both sets of tables are walked and checked
    for nrun in range(3):
        out[:] = np.nan
        check(lib().itb_contract_run(ctx, P._h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        got = out[:P.C.nreal]
        if P.C.nreal:
            err = np.abs(got - want).max() / scale if not np.isnan(got).any() else np.inf
            if not err < 1e-12:
                print(f"MISMATCH in {what} (execution {nrun + 1}): rel err {err}, nan {int(np.isnan(got).sum())} of {got.size}")
                sys.exit(1)
        if nrun == 0:
            stats["skinny_first"] = stats.get("skinny_first", 0) + P.info.n_skinny
        check(lib().itb_contract_plan_info(P._h, C.byref(P.info)))
    classes(P)
    stats["cases"] += 1
    if P.info.n_gemm_tiles > 0:
        # re-cut the tile partition with perturbed per-tile costs (what itb_contract_plan_refine does from measured cycles;
        # the mock perturbs pseudo-randomly) and walk the re-cut tables: same result
        out2 = np.full(max(P.C.nreal, 1), np.nan)
        check(lib().itb_contract_plan_refine(ctx, P._h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), out2.ctypes.data_as(C.c_void_p), 1, None))
        err2 = np.abs(out2[:P.C.nreal] - want).max() / scale if not np.isnan(out2[:P.C.nreal]).any() else np.inf
        if not err2 < 1e-12:
            print(f"MISMATCH after re-partition in {what}: rel err {err2}")
            sys.exit(1)
        stats["refined"] = stats.get("refined", 0) + 1
    return P.C, ref


rng = np.random.default_rng(11)
# random block-deficient QN pairs of every real/complex pairing (tiles 32x32, dots, streaming kernels, complex folds)
for trial in range(160):
    ra, rb = int(rng.integers(0, 5)), int(rng.integers(0, 5))
    nc = int(rng.integers(0, min(ra, rb) + 1))
    A, B = synth.random_qn_pair(rng, ra, rb, nc, dtype_a=int(rng.integers(0, 2)), dtype_b=int(rng.integers(0, 2)))
    run(A, synth.random_values(A, 2 * trial), B, synth.random_values(B, 2 * trial + 1), f"random pair {trial}")
# H_eff*phi chains: real (row groups for the MPO steps), complex (C-stationary streaming kernels), and <phi|H phi> dots
for sizes, dt in (([3, 9, 14, 8, 2], ITB_F64), ([3, 9, 14, 8, 2], ITB_C64), ([20, 70, 45], ITB_F64)):
    structs = synth.heff_chain(sizes, dtype=dt)
    vals = [synth.random_values(s, 50 + i) for i, s in enumerate(structs)]
    cur, cv = structs[0], vals[0]
    for k in range(1, 5):
        cur, cv = run(cur, cv, structs[k], vals[k], f"heff {sizes} dtype {dt} step {k}")
    # scalar product with the (conjugated-structure) bra: rank-0 result through the split-K dot items
    bra = BlockStruct([i.dag().prime(0) if False else i for i in cur.inds], cur.blocks, cur.dtype)
    braket_inds = [Index(ix.id, ix.sizes, ix.qns, -ix.dir, ix.mods, ix.plev) for ix in cur.inds]
    bra = BlockStruct(braket_inds, cur.blocks, cur.dtype)
    run(bra, synth.random_values(bra, 99), cur, cv, f"heff {sizes} dtype {dt} dot")
# complex state, real operators (a real Hamiltonian applied to a complex phi): the MPO steps stream through the ROW-GROUP kernel
# with the complex operand read as a real one of doubled leading extent
# ... and the other pairings: complex weights (complex operators: input / output slots are the (re,im) components, signed
# weights) and real phi with complex operators (strided C rows)
for sizes in ([3, 9, 14, 8, 2], [20, 70, 45]):
    sc, sr = synth.heff_chain(sizes, dtype=ITB_C64), synth.heff_chain(sizes, dtype=ITB_F64)
    for what, structs in (("complex phi x real operators", (sc[0],) + tuple(sr[1:])), ("complex phi x complex operators", sc),
                          ("real phi x complex operators", (sr[0],) + tuple(sc[1:]))):
        vals = [synth.random_values(s, 70 + i) for i, s in enumerate(structs)]
        cur, cv = structs[0], vals[0]
        before = stats["rowgroups"]
        for k in range(1, 5):
            cur, cv = run(cur, cv, structs[k], vals[k], f"heff {what} {sizes} step {k}")
        assert stats["rowgroups"] > before, f"expected row groups for {what}"
# long K on few tiles: the stream-K partition must cut tiles into pieces (workspace slots + ordered reduction)
for M, K, N, dt in ((64, 4096, 64, ITB_F64), (150, 3000, 40, ITB_F64), (40, 2500, 33, ITB_C64)):
    im, ik, inn = Index(1, (M,)), Index(2, (K,)), Index(3, (N,))
    A, B = BlockStruct.dense([ik, im], dt), BlockStruct.dense([inn, ik], dt)
    before = stats["split_pieces"]
    run(A, synth.random_values(A, 7), B, synth.random_values(B, 8), f"dense long-K {M}x{K}x{N}")
    assert stats["split_pieces"] > before, "expected cut tiles"
# itb_contract_run_mirrored (multi-GPU row exchange from the epilogue): on the mock the "peer copies" are plain host buffers
A, B = BlockStruct.dense([Index(2, (300,)), Index(1, (70,))], ITB_F64), BlockStruct.dense([Index(3, (50,)), Index(2, (300,))], ITB_F64)
av, bv = synth.random_values(A, 5), synth.random_values(B, 6)
P = itb.ContractPlan(A, B)
_, _, ref = orc.contract(A, av, B, bv)
out = np.full(P.C.nreal, np.nan)
peers = [np.full(P.C.nreal, np.nan) for _ in range(2)]
arr = (C.c_void_p * 2)(*[p_.ctypes.data for p_ in peers])
check(lib().itb_contract_run_mirrored(ctx, P._h, av.ctypes.data_as(C.c_void_p), bv.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), 2, arr))
want = np.ascontiguousarray(ref).view(np.float64).reshape(-1)
assert np.abs(out - want).max() <= 1e-12 * np.abs(want).max() and all(np.array_equal(p_, out) for p_ in peers)
# ---- permute / permuting accumulate: per-work-item records (4096-element chunks, PT x PT tiles, zero-fill items) ----
from itensor_b200.tensor import PermutePlan, permuted_struct

stats.update({"permutes": 0, "zero_filled": 0})


def run_permute(S, host, D, perm, alpha=1.0 + 0j, accumulate=False, d_host=None, what=""):
    want = orc.permute(S, host, D, perm, alpha=alpha, accumulate=accumulate, d_host=d_host)
    pp = PermutePlan(S, D, perm)
    nd = D.nreal
    out = np.full(max(nd, 1), np.nan)
    if accumulate:
        out[:nd] = np.ascontiguousarray(d_host).view(np.float64).reshape(-1)
    src = np.ascontiguousarray(host).view(np.float64).reshape(-1)
    alpha = complex(alpha)
    check(lib().itb_permute_run(ctx, pp._h, src.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), alpha.real, alpha.imag, 1 if accumulate else 0))
    got = out[:nd]
    w = np.ascontiguousarray(want).view(np.float64).reshape(-1)
    ok = np.array_equal(got, w) if alpha == 1 and not accumulate else np.allclose(got, w, rtol=1e-14, atol=1e-14)
    if not ok:
        print(f"PERMUTE MISMATCH in {what}: nan {int(np.isnan(got).sum())} of {got.size}")
        sys.exit(1)
    stats["permutes"] += 1
    stats["zero_filled"] += int(D.nelems > S.nelems and not accumulate)


for dtype in (ITB_F64, ITB_C64):
    prng = np.random.default_rng(3 + dtype)
    for trial in range(40):  # QN permutes fill in every flux-allowed block (zero-fill items / memset ranges)
        r = int(prng.integers(1, 6))
        S, _ = synth.random_qn_pair(prng, r, 1, 0, max_sect=3, max_size=7, dtype_a=dtype, drop=0.2)
        new_inds = [S.inds[i] for i in prng.permutation(r)]
        D, perm = permuted_struct(S, new_inds, flux=(0,))
        host = synth.random_values(S, trial)
        run_permute(S, host, D, perm, what=f"qn permute {trial}")
        # accumulate into existing destination data: blocks without a source must stay untouched
        base = synth.random_values(D, 1000 + trial)
        run_permute(S, host, D, perm, alpha=0.5 - (0.25j if dtype == ITB_C64 else 0), accumulate=True, d_host=base, what=f"qn accumulate {trial}")
    for dims in [(7,), (33, 65), (65, 33), (5, 6, 7), (40, 3, 50), (3, 40, 2, 50), (2, 3, 4, 5, 6), (64, 64, 9), (1, 70, 1, 90), (130, 131)]:
        inds = [Index(10 + j, (d,)) for j, d in enumerate(dims)]
        S = BlockStruct.dense(inds, dtype)
        host = synth.random_values(S, 1)
        for _ in range(3):
            new_inds = [inds[i] for i in prng.permutation(len(dims))]
            D, perm = permuted_struct(S, new_inds)
            run_permute(S, host, D, perm, what=f"dense permute {dims}")
print("tables emulation ok:", stats)
