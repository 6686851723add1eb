"""Deterministic case lists shared by tests/golden/make_golden.py (reference side) and
tests/test_golden.py / tests/test_oracle_vs_reference.py (oracle and CUDA side)."""
import numpy as np

from itensor_b200 import ITB_C64, ITB_F64, synth
from itensor_b200.tensor import BlockStruct, Index

F, Z = ITB_F64, ITB_C64


def contract_cases():
    rng = np.random.default_rng(20260101)
    cases = []
    k = 0
    while len(cases) < 24:
        ra, rb = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        nc = int(rng.integers(0, min(ra, rb) + 1))
        da, db = [(F, F), (Z, F), (F, Z), (Z, Z)][k % 4]
        A, B = synth.random_qn_pair(rng, ra, rb, nc, dtype_a=da, dtype_b=db, max_size=4)
        k += 1
        if A.nblocks == 0 or B.nblocks == 0:
            continue
        cases.append((f"qn{len(cases):02d}", A, B))
    # the four H_eff*phi operand pairs of a small chain (LocalOp::product), real and complex
    for tag, dt in (("r", F), ("z", Z)):
        st = synth.heff_chain([1, 3, 4, 2], [2, 4, 3, 1], dtype=dt)
        cases.append((f"heff1{tag}", st[0], st[1]))
    # dense (no QN) ITensors, all transpose layouts + non-matrix reshapes (contract_test.cc)
    ix = {l: Index(l, (d,)) for l, d in [(1, 7), (2, 5), (3, 6), (4, 3), (5, 4)]}
    for n, (la, lb) in enumerate([((1, 2), (2, 3)), ((2, 1), (2, 3)), ((1, 2), (3, 2)), ((2, 1), (3, 2)), ((1, 2, 3), (2, 4, 1)),
                                  ((4, 1, 5), (5, 4)), ((1, 2, 3), (1, 2, 3))]):
        for tag, (da, db) in (("rr", (F, F)), ("zz", (Z, Z)), ("rz", (F, Z))):
            cases.append((f"dense{n}{tag}", BlockStruct.dense([ix[l] for l in la], da), BlockStruct.dense([ix[l] for l in lb], db)))
    return cases


def permute_cases():
    rng = np.random.default_rng(20260102)
    cases = []
    while len(cases) < 16:
        r = int(rng.integers(2, 6))
        dt = F if len(cases) % 2 == 0 else Z
        S, _ = synth.random_qn_pair(rng, r, 1, 0, max_sect=3, max_size=5, dtype_a=dt, drop=0.3)
        if S.nblocks == 0:
            continue
        new_inds = [S.inds[i] for i in rng.permutation(r)]
        if all(a.same(b) for a, b in zip(new_inds, S.inds)):
            continue  # trivial permutation: the reference returns early without filling in blocks
        cases.append((f"perm{len(cases):02d}", S, new_inds))
    return cases


def pluseq_cases():
    rng = np.random.default_rng(20260103)
    cases = []
    while len(cases) < 16:
        r = int(rng.integers(1, 5))
        da, db = [(F, F), (Z, Z), (Z, F), (F, Z)][len(cases) % 4]
        full, _ = synth.random_qn_pair(rng, r, 1, 0, max_sect=3, max_size=4, drop=0.0)
        if full.nblocks < 2:
            continue

        def sub(dtype):
            keep = rng.uniform(size=full.nblocks) < 0.7
            if not keep.any():
                keep[0] = True
            return BlockStruct(full.inds, full.blocks[keep], dtype)

        A, Bn = sub(da), sub(db)
        order = rng.permutation(r)
        from itensor_b200.tensor import permuted_struct

        B, _ = permuted_struct(Bn, [Bn.inds[i] for i in order])
        alpha = complex(round(float(rng.uniform(-2, 2)), 3), round(float(rng.uniform(-2, 2)), 3) if da == Z else 0.0)
        cases.append((f"add{len(cases):02d}", A, B, alpha))
    return cases
