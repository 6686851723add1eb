"""Shared helpers for the parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import numpy as np

import itensor_b200 as itb
from itensor_b200 import synth
from oracle import orc

REL_TOL = 1e-12  # north_star: element values within a relative 1e-12 (scaled by max|C|)


def assert_struct_equal(S, O):
    """block structure, offsets and index order must be bit-exact"""
    assert S.order == O.order
    assert np.array_equal(np.asarray(S.labels, np.int64), np.asarray(O.labels, np.int64))
    assert np.array_equal(S.blocks, O.blocks)
    assert np.array_equal(S.offsets, O.offsets)
    assert S.nelems == O.nelems
    assert S.dtype == O.dtype


def assert_close(got, want, tol=REL_TOL, what=""):
    got = np.asarray(got).astype(np.complex128)
    want = np.asarray(want).astype(np.complex128)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if want.size == 0:
        return
    scale = max(np.abs(want).max(), 1e-300)
    err = np.abs(got - want).max() / scale
    assert err <= tol, (what, err)


def gpu_contract_vs_oracle(ctx, A, B, seed=0, tol=REL_TOL):
    a, b = synth.random_values(A, seed), synth.random_values(B, seed + 7919)
    Cs, triples, want = orc.contract(A, a, B, b)
    tA, tB = itb.QTensor.from_host(ctx, A, a), itb.QTensor.from_host(ctx, B, b)
    plan = itb.ContractPlan(A, B)
    assert_struct_equal(plan.C, Cs)
    assert np.array_equal(plan.pairs(), triples)
    got = itb.contract(tA, tB, plan).to_host()
    assert_close(got, want, tol, "contract")
    return plan
