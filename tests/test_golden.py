"""Replay of the committed golden vectors (tests/golden/golden_v1.npz, produced by the UNMODIFIED
reference through tests/golden/make_golden.py):
  * CPU: the oracle restatement must reproduce them      -> pins the oracle
  * GPU: the CUDA path through the C ABI must reproduce them (block structure, offsets, index order
    bit-exact; element values within relative 1e-12)
"""
import os

import numpy as np
import pytest

import itensor_b200 as itb
from cases import contract_cases, permute_cases, pluseq_cases
from itensor_b200 import synth
from itensor_b200.tensor import permuted_struct
from oracle import orc
from util import assert_close

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def _check_struct(name, labels, blocks, offsets, is_qn=True):
    assert np.array_equal(np.asarray(labels, np.int64), G[f"{name}/labels"]), name
    if is_qn:
        assert np.array_equal(blocks, G[f"{name}/blocks"]), name
        assert np.array_equal(offsets, G[f"{name}/offsets"]), name


def test_oracle_contract_matches_golden():
    for name, A, B in contract_cases():
        a, b = synth.random_values(A, 1), synth.random_values(B, 2)
        Cs, tr, val = orc.contract(A, a, B, b)
        is_qn, dtype, nelems = G[f"{name}/is_qn"]
        if Cs.order > 0:
            _check_struct(name, Cs.labels, Cs.blocks, Cs.offsets, bool(is_qn))
        assert Cs.nelems == nelems or (Cs.order == 0 and nelems == 1), name
        assert_close(val, G[f"{name}/data"], 1e-12, name)


def test_oracle_permute_and_norm_match_golden():
    for name, S, new_inds in permute_cases():
        s = synth.random_values(S, 3)
        D, perm = permuted_struct(S, new_inds, flux=(0,))
        _check_struct(name, [i.label for i in D.inds], D.blocks, D.offsets)
        assert np.array_equal(orc.permute(S, s, D, perm), G[f"{name}/data"]), name
        assert abs(orc.nrm2(s) - G[f"{name}/norm"][0]) <= 1e-14 * G[f"{name}/norm"][0]


def _add_oracle(A, a, B, b, alpha):
    """widen A to the merged block list (reference merge, qdense.cc:586-653), then accumulate"""
    perm = [[j for j, aj in enumerate(A.inds) if aj.same(ix)][0] for ix in B.inds]
    pb = np.zeros_like(B.blocks)
    for i, p in enumerate(perm):
        pb[:, p] = B.blocks[:, i]
    from itensor_b200.tensor import BlockStruct, _merge_blocks

    have = {tuple(r) for r in A.blocks.tolist()}
    missing = any(tuple(r) not in have for r in pb.tolist())
    cplx = A.is_complex or B.is_complex or alpha.imag != 0
    blocks = _merge_blocks(A.blocks, pb) if missing else A.blocks
    W = BlockStruct(A.inds, blocks, itb.ITB_C64 if cplx else itb.ITB_F64)
    base = orc.permute(A, a, W, list(range(A.order)))
    return W, orc.permute(B, b, W, perm, alpha=alpha, accumulate=True, d_host=base)


def test_oracle_pluseq_matches_golden():
    for name, A, B, alpha in pluseq_cases():
        a, b = synth.random_values(A, 4), synth.random_values(B, 5)
        W, val = _add_oracle(A, a, B, b, alpha)
        _check_struct(name, [i.label for i in W.inds], W.blocks, W.offsets)
        assert int(W.is_complex) == int(G[f"{name}/dtype"][0]), name
        assert_close(val, G[f"{name}/data"], 1e-14, name)


@pytest.mark.gpu
def test_gpu_contract_matches_golden(ctx):
    for name, A, B in contract_cases():
        a, b = synth.random_values(A, 1), synth.random_values(B, 2)
        plan = itb.ContractPlan(A, B)
        is_qn, dtype, nelems = G[f"{name}/is_qn"]
        if plan.C.order > 0:
            _check_struct(name, plan.C.labels, plan.C.blocks, plan.C.offsets, bool(is_qn))
        got = itb.contract(itb.QTensor.from_host(ctx, A, a), itb.QTensor.from_host(ctx, B, b), plan).to_host()
        assert_close(got, G[f"{name}/data"], 1e-12, name)


@pytest.mark.gpu
def test_gpu_permute_norm_pluseq_match_golden(ctx):
    for name, S, new_inds in permute_cases():
        s = synth.random_values(S, 3)
        t = itb.QTensor.from_host(ctx, S, s)
        out = itb.permute(t, new_inds, flux=(0,))
        _check_struct(name, [i.label for i in out.inds], out.struct.blocks, out.struct.offsets)
        assert np.array_equal(out.to_host(), G[f"{name}/data"]), name
        assert abs(itb.norm(t) - G[f"{name}/norm"][0]) <= 1e-13 * G[f"{name}/norm"][0]
    for name, A, B, alpha in pluseq_cases():
        a, b = synth.random_values(A, 4), synth.random_values(B, 5)
        out = itb.add(itb.QTensor.from_host(ctx, A, a), alpha, itb.QTensor.from_host(ctx, B, b))
        _check_struct(name, [i.label for i in out.inds], out.struct.blocks, out.struct.offsets)
        assert int(out.struct.is_complex) == int(G[f"{name}/dtype"][0]), name
        assert_close(out.to_host(), G[f"{name}/data"], 1e-14, name)
