"""GPU parity: permute / permuting accumulate / flat-store tasks vs the CPU oracle."""
import numpy as np
import pytest

import itensor_b200 as itb
from itensor_b200 import synth
from itensor_b200.tensor import BlockStruct, Index, PermutePlan, permuted_struct
from oracle import orc
from util import assert_close

pytestmark = pytest.mark.gpu
F, Z = itb.ITB_F64, itb.ITB_C64


def _rand_qn_tensor(rng, order, dtype, max_size=7, max_sect=3, drop=0.2):
    A, _ = synth.random_qn_pair(rng, order, 1, 0, max_sect=max_sect, max_size=max_size, dtype_a=dtype, drop=drop)
    return A


@pytest.mark.parametrize("dtype", [F, Z])
def test_permute_qn_fills_all_flux_blocks(ctx, dtype):
    """permuteQDense allocates EVERY flux-allowed block (qdense.cc:862)"""
    rng = np.random.default_rng(3 + dtype)
    for trial in range(40):
        r = int(rng.integers(1, 6))
        S = _rand_qn_tensor(rng, r, dtype)
        new_inds = [S.inds[i] for i in rng.permutation(r)]
        host = synth.random_values(S, trial)
        D, perm = permuted_struct(S, new_inds, flux=(0,))
        assert np.array_equal(D.blocks, orc.flux_blocks(D.inds, (0,)))
        want = orc.permute(S, host, D, perm)
        got = itb.permute(itb.QTensor.from_host(ctx, S, host), new_inds, flux=(0,))
        assert np.array_equal(got.struct.blocks, D.blocks) and np.array_equal(got.struct.offsets, D.offsets)
        assert np.array_equal(got.to_host(), want)  # a permute moves bits: exact


@pytest.mark.parametrize("dtype", [F, Z])
def test_permute_dense(ctx, dtype):
    rng = np.random.default_rng(11)
    for dims in [(7,), (33, 65), (65, 33), (5, 6, 7), (40, 3, 50), (3, 40, 2, 50), (2, 3, 4, 5, 6), (64, 64, 9), (1, 70, 1, 90)]:
        r = len(dims)
        inds = [Index(10 + j, (d,)) for j, d in enumerate(dims)]
        S = BlockStruct.dense(inds, dtype)
        host = synth.random_values(S, 1)
        for _ in range(3):
            new_inds = [inds[i] for i in rng.permutation(r)]
            D, perm = permuted_struct(S, new_inds)
            want = orc.permute(S, host, D, perm)
            got = itb.permute(itb.QTensor.from_host(ctx, S, host), new_inds)
            assert np.array_equal(got.to_host(), want)
            # cross-check against numpy's own transpose
            arr = host.reshape(dims, order="F")
            axes = [list(inds).index(i) for i in new_inds]
            assert np.array_equal(np.transpose(arr, axes).reshape(-1, order="F"), want)


@pytest.mark.parametrize("da,db", [(F, F), (Z, Z), (Z, F), (F, Z)])
def test_add_with_permutation_and_block_merge(ctx, da, db):
    """A += alpha*B: permuted index order, blocks of B missing in A, real+complex promotion
    (qdense.cc:551-668; itensor_test.cc SumDifference :808-1027, block-deficient :2734-2822)"""
    rng = np.random.default_rng(17 + da + 2 * db)
    for trial in range(30):
        r = int(rng.integers(1, 5))
        full = _rand_qn_tensor(rng, r, F, drop=0.0)
        if full.nblocks == 0:
            continue
        def sub(dtype):
            keep = rng.uniform(size=full.nblocks) < 0.7
            if not keep.any():
                keep[0] = True
            return BlockStruct(full.inds, full.blocks[keep], dtype)
        A = sub(da)
        Bn = sub(db)
        order = rng.permutation(r)
        b_inds = [Bn.inds[i] for i in order]
        B, _ = permuted_struct(Bn, b_inds)  # same blocks, permuted index order
        a, b = synth.random_values(A, trial), synth.random_values(B, 99 + trial)
        alpha = complex(rng.uniform(-2, 2), rng.uniform(-2, 2) if (da == Z and trial % 2) else 0.0)
        out = itb.add(itb.QTensor.from_host(ctx, A, a), alpha, itb.QTensor.from_host(ctx, B, b))
        # oracle: widen A to the merged block list, then accumulate alpha*perm(B)
        want_struct = out.struct
        perm = [[j for j, aj in enumerate(A.inds) if aj.same(ix)][0] for ix in B.inds]
        base = orc.permute(A, a, want_struct, list(range(r)))
        want = orc.permute(B, b, want_struct, perm, alpha=alpha, accumulate=True, d_host=base)
        # merged block list == sorted union (reference merge, qdense.cc:586-639)
        pb = np.zeros_like(B.blocks)
        for i, p in enumerate(perm):
            pb[:, p] = B.blocks[:, i]
        union = {tuple(x) for x in A.blocks.tolist()} | {tuple(x) for x in pb.tolist()}
        assert {tuple(x) for x in want_struct.blocks.tolist()} == union
        assert want_struct.is_complex == (da == Z or db == Z or alpha.imag != 0)
        assert_close(out.to_host(), want, 1e-15, "add")


def test_blas1(ctx):
    rng = np.random.default_rng(0)
    for dtype in (F, Z):
        for n in (1, 2, 3, 255, 4097, 1_000_003):
            S = BlockStruct.dense([Index(1, (n,))], dtype)
            x, y = synth.random_values(S, 1), synth.random_values(S, 2)
            tx, ty = itb.QTensor.from_host(ctx, S, x), itb.QTensor.from_host(ctx, S, y)
            assert abs(itb.norm(tx) - orc.nrm2(x)) <= 1e-13 * orc.nrm2(x)
            alpha = 0.75 if dtype == F else 0.75 - 0.5j
            itb.scale(tx, alpha)
            assert_close(tx.to_host(), alpha * x, 1e-15)
            out = itb.add(ty, -1.25, tx)
            assert_close(out.to_host(), y - 1.25 * alpha * x, 1e-15)
            itb.fill(ty, 3.5)
            assert np.all(ty.to_host() == 3.5)
    # dnrm2-style range safety
    S = BlockStruct.dense([Index(1, (1000,))], F)
    big = np.full(1000, 1e200)
    assert abs(itb.norm(itb.QTensor.from_host(ctx, S, big)) / orc.nrm2(big) - 1) < 1e-14
    tiny = np.full(1000, 1e-200)
    assert abs(itb.norm(itb.QTensor.from_host(ctx, S, tiny)) / orc.nrm2(tiny) - 1) < 1e-14


def test_dag_conj_and_complex_scale_promotes(ctx):
    S = synth.heff_chain([2, 3, 2], dtype=Z)[0]
    x = synth.random_values(S, 4)
    t = itb.dag(itb.QTensor.from_host(ctx, S, x))
    assert np.array_equal(t.to_host(), np.conj(x))
    Sr = synth.heff_chain([2, 3, 2], dtype=F)[0]
    xr = synth.random_values(Sr, 5)
    t2 = itb.scale(itb.QTensor.from_host(ctx, Sr, xr), 1j)
    assert t2.struct.is_complex
    assert_close(t2.to_host(), 1j * xr, 1e-15)
