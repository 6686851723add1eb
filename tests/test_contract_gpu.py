"""GPU parity: block-sparse / dense contraction through the C ABI vs the CPU oracle.

Mirrors the reference's own test style (random inputs vs explicit loops): unittest/contract_test.cc:161-527,
unittest/itensor_test.cc:1029-1406 (dense / complex), :2734-2885 (block-deficient QN tensors, empty results).
"""
import numpy as np
import pytest

import itensor_b200 as itb
from itensor_b200 import synth
from itensor_b200._lib import lib
from itensor_b200.tensor import BlockStruct, Index
from oracle import orc
from util import assert_close, assert_struct_equal, gpu_contract_vs_oracle

pytestmark = pytest.mark.gpu
F, Z = itb.ITB_F64, itb.ITB_C64


@pytest.mark.parametrize("da,db", [(F, F), (Z, F), (F, Z), (Z, Z)])
def test_random_qn_pairs(ctx, da, db):
    rng = np.random.default_rng(100 + 2 * da + db)
    nonempty = 0
    for trial in range(60):
        ra, rb = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        nc = int(rng.integers(0, min(ra, rb) + 1))
        A, B = synth.random_qn_pair(rng, ra, rb, nc, dtype_a=da, dtype_b=db, max_size=6)
        plan = gpu_contract_vs_oracle(ctx, A, B, seed=trial)
        nonempty += plan.C.nelems > 0
    assert nonempty > 20


@pytest.mark.parametrize("da,db", [(F, F), (Z, Z), (F, Z), (Z, F)])
def test_random_qn_pairs_larger_sectors(ctx, da, db):
    """sector sizes large enough to reach the tile (DMMA) kernels and multi-chunk K loops"""
    rng = np.random.default_rng(7 + da + 3 * db)
    for trial in range(12):
        ra, rb = int(rng.integers(2, 5)), int(rng.integers(2, 5))
        nc = int(rng.integers(1, min(ra, rb) + 1))
        A, B = synth.random_qn_pair(rng, ra, rb, nc, dtype_a=da, dtype_b=db, max_size=23, max_sect=3, drop=0.1)
        gpu_contract_vs_oracle(ctx, A, B, seed=trial)


def _dense_pair(rng, dims_a, labs_a, dims_b, labs_b, da, db):
    mk = {}

    def ix(l, d):
        if l not in mk:
            mk[l] = Index(l, (d,))
        return mk[l]

    A = BlockStruct.dense([ix(l, d) for l, d in zip(labs_a, dims_a)], da)
    B = BlockStruct.dense([ix(l, d) for l, d in zip(labs_b, dims_b)], db)
    return A, B


@pytest.mark.parametrize("da,db", [(F, F), (Z, Z), (F, Z), (Z, F)])
def test_dense_pairs(ctx, da, db):
    """dense ITensor contraction incl. every transpose layout (contract_test.cc:30-158,161-462)"""
    rng = np.random.default_rng(5)
    cases = [
        ((37, 41), (1, 2), (41, 29), (2, 3)),          # A(m,k) B(k,n)
        ((41, 37), (2, 1), (41, 29), (2, 3)),          # A^T
        ((37, 41), (1, 2), (29, 41), (3, 2)),          # B^T
        ((41, 37), (2, 1), (29, 41), (3, 2)),          # both
        ((5, 7, 6), (1, 2, 3), (7, 4, 5), (2, 4, 1)),  # non-matrix reshapes
        ((4, 5, 6, 3), (1, 2, 3, 4), (6, 7, 4), (3, 5, 1)),
        ((150, 70), (1, 2), (70, 130), (2, 3)),        # several tiles / k-chunks
        ((9, 140, 8), (1, 2, 3), (8, 140, 11), (3, 2, 4)),
        ((3, 4, 5), (1, 2, 3), (3, 4, 5), (1, 2, 3)),  # full contraction -> rank 0
        ((6,), (1,), (7,), (2,)),                      # outer product
        ((300, 3), (1, 2), (3, 2), (2, 3)),            # skinny (streaming kernel)
        ((2, 3), (1, 2), (3, 500), (2, 3)),            # skinny, long side n
        ((1, 5000), (1, 2), (5000, 1), (2, 3)),        # long dot
    ]
    for i, (dimsa, la, dimsb, lb) in enumerate(cases):
        A, B = _dense_pair(rng, dimsa, la, dimsb, lb, da, db)
        gpu_contract_vs_oracle(ctx, A, B, seed=i)


def test_contract_test_2x2_exact(ctx):
    """the reference's exact integer 2x2 cases (unittest/contract_test.cc:30-158): REQUIRE =="""
    vals_a = np.array([11.0, 21.0, 12.0, 22.0])  # column-major A(i,j): A11,A21,A12,A22
    vals_b = np.array([110.0, 210.0, 120.0, 220.0])
    i, j, k = Index(1, (2,)), Index(2, (2,)), Index(3, (2,))
    for a_inds, b_inds in [((i, j), (j, k)), ((j, i), (j, k)), ((i, j), (k, j)), ((j, i), (k, j))]:
        A, B = BlockStruct.dense(a_inds), BlockStruct.dense(b_inds)
        tA, tB = itb.QTensor.from_host(ctx, A, vals_a), itb.QTensor.from_host(ctx, B, vals_b)
        got = itb.contract(tA, tB).to_host()
        a = vals_a.reshape(2, 2, order="F")
        b = vals_b.reshape(2, 2, order="F")
        am = a if a_inds[0] is i else a.T
        bm = b if b_inds[0] is j else b.T
        want = (am @ bm).reshape(-1, order="F")
        assert np.array_equal(got, want)  # small integers: exact in fp64 whatever the summation order


def test_no_output_blocks(ctx):
    """contraction with no matching block pair -> empty storage (itensor_test.cc:2849-2865)"""
    l = Index(1, (2, 3), ((0,), (1,)), 1)
    s = Index(2, (1, 1), ((0,), (1,)), 1)
    A = BlockStruct([l, s], np.array([[0, 0]], np.int32))
    B = BlockStruct([l.dag(), Index(3, (2,), ((0,),), 1)], np.array([[1, 0]], np.int32))
    plan = gpu_contract_vs_oracle(ctx, A, B)
    assert plan.C.nblocks == 0 and plan.C.nelems == 0 and plan.npairs == 0


@pytest.mark.parametrize("dtype", [F, Z])
@pytest.mark.parametrize("order", ["l s1 s2 r", "s1 l r s2"])
def test_heff_product_chain(ctx, dtype, order):
    """LocalOp::product = 4 chained contractions (mps/localop.h:324-365), checked step by step"""
    structs = synth.heff_chain([3, 11, 17, 9, 2], [2, 8, 19, 10, 4], dtype=dtype, phi_order=order)
    hosts = [synth.random_values(s, 10 + i) for i, s in enumerate(structs)]
    t = itb.QTensor.from_host(ctx, structs[0], hosts[0])
    ref_s, ref_v = structs[0], hosts[0]
    for s, h in zip(structs[1:], hosts[1:]):
        plan = itb.ContractPlan(t.struct, s)
        t = itb.contract(t, itb.QTensor.from_host(ctx, s, h), plan)
        cs, tr, ref_v = orc.contract(ref_s, ref_v, s, h)
        assert_struct_equal(plan.C, cs)
        assert_close(t.to_host(), ref_v, what="heff step")
        ref_s = plan.C
    assert [i.label for i in t.inds] == [i.prime().label for i in
                                         (structs[0].inds[k] for k in _order_lsr(structs[0]))]


@pytest.mark.parametrize("mix", ["complex_state_real_ops", "complex_state_complex_ops", "real_state_complex_ops"])
@pytest.mark.parametrize("sizes", [[3, 11, 17, 9, 2], [40, 130, 90]])
def test_heff_chain_mixed_dtypes_rowgroups(ctx, sizes, mix):
    """The three pairings with a complex tensor in the chain: complex phi with real L / W1 / W2 / R (a real Hamiltonian
    applied to a complex state: the complex operand is streamed as a real one of doubled leading extent), complex phi with
    complex operators (input / output slots = (re,im) components, signed weights) and real phi with complex operators
    (strided C rows). The MPO steps run on the row-group streaming kernel. Every plan is executed three times: the planner is
    tiered (C-stationary kernels first, row groups from the third execution of a small streaming class), all three results
    must match the oracle."""
    sc, sr = synth.heff_chain(sizes, dtype=Z), synth.heff_chain(sizes, dtype=F)
    structs = {"complex_state_real_ops": (sc[0],) + tuple(sr[1:]), "complex_state_complex_ops": tuple(sc),
               "real_state_complex_ops": (sr[0],) + tuple(sc[1:])}[mix]
    hosts = [synth.random_values(s, 30 + i) for i, s in enumerate(structs)]
    t = itb.QTensor.from_host(ctx, structs[0], hosts[0])
    ref_s, ref_v = structs[0], hosts[0]
    saw_rowgroups = False
    for s, h in zip(structs[1:], hosts[1:]):
        plan = itb.ContractPlan(t.struct, s)
        cs, tr, ref_v = orc.contract(ref_s, ref_v, s, h)
        assert_struct_equal(plan.C, cs)
        tb = itb.QTensor.from_host(ctx, s, h)
        for _ in range(3):
            out = itb.contract(t, tb, plan)
            assert_close(out.to_host(), ref_v, what="heff step (complex state, real operators)")
        saw_rowgroups |= lib().itb_contract_plan_rowgroups(plan._h, None, 0) > 0
        t, ref_s = out, plan.C
    assert saw_rowgroups


def _order_lsr(phi):
    # result order is (l', s1', s2', r') in the order the primed partners were appended:
    # L brings l', W1 brings s1', W2 brings s2', R brings r'
    ids = [i.id for i in phi.inds]
    return [ids.index(1), ids.index(3), ids.index(4), ids.index(2)]


def test_scalar_product(ctx):
    """<V|q> : full contraction to rank 0 with many block pairs (iterativesolvers.h:169,315)"""
    for dtype in (F, Z):
        phi = synth.heff_chain([5, 40, 77, 31, 6], dtype=dtype)[0]
        v = synth.random_values(phi, 1)
        q = synth.random_values(phi, 2)
        tq = itb.QTensor.from_host(ctx, phi, q)
        tv = itb.dag(itb.QTensor.from_host(ctx, phi, v))
        out = itb.contract(tv, tq)
        assert out.struct.order == 0 and out.struct.nelems == 1
        want = np.vdot(v, q)
        got = itb.elt(out)
        assert abs(got - want) <= 1e-12 * abs(np.abs(v) @ np.abs(q))


def test_large_blocks_linearity(ctx):
    """full-size property check where the O(MNK) oracle is too slow: (A1+2*A2)*B == A1*B + 2*A2*B"""
    structs = synth.heff_chain(synth.gaussian_sectors(600, 7))
    phi, L = structs[0], structs[1]
    a1, a2, b = synth.random_values(phi, 1), synth.random_values(phi, 2), synth.random_values(L, 3)
    plan = itb.ContractPlan(phi, L)
    tb = itb.QTensor.from_host(ctx, L, b)
    c1 = itb.contract(itb.QTensor.from_host(ctx, phi, a1), tb, plan).to_host()
    c2 = itb.contract(itb.QTensor.from_host(ctx, phi, a2), tb, plan).to_host()
    c12 = itb.contract(itb.QTensor.from_host(ctx, phi, a1 + 2 * a2), tb, plan).to_host()
    assert_close(c12, c1 + 2 * c2, 1e-12, "linearity")
    # and a spot check of one C block against NumPy on that block pair list
    pairs = plan.pairs()
    ic = int(pairs[len(pairs) // 2, 2])
    want = np.zeros(plan.C.block_shape(ic))
    for ia, ib, jc in pairs:
        if jc != ic:
            continue
        ab = a1[phi.offsets[ia]: phi.offsets[ia] + int(np.prod(phi.block_shape(ia)))].reshape(phi.block_shape(ia), order="F")
        bb = b[L.offsets[ib]: L.offsets[ib] + int(np.prod(L.block_shape(ib)))].reshape(L.block_shape(ib), order="F")
        # phi(l,s1,s2,r) . L(l+,k,l') over l
        want += np.tensordot(ab, bb, axes=([0], [0]))
    got = c1[plan.C.offsets[ic]: plan.C.offsets[ic] + want.size].reshape(want.shape, order="F")
    assert_close(got, want, 1e-12, "block spot check")


def test_bench_workload_vs_reference(ctx):
    """The EXACT bench.py workload (H_eff*phi at maxdim 2000: 148-CTA stream-K cuts, split-K reductions, row groups) against
    the unmodified reference build (oracle/_ref/libitref.so travels to the GPU box): block list, offsets and index order
    bit-exact, values <= 1e-12 relative, for each of the four contractions (every step is fed with the reference's own
    previous intermediate so that errors cannot compound or cancel)."""
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libitref.so not built")
    sizes = synth.gaussian_sectors(2000, 9)
    structs = synth.heff_chain(sizes)
    hosts = [synth.random_values(s, 10 + i) for i, s in enumerate(structs)]
    cur_s, cur_h = structs[0], hosts[0]
    for k in range(4):
        rr = orc.ref_contract(cur_s, cur_h, structs[k + 1], hosts[k + 1])
        plan = itb.ContractPlan(cur_s, structs[k + 1])
        assert np.array_equal(rr.blocks, plan.C.blocks) and np.array_equal(rr.offsets, plan.C.offsets)
        assert rr.nelems == plan.C.nelems and list(rr.labels) == [int(x) for x in plan.C.labels]
        got = itb.contract(itb.QTensor.from_host(ctx, cur_s, cur_h), itb.QTensor.from_host(ctx, structs[k + 1], hosts[k + 1]), plan).to_host()
        assert_close(got, rr.data, 1e-12, f"bench workload step {k + 1}")
        cur_s, cur_h = plan.C, rr.data
    assert cur_s.order == 4 and cur_s.nelems == structs[0].nelems


def test_refined_partition_vs_reference(ctx):
    """itb_contract_plan_refine re-cuts the stream-K partition from measured per-CTA cycles: the refined plans of the bench
    workload (steps 1 and 4, the two tile-kernel launches) must still reproduce the reference to 1e-12, the refinement run
    itself must leave the correct result in C, and re-running a refined plan must be bitwise reproducible."""
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libitref.so not built")
    sizes = synth.gaussian_sectors(2000, 9)
    structs = synth.heff_chain(sizes)
    hosts = [synth.random_values(s, 10 + i) for i, s in enumerate(structs)]
    cur_s, cur_h = structs[0], hosts[0]
    for k in range(4):
        rr = orc.ref_contract(cur_s, cur_h, structs[k + 1], hosts[k + 1])
        plan = itb.ContractPlan(cur_s, structs[k + 1])
        ta, tb = itb.QTensor.from_host(ctx, cur_s, cur_h), itb.QTensor.from_host(ctx, structs[k + 1], hosts[k + 1])
        tc = itb.QTensor(ctx, plan.C, ctx.empty(plan.C.nreal))
        gain = plan.refine(ctx, ta.ptr, tb.ptr, tc.ptr, rounds=3)
        assert gain >= 1.0 - 1e-12  # the best measured partition is kept, the first one included
        assert_close(tc.to_host(), rr.data, 1e-12, f"result left by the refinement, step {k + 1}")
        got1 = itb.contract(ta, tb, plan).to_host()
        got2 = itb.contract(ta, tb, plan).to_host()
        assert_close(got1, rr.data, 1e-12, f"refined plan, step {k + 1}")
        assert np.array_equal(got1, got2)
        cur_s, cur_h = plan.C, rr.data


def test_bench_workload_complex_vs_reference(ctx):
    """same chain with complex tensors (folded real-GEMM decomposition) at maxdim 600 against the reference"""
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libitref.so not built")
    structs = synth.heff_chain(synth.gaussian_sectors(600, 7), dtype=Z)
    hosts = [synth.random_values(s, 20 + i) for i, s in enumerate(structs)]
    cur_s, cur_h = structs[0], hosts[0]
    for k in range(4):
        rr = orc.ref_contract(cur_s, cur_h, structs[k + 1], hosts[k + 1])
        plan = itb.ContractPlan(cur_s, structs[k + 1])
        assert np.array_equal(rr.blocks, plan.C.blocks) and np.array_equal(rr.offsets, plan.C.offsets)
        got = itb.contract(itb.QTensor.from_host(ctx, cur_s, cur_h), itb.QTensor.from_host(ctx, structs[k + 1], hosts[k + 1]), plan).to_host()
        assert_close(got, rr.data, 1e-12, f"complex chain step {k + 1}")
        cur_s, cur_h = plan.C, rr.data
