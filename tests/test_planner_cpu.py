"""Host integer planning (libitb200.so, no device) vs the CPU oracle: block structure, offsets, index
order and pair enumeration must be bit-exact. Also pins the reference's own integer KATs."""
import numpy as np
import pytest

import itensor_b200 as itb
from itensor_b200 import synth
from itensor_b200.tensor import BlockStruct, Index, flux_blocks, permuted_struct
from oracle import orc
from util import assert_struct_equal

F, Z = itb.ITB_F64, itb.ITB_C64


def test_contract_structure_matches_oracle_random():
    rng = np.random.default_rng(0)
    nonempty = 0
    for trial in range(400):
        ra, rb = int(rng.integers(0, 6)), int(rng.integers(0, 6))
        nc = int(rng.integers(0, min(ra, rb) + 1))
        A, B = synth.random_qn_pair(rng, ra, rb, nc, dtype_a=int(rng.integers(0, 2)), dtype_b=int(rng.integers(0, 2)))
        P = itb.ContractPlan(A, B)
        Cs, tr = orc.contract_structure(A, B)
        assert_struct_equal(P.C, Cs)
        assert np.array_equal(P.pairs(), tr)
        nonempty += Cs.nelems > 0
    assert nonempty > 100


def test_flops_count():
    """flops = sum_pairs 2*M*N*K, x2 real*cplx, x4 cplx*cplx"""
    structs = synth.heff_chain([3, 9, 14, 8, 2])
    phi, L = structs[0], structs[1]
    P = itb.ContractPlan(phi, L)
    tot = 0
    for ia, ib, ic in P.pairs():
        sa, sb = phi.block_shape(ia), L.block_shape(ib)
        tot += 2 * np.prod(sa) * np.prod(sb) / sa[0]  # contracted index is phi's first (l), extent sa[0]
    assert abs(P.flops - tot) < 1e-6 * tot
    for da, db, mult in [(Z, F, 2), (F, Z, 2), (Z, Z, 4)]:
        a = BlockStruct(phi.inds, phi.blocks, da)
        b = BlockStruct(L.inds, L.blocks, db)
        assert abs(itb.ContractPlan(a, b).flops - mult * tot) < 1e-6 * tot


def test_heff_census_matches_survey():
    """S=1/2 Heisenberg centre bond at m=400 (SURVEY 8a census: sectors 12,83,158,113,29,5):
    60/102/106/60 pairs and 60/62/62/22 C blocks for the four LocalOp::product steps."""
    structs = synth.heff_chain([12, 83, 158, 113, 29, 5])
    s = structs[0]
    pairs, cblocks = [], []
    for t in structs[1:]:
        P = itb.ContractPlan(s, t)
        pairs.append(P.npairs)
        cblocks.append(P.C.nblocks)
        s = P.C
    assert pairs == [60, 102, 106, 60]
    assert cblocks == [60, 62, 62, 22]


def test_block_deficient_kat():
    """unittest/itensor_test.cc:2734-2765: A = a*prime(dag(a),i) has 4 blocks / 64 elements"""
    i = Index(1, (2, 3, 4, 5, 6), ((0,), (1,), (2,), (1,), (3,)), 1)
    l = Index(2, (3,), ((0,),), 1)
    a_blocks = flux_blocks([i, l], (1,))
    assert a_blocks.tolist() == [[1, 0], [3, 0]]
    a = BlockStruct([i, l], a_blocks)
    adag = BlockStruct([i.dag().prime(), l.dag()], a_blocks)
    P = itb.ContractPlan(a, adag)
    assert P.C.nblocks == 4 and P.C.nelems == 64
    assert P.C.blocks.tolist() == [[1, 1], [3, 1], [1, 3], [3, 3]]  # last index most significant
    assert P.C.offsets.tolist() == [0, 9, 24, 39]


def test_flux_blocks_matches_oracle_incl_modular_qns():
    rng = np.random.default_rng(3)
    for trial in range(200):
        r = int(rng.integers(1, 5))
        nqn = int(rng.integers(1, 3))
        mods = tuple(int(rng.choice([1, 2, 3, -2])) for _ in range(nqn))
        inds = []
        for j in range(r):
            ns = int(rng.integers(1, 4))
            qns = tuple(tuple(int(rng.integers(-2, 3)) for _ in range(nqn)) for _ in range(ns))
            inds.append(Index(10 + j, tuple(int(v) for v in rng.integers(1, 4, ns)), qns, int(rng.choice([-1, 1])), mods))
        flux = tuple(int(rng.integers(-2, 3)) for _ in range(nqn))
        got, want = flux_blocks(inds, flux), orc.flux_blocks(inds, flux)
        assert np.array_equal(got, want)
        # sorted in the reference Block order (reverse lexicographic)
        if len(got) > 1:
            keys = [tuple(reversed(b)) for b in got.tolist()]
            assert keys == sorted(keys)


def test_permute_struct_fill_in_rule():
    """QN permute allocates every flux-allowed block; dense/no-QN keeps the block list"""
    rng = np.random.default_rng(5)
    A, _ = synth.random_qn_pair(rng, 3, 1, 0, drop=0.6, max_sect=3)
    new = [A.inds[2], A.inds[0], A.inds[1]]
    D, perm = permuted_struct(A, new, flux=(0,))
    assert perm == [1, 2, 0]
    assert np.array_equal(D.blocks, orc.flux_blocks(new, (0,)))
    assert D.nblocks >= A.nblocks
    D2, _ = permuted_struct(A, new)  # no fill-in requested
    assert D2.nblocks == A.nblocks


def test_cblock_range_restriction():
    """output-block sharding unit for multi-GPU: ranges partition the work, structure is unchanged"""
    structs = synth.heff_chain([3, 9, 14, 8, 2])
    P = itb.ContractPlan(structs[0], structs[1])
    full = [P.info.n_gemm_tiles, P.info.n_skinny, P.info.n_dot]
    fl_full = sum(P.info.class_flops)
    nb = P.C.nblocks
    parts, fl = [0, 0, 0], 0.0
    for lo, hi in [(0, nb // 3), (nb // 3, 2 * nb // 3), (2 * nb // 3, nb)]:
        P.set_cblock_range(lo, hi)
        parts = [a + b for a, b in zip(parts, [P.info.n_gemm_tiles, P.info.n_skinny, P.info.n_dot])]
        fl += sum(P.info.class_flops)
    assert abs(fl - fl_full) < 1e-9 * fl_full
    # (whether skinny C blocks stream or ride the tile queue is decided per executed range, so only the split-K
    # dot items and the flops are additive)
    assert parts[2] == full[2]
    assert P.C.nblocks == nb


def _introspect(P):
    import ctypes as C

    from itensor_b200._lib import lib

    n = lib().itb_contract_plan_tiles(P._h, None, 0)
    t = np.zeros((n, 8), np.int32)
    lib().itb_contract_plan_tiles(P._h, t.ctypes.data_as(C.POINTER(C.c_int32)), n)
    nc = lib().itb_contract_plan_cblks(P._h, None, 0)
    c = np.zeros((nc, 4), np.int64)
    lib().itb_contract_plan_cblks(P._h, c.ctypes.data_as(C.POINTER(C.c_int64)), nc)
    ng = lib().itb_contract_plan_cta_begin(P._h, None, 0)
    g = np.zeros(ng, np.int32)
    lib().itb_contract_plan_cta_begin(P._h, g.ctypes.data_as(C.POINTER(C.c_int32)), ng)
    nr = lib().itb_contract_plan_rowgroups(P._h, None, 0)
    r = np.zeros((nr, 4), np.int64)
    lib().itb_contract_plan_rowgroups(P._h, r.ctypes.data_as(C.POINTER(C.c_int64)), nr)
    return t, c, g, r


@pytest.mark.parametrize("sizes,dtype", [([6, 47, 214, 507, 634, 418, 145, 27, 3], F), ([12, 83, 158, 113, 29, 5], F),
                                         ([3, 9, 14, 8, 2], Z), ([40, 70, 33], F)])
def test_stream_k_partition_covers_every_tile_once(sizes, dtype):
    """The DMMA tile schedule: every (C block, tile) of the tile class appears with K-chunk ranges that partition its
    chunk loop exactly (pieces of a cut tile are consecutive items with consecutive workspace slots), the static per-CTA
    ranges plus the shared dynamic queue partition the item list, and the modelled work is balanced."""
    structs = synth.heff_chain(sizes, dtype=dtype)
    s = structs[0]
    for tt in structs[1:]:
        P = itb.ContractPlan(s, tt)
        s = P.C
        t, c, g, r = _introspect(P)
        if len(t) == 0:
            continue
        # hybrid schedule: g[b]..g[b+1] is CTA b's static range (b < 148), g[148]..g[149] the shared dynamic queue
        assert g[0] == 0 and g[-1] == len(t) and np.all(np.diff(g) >= 0) and len(g) == 150
        by_tile = {}
        for i, (cb, m0, n0, tm, tn, c0, c1, slot) in enumerate(t.tolist()):
            assert 0 <= m0 < c[cb, 0] and 0 <= n0 < c[cb, 1] and m0 % tm == 0 and n0 % tn == 0 and c1 > c0 >= 0
            by_tile.setdefault((cb, m0, n0, tm, tn), []).append((i, c0, c1, slot))
        covered = {}
        for (cb, m0, n0, tm, tn), pieces in by_tile.items():
            pieces.sort(key=lambda p: p[1])
            assert pieces[0][1] == 0
            for a, b in zip(pieces, pieces[1:]):
                assert a[2] == b[1] and b[0] == a[0] + 1 and b[3] == a[3] + 1  # consecutive chunks, items and slots
            if len(pieces) == 1:
                assert pieces[0][3] == -1  # an uncut tile writes C directly
            else:
                assert all(p[3] >= 0 for p in pieces)
            covered.setdefault(cb, {"nch": pieces[-1][2], "area": 0})
            assert covered[cb]["nch"] == pieces[-1][2]  # every tile of a C block walks the same K loop
            covered[cb]["area"] += min(tm, c[cb, 0] - m0) * min(tn, c[cb, 1] - n0)
        for cb, v in covered.items():
            assert v["area"] == c[cb, 0] * c[cb, 1]  # tiles tile the block exactly
        if P.flops > 1e9:  # enough work to balance: static ranges carry ~85 % of the work, none more than 1.5x their mean
            w = np.array([(c1 - c0) * tm * tn for _, _, _, tm, tn, c0, c1, _ in t.tolist()], float)
            load = np.array([w[g[b]:g[b + 1]].sum() for b in range(148)])
            tail = w[g[148]:g[149]]
            if g[148] > 0 and len(tail):  # hybrid schedule (ITB_STATIC_FRAC > 0): the static ranges carry most of the work
                assert load.max() <= 1.5 * load.mean() and 0.05 <= tail.sum() / w.sum() <= 0.30
            if len(tail):  # the queue ends in small pieces: that is what lets the CTAs finish together
                assert tail[-148:].max() <= 0.35 * w.sum() / 148
            else:          # ITB_SCHED=streamk: pure static partition, balanced by the cycle model
                assert load.max() <= 1.5 * load.mean()


def test_row_groups_read_every_input_once():
    """Streaming class (MPO steps of H_eff*phi): the row groups' input slots x rows account for every element of A
    exactly once and their output slots x rows for every element of C exactly once."""
    structs = synth.heff_chain([6, 47, 214, 507, 634, 418, 145, 27, 3])
    s = structs[0]
    seen = 0
    for k, tt in enumerate(structs[1:]):
        P = itb.ContractPlan(s, tt)
        t, c, g, r = _introspect(P)
        if k in (1, 2):  # *W1, *W2
            assert len(r) > 0 and len(t) == 0
            assert int((r[:, 0] * r[:, 3]).sum()) == s.nelems
            assert int((r[:, 1] * r[:, 3]).sum()) == P.C.nelems
            assert r[:, 0].max() <= 32 and r[:, 1].max() <= 16 and r[:, 2].max() <= 3
            seen += 1
        s = P.C
    assert seen == 2


def test_generic_key_path_matches_oracle_too():
    """Block tuples that do not fit a 64-bit key take the vector-key path of the planner; ITB_PLAN_GENERIC forces it so
    that the same oracle comparison covers it (subprocess: the switch is read once per process). The same subprocess runs
    the guided-queue schedule (ITB_SCHED=guided) through the partition checks; the default is the static stream-K one."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", os.path.join(root, "tests", "test_planner_cpu.py"), "-k",
                          "structure_matches_oracle_random or heff_census or stream_k"], env=dict(os.environ, ITB_PLAN_GENERIC="1", ITB_SCHED="guided"),
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-2000:]
