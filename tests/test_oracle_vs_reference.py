"""Pins the CPU oracle against the UNMODIFIED reference built into oracle/_ref/libitref.so
(only available where /root/reference exists at build time; skipped elsewhere — the committed golden
vectors in tests/golden/ carry the same pin to boxes without the reference)."""
import numpy as np
import pytest

import itensor_b200 as itb
from itensor_b200 import synth
from itensor_b200.tensor import permuted_struct
from oracle import orc
from util import assert_close

pytestmark = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref/libitref.so not built")
F, Z = itb.ITB_F64, itb.ITB_C64


def test_contract_random_pairs():
    rng = np.random.default_rng(42)
    checked = 0
    for trial in range(250):
        ra, rb = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        nc = int(rng.integers(0, min(ra, rb) + 1))
        A, B = synth.random_qn_pair(rng, ra, rb, nc, dtype_a=int(rng.integers(0, 2)), dtype_b=int(rng.integers(0, 2)))
        if A.nblocks == 0 or B.nblocks == 0:
            continue
        a, b = synth.random_values(A, trial), synth.random_values(B, 500 + trial)
        Cs, tr, val = orc.contract(A, a, B, b)
        R = orc.ref_contract(A, a, B, b)
        if Cs.order > 0:
            assert np.array_equal(R.labels, Cs.labels.astype(np.int64))
            assert np.array_equal(R.blocks, Cs.blocks) and np.array_equal(R.offsets, Cs.offsets)
        if Cs.nelems == 0:
            assert R.nelems == 0  # "no output blocks => empty storage"
            continue
        assert_close(val, R.data, 1e-12, f"trial {trial}")
        checked += 1
    assert checked > 80


def test_heff_chain_vs_reference_product():
    """the whole LocalOp::product chain through the reference's operator*= vs the oracle, step by step"""
    for dt in (F, Z):
        structs = synth.heff_chain([2, 6, 9, 5, 1], [1, 5, 8, 6, 2], dtype=dt)
        hosts = [synth.random_values(s, 10 + i) for i, s in enumerate(structs)]
        _, R = orc.ref_time_heff(structs, hosts, 1, want_result=True)
        s, v = structs[0], hosts[0]
        for t, h in zip(structs[1:], hosts[1:]):
            Cs, tr, v = orc.contract(s, v, t, h)
            s = itb.ContractPlan(s, t).C
        assert np.array_equal(R.blocks, Cs.blocks) and np.array_equal(R.offsets, Cs.offsets)
        assert_close(v, R.data, 1e-12, "heff")


def test_permute_fill_in_and_values():
    rng = np.random.default_rng(7)
    for trial in range(60):
        r = int(rng.integers(2, 6))
        S, _ = synth.random_qn_pair(rng, r, 1, 0, max_sect=3, max_size=5, dtype_a=int(rng.integers(0, 2)), drop=0.3)
        if S.nblocks == 0:
            continue
        new_inds = [S.inds[i] for i in rng.permutation(r)]
        if all(a.same(b) for a, b in zip(new_inds, S.inds)):
            continue
        s = synth.random_values(S, trial)
        R = orc.ref_permute(S, s, new_inds)
        D, perm = permuted_struct(S, new_inds, flux=(0,))
        assert np.array_equal(R.blocks, D.blocks) and np.array_equal(R.offsets, D.offsets)  # F7 fill-in rule
        assert np.array_equal(orc.permute(S, s, D, perm), R.data)
        assert abs(orc.nrm2(s) - orc.ref_norm(S, s)) <= 1e-14 * orc.nrm2(s)


def test_dmrg_energy_kat():
    """sample/dmrg.cc logic (Neel start, AutoMPO Heisenberg, Sz QNs, schedule 10,20,100,100,200 / cutoff 1e-10 /
    niter 2 / noise 1e-7,1e-8,0) on a short S=1/2 chain: the QN run is reproducible to 1e-12, so the
    value this container's reference build produced is a usable pin for later GPU-storage DMRG runs."""
    e, secs, maxlink = orc.ref_dmrg_heisenberg(20, 1, True, [10, 20, 100, 100, 200], [1e-10] * 5, [2] * 5, [1e-7, 1e-8, 0, 0, 0])
    assert abs(e - (-8.682473331005815)) < 1e-10
    assert maxlink == 39
