"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when present, the reference harness
(oracle/_ref/libitref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs — never by the product (itensor_b200/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libitref.so")
REF_OMP_SO = os.path.join(_HERE, "_ref", "libitref_omp.so")  # same sources built with -DITENSOR_USE_OMP -fopenmp


class OrcDesc(C.Structure):
    _fields_ = [
        ("order", C.c_int32),
        ("dtype", C.c_int32),
        ("nsect", C.POINTER(C.c_int32)),
        ("sect", C.POINTER(C.c_int64)),
        ("nblocks", C.c_int64),
        ("blocks", C.POINTER(C.c_int32)),
        ("offsets", C.POINTER(C.c_int64)),
        ("nelems", C.c_int64),
    ]


class RefTensor(C.Structure):
    _fields_ = [
        ("order", C.c_int32),
        ("dtype", C.c_int32),
        ("nqn", C.c_int32),
        ("labels", C.POINTER(C.c_int64)),
        ("dirs", C.POINTER(C.c_int32)),
        ("nsect", C.POINTER(C.c_int32)),
        ("sect", C.POINTER(C.c_int64)),
        ("qn", C.POINTER(C.c_int32)),
        ("mods", C.POINTER(C.c_int32)),
        ("nblocks", C.c_int64),
        ("blocks", C.POINTER(C.c_int32)),
        ("offsets", C.POINTER(C.c_int64)),
        ("nelems", C.c_int64),
        ("data", C.POINTER(C.c_double)),
    ]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


_orc = None
_ref = None
_ref_omp = None


def build_oracle() -> None:
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def oracle():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.orc_nrm2.restype = C.c_double
        L.orc_nrm2.argtypes = [C.c_int64, C.POINTER(C.c_double)]
        L.orc_flux_blocks.restype = C.c_int64
        _orc = L
    return _orc


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def have_ref_omp() -> bool:
    return os.path.exists(REF_OMP_SO)


def host_cores() -> int:
    """cores this process may run on (cgroup/affinity aware), the thread count of CPU modes B and C"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def ref(omp: bool = False):
    global _ref, _ref_omp
    if omp:
        if _ref_omp is None:
            ref()  # preloads the BLAS siblings
            _ref_omp = _bind(C.CDLL(REF_OMP_SO))
        return _ref_omp
    if _ref is None:
        # the wheel-bundled OpenBLAS needs its sibling libquadmath/libgfortran: preload them
        import glob
        import sysconfig
        blas_dir = os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs")
        for pat in ("libquadmath-*.so*", "libgfortran-*.so*"):
            for f in sorted(glob.glob(os.path.join(blas_dir, pat))):
                C.CDLL(f, mode=C.RTLD_GLOBAL)
        _ref = _bind(C.CDLL(REF_SO))
    return _ref


def _bind(L):
    if True:
        for n in ("ref_contract", "ref_permute", "ref_pluseq"):
            getattr(L, n).restype = C.c_void_p
        L.ref_time_contract.restype = C.c_double
        L.ref_time_permute.restype = C.c_double
        L.ref_norm.restype = C.c_double
        L.ref_result_nblocks.restype = C.c_int64
        L.ref_result_nelems.restype = C.c_int64
        for n in ("ref_result_order", "ref_result_dtype", "ref_result_is_qn", "ref_result_nblocks", "ref_result_nelems",
                  "ref_result_nflux", "ref_result_labels", "ref_result_blocks", "ref_result_offsets", "ref_result_data",
                  "ref_result_flux", "ref_result_free"):
            getattr(L, n).argtypes = [C.c_void_p] + ([C.c_void_p] if n.split("_")[-1] in ("labels", "blocks", "offsets", "data", "flux") else [])
        L.ref_time_heff.restype = C.c_double
    return L


# The three CPU modes of SURVEY §8(d): (A) 1 BLAS thread, no OMP; (B) BLAS threads = cores; (C) the reference's own
# OpenMP loop over C blocks (ITENSOR_USE_OMP build, options.mk.sample:114-150) with cores threads and 1 BLAS thread.
# Threads are set through the libraries' own setters, so a launcher's OMP_NUM_THREADS=1 (torchrun) cannot change them.
def ref_mode(mode: str):
    """(library, description, threads) with the thread counts of CPU mode 'A' | 'B' | 'C' applied"""
    n = host_cores()
    if mode == "A":
        L = ref()
        L.ref_set_threads(1, 0)
        return L, "1 BLAS thread, no OpenMP", 1
    if mode == "B":
        L = ref()
        L.ref_set_threads(n, 0)
        return L, f"OpenBLAS threads={L.ref_get_blas_threads()}, no OpenMP", n
    if mode == "C":
        L = ref(omp=True)
        assert L.ref_set_threads(1, n) == 1
        return L, f"ITENSOR_USE_OMP=1, OMP threads={L.ref_get_omp_threads()}, OpenBLAS threads=1", n
    raise ValueError(mode)


# ---- oracle wrappers (operate on itensor_b200.tensor.BlockStruct-like objects: duck-typed) ----------
def _desc(s) -> OrcDesc:
    d = OrcDesc()
    d.order, d.dtype = s.order, s.dtype
    d.nsect, d.sect = _p(s.nsect, C.c_int32), _p(s.sect, C.c_int64)
    d.nblocks, d.blocks, d.offsets, d.nelems = s.nblocks, _p(s.blocks, C.c_int32), _p(s.offsets, C.c_int64), s.nelems
    return d


class OracleStruct:
    """minimal structure holder for oracle outputs"""

    def __init__(self, order, dtype, labels, nsect, sect, blocks, offsets, nelems):
        self.order, self.dtype, self.labels = order, dtype, labels
        self.nsect, self.sect, self.blocks, self.offsets, self.nelems = nsect, sect, blocks, offsets, nelems
        self.nblocks = blocks.shape[0]


def contract_structure(A, B):
    """(C structure, triples) from the oracle's restatement of getContractedOffsets."""
    L = oracle()
    la, lb = np.ascontiguousarray(A.labels, np.int32), np.ascontiguousarray(B.labels, np.int32)
    cap_p = max(1, A.nblocks * B.nblocks)
    rmax = A.order + B.order
    c_order = C.c_int32()
    c_labels = np.zeros(max(rmax, 1), np.int32)
    c_nsect = np.zeros(max(rmax, 1), np.int32)
    c_sect = np.zeros(max(len(A.sect) + len(B.sect), 1), np.int64)
    npairs, c_nblocks, c_nelems = C.c_int64(), C.c_int64(), C.c_int64()
    triples = np.zeros((cap_p, 3), np.int64)
    c_blocks = np.zeros((cap_p, max(rmax, 1)), np.int32)
    c_offsets = np.zeros(cap_p, np.int64)
    da, db = _desc(A), _desc(B)
    rc = L.orc_contract_structure(C.byref(da), _p(la, C.c_int32), C.byref(db), _p(lb, C.c_int32), C.byref(c_order),
                                  _p(c_labels, C.c_int32), _p(c_nsect, C.c_int32), _p(c_sect, C.c_int64), C.byref(npairs),
                                  _p(triples, C.c_int64), C.c_int64(cap_p), C.byref(c_nblocks), _p(c_blocks, C.c_int32),
                                  _p(c_offsets, C.c_int64), C.c_int64(cap_p), C.byref(c_nelems))
    assert rc == 0
    r, nb, np_ = c_order.value, c_nblocks.value, npairs.value
    ns = c_nsect[:r].copy()
    blocks = c_blocks.reshape(-1)[: nb * r].reshape(nb, r).copy()
    dtype = 1 if (A.dtype == 1 or B.dtype == 1) else 0
    S = OracleStruct(r, dtype, c_labels[:r].copy(), ns, c_sect[: int(ns.sum())].copy(), blocks, c_offsets[:nb].copy(), c_nelems.value)
    return S, triples[:np_].copy()


def contract_values(A, a_host, B, b_host, Cs, triples) -> np.ndarray:
    L = oracle()
    la, lb = np.ascontiguousarray(A.labels, np.int32), np.ascontiguousarray(B.labels, np.int32)
    a = np.ascontiguousarray(a_host).view(np.float64).reshape(-1)
    b = np.ascontiguousarray(b_host).view(np.float64).reshape(-1)
    out = np.zeros(max(Cs.nelems * (2 if Cs.dtype == 1 else 1), 1), np.float64)
    t = np.ascontiguousarray(triples, np.int64)
    da, db, dc = _desc(A), _desc(B), _desc(Cs)
    rc = L.orc_contract_values(C.byref(da), _p(la, C.c_int32), _p(a, C.c_double), C.byref(db), _p(lb, C.c_int32),
                               _p(b, C.c_double), C.byref(dc), _p(t, C.c_int64), C.c_int64(len(t)), _p(out, C.c_double))
    assert rc == 0
    out = out[: Cs.nelems * (2 if Cs.dtype == 1 else 1)]
    return out.view(np.complex128) if Cs.dtype == 1 else out


def contract(A, a_host, B, b_host):
    Cs, triples = contract_structure(A, B)
    return Cs, triples, contract_values(A, a_host, B, b_host, Cs, triples)


def permute(S, s_host, D, perm, alpha=1.0 + 0j, accumulate=False, d_host=None) -> np.ndarray:
    L = oracle()
    s = np.ascontiguousarray(s_host).view(np.float64).reshape(-1)
    nd = D.nelems * (2 if D.dtype == 1 else 1)
    out = np.zeros(max(nd, 1), np.float64)
    if accumulate:
        out[:nd] = np.ascontiguousarray(d_host).view(np.float64).reshape(-1)
    pm = np.ascontiguousarray(perm, np.int32)
    ds, dd = _desc(S), _desc(D)
    alpha = complex(alpha)
    rc = L.orc_permute(C.byref(ds), _p(s, C.c_double), C.byref(dd), _p(out, C.c_double), _p(pm, C.c_int32),
                       C.c_double(alpha.real), C.c_double(alpha.imag), C.c_int(1 if accumulate else 0))
    assert rc == 0, "oracle permute: source block without image"
    out = out[:nd]
    return out.view(np.complex128) if D.dtype == 1 else out


def nrm2(x: np.ndarray) -> float:
    v = np.ascontiguousarray(x).view(np.float64).reshape(-1)
    return float(oracle().orc_nrm2(C.c_int64(len(v)), _p(v, C.c_double)))


def flux_blocks(inds, flux) -> np.ndarray:
    r = len(inds)
    if r == 0:
        return np.zeros((1, 0), np.int32)
    nqn = len(inds[0].mods)
    nsect = np.array([i.nsect for i in inds], np.int32)
    qn = np.array([c for i in inds for s in i.qns for c in s], np.int32)
    mods = np.array(inds[0].mods, np.int32)
    dirs = np.array([i.dir for i in inds], np.int32)
    fl = np.array(list(flux), np.int32)
    args = [C.c_int32(r), _p(nsect, C.c_int32), _p(qn, C.c_int32), C.c_int32(nqn), _p(mods, C.c_int32), _p(dirs, C.c_int32), _p(fl, C.c_int32)]
    n = oracle().orc_flux_blocks(*args, None, C.c_int64(0))
    out = np.zeros((max(n, 1), r), np.int32)
    oracle().orc_flux_blocks(*args, _p(out, C.c_int32), C.c_int64(n))
    return out[:n]


# ---- reference harness wrappers -----------------------------------------------------------------------
class _Keep:
    """RefTensor + the numpy arrays it points into"""

    def __init__(self, S, host):
        inds = S.inds
        self.labels = np.array([i.label for i in inds], np.int64)
        self.dirs = np.array([i.dir for i in inds], np.int32)
        has_qn = len(inds) > 0 and inds[0].qns is not None
        self.nqn = len(inds[0].mods) if has_qn else 0
        self.qn = np.array([c for i in inds for s in i.qns for c in s], np.int32) if has_qn else np.zeros(1, np.int32)
        self.mods = np.array(inds[0].mods if has_qn else (1,), np.int32)
        self.data = np.ascontiguousarray(host).view(np.float64).reshape(-1).copy()
        t = RefTensor()
        t.order, t.dtype, t.nqn = S.order, S.dtype, self.nqn
        t.labels, t.dirs = _p(self.labels, C.c_int64), _p(self.dirs, C.c_int32)
        t.nsect, t.sect = _p(S.nsect, C.c_int32), _p(S.sect, C.c_int64)
        t.qn, t.mods = _p(self.qn, C.c_int32), _p(self.mods, C.c_int32)
        t.nblocks, t.blocks, t.offsets, t.nelems = S.nblocks, _p(S.blocks, C.c_int32), _p(S.offsets, C.c_int64), S.nelems
        t.data = _p(self.data, C.c_double)
        self.t, self.S = t, S


class RefResult:
    def __init__(self, h):
        L = ref()
        self.order = L.ref_result_order(h)
        self.dtype = L.ref_result_dtype(h)
        self.is_qn = L.ref_result_is_qn(h)
        self.nblocks = L.ref_result_nblocks(h)
        self.nelems = L.ref_result_nelems(h)
        self.labels = np.zeros(max(self.order, 1), np.int64)
        L.ref_result_labels(h, self.labels.ctypes.data)
        self.labels = self.labels[: self.order]
        self.blocks = np.zeros((max(self.nblocks, 1), max(self.order, 1)), np.int32)
        L.ref_result_blocks(h, self.blocks.ctypes.data)
        self.blocks = self.blocks.reshape(-1)[: self.nblocks * self.order].reshape(self.nblocks, self.order)
        self.offsets = np.zeros(max(self.nblocks, 1), np.int64)
        L.ref_result_offsets(h, self.offsets.ctypes.data)
        self.offsets = self.offsets[: self.nblocks]
        nreal = self.nelems * (2 if self.dtype == 1 else 1)
        d = np.zeros(max(nreal, 1), np.float64)
        L.ref_result_data(h, d.ctypes.data)
        d = d[:nreal]
        self.data = d.view(np.complex128) if self.dtype == 1 else d
        nf = L.ref_result_nflux(h)
        self.flux = np.zeros(max(nf, 1), np.int32)
        L.ref_result_flux(h, self.flux.ctypes.data)
        self.flux = self.flux[:nf]
        L.ref_result_free(h)


def ref_contract(A, a_host, B, b_host) -> RefResult:
    ka, kb = _Keep(A, a_host), _Keep(B, b_host)
    return RefResult(ref().ref_contract(C.byref(ka.t), C.byref(kb.t)))


def ref_time_contract(A, a_host, B, b_host, reps=3) -> float:
    ka, kb = _Keep(A, a_host), _Keep(B, b_host)
    return float(ref().ref_time_contract(C.byref(ka.t), C.byref(kb.t), C.c_int(reps)))


def ref_permute(S, s_host, new_inds) -> RefResult:
    k = _Keep(S, s_host)
    nl = np.array([i.label for i in new_inds], np.int64)
    return RefResult(ref().ref_permute(C.byref(k.t), _p(nl, C.c_int64)))


def ref_time_permute(S, s_host, new_inds, reps=3) -> float:
    k = _Keep(S, s_host)
    nl = np.array([i.label for i in new_inds], np.int64)
    return float(ref().ref_time_permute(C.byref(k.t), _p(nl, C.c_int64), C.c_int(reps)))


def ref_pluseq(A, a_host, B, b_host, alpha=1.0 + 0j) -> RefResult:
    ka, kb = _Keep(A, a_host), _Keep(B, b_host)
    alpha = complex(alpha)
    return RefResult(ref().ref_pluseq(C.byref(ka.t), C.byref(kb.t), C.c_double(alpha.real), C.c_double(alpha.imag)))


def ref_norm(S, s_host) -> float:
    k = _Keep(S, s_host)
    return float(ref().ref_norm(C.byref(k.t)))


def ref_dmrg_heisenberg(N, spin2, conserve_qns, maxdim, cutoff, niter, noise):
    n = len(maxdim)
    md = np.array(maxdim, np.int32)
    co = np.array(cutoff, np.float64)
    ni = np.array(niter, np.int32)
    no = np.array(noise, np.float64)
    e, s, ml = C.c_double(), C.c_double(), C.c_int()
    ref().ref_dmrg_heisenberg(C.c_int(N), C.c_int(spin2), C.c_int(1 if conserve_qns else 0), C.c_int(n), _p(md, C.c_int32),
                              _p(co, C.c_double), _p(ni, C.c_int32), _p(no, C.c_double), C.byref(e), C.byref(s), C.byref(ml))
    return e.value, s.value, ml.value


def ref_time_heff(structs, hosts, reps=1, want_result=False, mode=None):
    """seconds for one LocalOp::product chain phi*L*W1*W2*R through the reference (best of reps); mode: see ref_mode
    (None = mode B, all cores in OpenBLAS)"""
    L = ref_mode(mode or "B")[0]
    keeps = [_Keep(s, h) for s, h in zip(structs, hosts)]
    out = C.c_void_p()
    secs = L.ref_time_heff(*[C.byref(k.t) for k in keeps], C.c_int(reps), C.byref(out) if want_result else None)
    res = RefResult(out.value) if want_result else None
    return float(secs), res
