"""Host-side mirror (Python) of the slice of the ITensor interface that sits on the hot path.

Index / QN sectors / block lists live here on the host, exactly like IndexSet + BlockOffsets do in the
reference (itensor/index.h:78-277, itensor/itdata/qdense.h:34-187); element data lives in HBM as a flat
torch tensor and is only ever touched through the C ABI (include/itb200.h). Names follow the reference:

    contract(A, B)            ITensor::operator*=      itensor/itensor.cc:935-981
    permute(T, inds)          ITensor::permute         itensor/itensor.cc:629-671   (QN: fills in all
                                                       flux-allowed blocks, itdata/qdense.cc:862)
    add(A, alpha, B)          daxpy / operator+=       itensor/itensor.cc:1207-1264 (permuting, merges blocks)
    norm / scale / fill / dag / elt                    itensor/itensor.cc:753-766,1017-1035,331-364,246

This is the test/bench harness side of the boundary; the C++ storage-type plugin
(itensor_b200/plugin/) is the drop-in for the reference itself.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import ITB_C64, ITB_F64, ContractInfo, TensorDesc, check, lib


def _i32p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _i64p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


@dataclass(frozen=True)
class Index:
    """An index with QN sectors. Identity = (id, plev); dir = +1 (Out) / -1 (In)."""

    id: int
    sizes: tuple  # sector sizes (one entry for a non-QN index)
    qns: Optional[tuple] = None  # per sector: tuple of QN components; None => no QNs
    dir: int = 1
    mods: tuple = (1,)
    plev: int = 0

    @property
    def dim(self) -> int:
        return int(sum(self.sizes))

    @property
    def nsect(self) -> int:
        return len(self.sizes)

    @property
    def label(self) -> int:
        return self.id * 8 + self.plev

    def dag(self) -> "Index":
        return Index(self.id, self.sizes, self.qns, -self.dir, self.mods, self.plev)

    def prime(self, n: int = 1) -> "Index":
        return Index(self.id, self.sizes, self.qns, self.dir, self.mods, self.plev + n)

    def same(self, other: "Index") -> bool:
        return self.id == other.id and self.plev == other.plev


def flux_blocks(inds: Sequence[Index], flux: Sequence[int]) -> np.ndarray:
    """All blocks allowed by `flux`, reference-sorted (getBlockOffsets, itdata/qdense.cc:133-173)."""
    r = len(inds)
    if r == 0:
        return np.zeros((1, 0), np.int32)
    nqn = len(inds[0].mods)
    nsect = np.array([i.nsect for i in inds], np.int32)
    qn = np.array([c for i in inds for s in i.qns for c in s], np.int32)
    mods = np.array(inds[0].mods, np.int32)
    dirs = np.array([i.dir for i in inds], np.int32)
    fl = np.array(list(flux), np.int32)
    n = lib().itb_flux_blocks(r, _i32p(nsect), _i32p(qn), nqn, _i32p(mods), _i32p(dirs), _i32p(fl), None, 0)
    if n < 0:
        check(int(n))
    out = np.zeros((max(n, 1), r), np.int32)
    lib().itb_flux_blocks(r, _i32p(nsect), _i32p(qn), nqn, _i32p(mods), _i32p(dirs), _i32p(fl), _i32p(out), n)
    return out[:n]


class BlockStruct:
    """Integer structure of a tensor: indices + sorted block list + element offsets."""

    def __init__(self, inds: Sequence[Index], blocks: np.ndarray, dtype: int = ITB_F64, offsets: Optional[np.ndarray] = None):
        self.inds = list(inds)
        self.order = len(self.inds)
        self.dtype = dtype
        blocks = np.asarray(blocks, np.int32)
        if blocks.ndim != 2:
            blocks = blocks.reshape(-1, self.order) if self.order else np.zeros((1, 0), np.int32)
        self.blocks = np.ascontiguousarray(blocks)
        self.nsect = np.array([i.nsect for i in self.inds], np.int32)
        self.sect = np.array([s for i in self.inds for s in i.sizes], np.int64)
        self.nblocks = self.blocks.shape[0]
        sizes = self.block_sizes()
        if offsets is None:
            offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]) if self.nblocks else np.zeros(0)
        self.offsets = np.ascontiguousarray(np.asarray(offsets, np.int64))
        self.nelems = int(sizes.sum()) if self.nblocks else 0

    @staticmethod
    def dense(inds: Sequence[Index], dtype: int = ITB_F64) -> "BlockStruct":
        return BlockStruct(inds, np.zeros((1, len(inds)), np.int32), dtype)

    def block_sizes(self) -> np.ndarray:
        sz = np.ones(self.nblocks, np.int64)
        for j, ix in enumerate(self.inds):
            sz *= np.asarray(ix.sizes, np.int64)[self.blocks[:, j]]
        return sz

    def block_shape(self, b: int) -> tuple:
        return tuple(int(self.inds[j].sizes[self.blocks[b, j]]) for j in range(self.order))

    @property
    def labels(self) -> np.ndarray:
        return np.array([i.label for i in self.inds], np.int32)

    @property
    def is_complex(self) -> bool:
        return self.dtype == ITB_C64

    @property
    def nreal(self) -> int:
        return self.nelems * (2 if self.is_complex else 1)

    def desc(self) -> TensorDesc:
        d = TensorDesc()
        d.order = self.order
        d.dtype = self.dtype
        d.nsect = _i32p(self.nsect)
        d.sect = _i64p(self.sect)
        d.nblocks = self.nblocks
        d.blocks = _i32p(self.blocks)
        d.offsets = _i64p(self.offsets)
        d.nelems = self.nelems
        return d


class ContractPlan:
    """Host-computed block-pair tables for C = A*B (itb_contract_plan)."""

    def __init__(self, A: BlockStruct, B: BlockStruct):
        self.A, self.B = A, B
        self._h = C.c_void_p()
        # labels are assigned per contraction: position of the (id, plev) pair among the distinct indices of A and B
        # (Index.label = id*8+plev would collide for plev >= 8 and overflow int32 for large ids)
        key_of = lambda ix: (ix.id, ix.plev)
        assert len({key_of(i) for i in A.inds}) == A.order, "contract: A carries the same index twice"
        assert len({key_of(i) for i in B.inds}) == B.order, "contract: B carries the same index twice"
        pos, by_label = {}, {}
        for ix in list(A.inds) + list(B.inds):
            if key_of(ix) not in pos:
                pos[key_of(ix)] = len(pos)
                by_label[pos[key_of(ix)]] = ix
        la = np.array([pos[key_of(i)] for i in A.inds], np.int32)
        lb = np.array([pos[key_of(i)] for i in B.inds], np.int32)
        da, db = A.desc(), B.desc()
        check(lib().itb_contract_plan_create(C.byref(da), _i32p(la), C.byref(db), _i32p(lb), C.byref(self._h)))
        info = ContractInfo()
        check(lib().itb_contract_plan_info(self._h, C.byref(info)))
        self.info = info
        r = info.c_order
        labels = np.zeros(max(r, 1), np.int32)
        lib().itb_contract_plan_c_labels(self._h, _i32p(labels))
        c_inds = [by_label[int(l)] for l in labels[:r]]
        blocks = np.zeros((max(info.c_nblocks, 1), max(r, 1)), np.int32)
        lib().itb_contract_plan_c_blocks(self._h, _i32p(blocks))
        offsets = np.zeros(max(info.c_nblocks, 1), np.int64)
        lib().itb_contract_plan_c_offsets(self._h, _i64p(offsets))
        cb = np.ascontiguousarray(blocks.reshape(-1)[: info.c_nblocks * r]).reshape(info.c_nblocks, r)
        self.C = BlockStruct(c_inds, cb, info.c_dtype, offsets[: info.c_nblocks])
        assert self.C.nelems == info.c_nelems, (self.C.nelems, info.c_nelems)
        self.flops = info.flops
        self.npairs = info.npairs

    def pairs(self) -> np.ndarray:
        t = np.zeros((max(self.npairs, 1), 3), np.int64)
        lib().itb_contract_plan_pairs(self._h, _i64p(t))
        return t[: self.npairs]

    def set_cblock_range(self, first: int, last: int) -> None:
        check(lib().itb_contract_plan_set_cblock_range(self._h, first, last))
        check(lib().itb_contract_plan_info(self._h, C.byref(self.info)))

    def set_cblock_mask(self, mask) -> None:
        if mask is None:
            check(lib().itb_contract_plan_set_cblock_mask(self._h, None))
        else:
            m = np.ascontiguousarray(mask, np.uint8)
            assert m.shape == (self.C.nblocks,)
            check(lib().itb_contract_plan_set_cblock_mask(self._h, m.ctypes.data_as(C.POINTER(C.c_uint8))))
        check(lib().itb_contract_plan_info(self._h, C.byref(self.info)))

    def set_index_slices(self, c_index: int, lo, hi) -> None:
        """restrict execution to rows [lo[s], hi[s]) of C's index c_index in every block whose coordinate there is sector s
        (None: remove the restriction) — the multi-GPU sharding unit inside QN sectors"""
        if lo is None:
            check(lib().itb_contract_plan_set_index_slices(self._h, -1, None, None))
        else:
            lo_, hi_ = np.ascontiguousarray(lo, np.int64), np.ascontiguousarray(hi, np.int64)
            assert lo_.shape == hi_.shape == (self.C.inds[c_index].nsect,)
            check(lib().itb_contract_plan_set_index_slices(self._h, int(c_index), _i64p(lo_), _i64p(hi_)))
        check(lib().itb_contract_plan_info(self._h, C.byref(self.info)))

    def executed_flops(self) -> float:
        """flops of the C blocks / slices this plan executes (== flops when nothing is masked or sliced)"""
        return float(sum(self.info.class_flops))

    def model_cycles(self, sms: int = 148, stream_bytes_per_cycle: float = 2100.0, launch_cycles: float = 30000.0) -> float:
        """modelled device time of the plan as currently sliced, in SM cycles: tile cycles / grid width + streaming bytes at
        the measured 4.2 TB/s (2100 B per 1.965 GHz cycle) + a fixed launch / ramp / reduce cost when there is any work"""
        cyc, byt = C.c_double(), C.c_double()
        check(lib().itb_contract_plan_model_work(self._h, C.byref(cyc), C.byref(byt)))
        t = cyc.value / sms + byt.value / stream_bytes_per_cycle
        return t + (launch_cycles if t > 0 else 0.0)

    def refine(self, ctx, a_ptr, b_ptr, c_ptr, rounds: int = 3) -> float:
        """itb_contract_plan_refine: re-partition the tile work from measured per-CTA cycles; returns longest CTA span
        before / after"""
        gain = C.c_double(1.0)
        check(lib().itb_contract_plan_refine(ctx.handle, self._h, a_ptr, b_ptr, c_ptr, rounds, C.byref(gain)))
        return gain.value

    def close(self):
        if self._h:
            lib().itb_contract_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PermutePlan:
    def __init__(self, S: BlockStruct, D: BlockStruct, perm: Sequence[int]):
        self.S, self.D = S, D
        self.perm = np.array(list(perm), np.int32)
        self._h = C.c_void_p()
        ds, dd = S.desc(), D.desc()
        check(lib().itb_permute_plan_create(C.byref(ds), C.byref(dd), _i32p(self.perm), C.byref(self._h)))
        self.bytes = int(lib().itb_permute_plan_bytes(self._h))

    def close(self):
        if self._h:
            lib().itb_permute_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One per (process, GPU): wraps itb_ctx and adopts torch's current CUDA stream."""

    def __init__(self, device: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("itensor_b200.Context needs a CUDA device: there is no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(device)
        # the library's background read of the cuSOLVER/cuBLAS objects (solver warm-up for long DMRG runs, see
        # itb_solver_ready in include/itb200.h) only competes for I/O in the short test / bench processes this Python
        # mirror serves: off unless asked for
        import os

        os.environ.setdefault("ITB_WARM_LIBS", "0")
        self._h = C.c_void_p()
        check(lib().itb_ctx_create(device, C.byref(self._h)))
        self.stream = torch.cuda.current_stream(self.device)
        check(lib().itb_ctx_set_stream(self._h, C.c_void_p(self.stream.cuda_stream)))

    @property
    def handle(self):
        return self._h

    def launches(self) -> int:
        return int(lib().itb_launch_count(self._h))

    def empty(self, nreal: int):
        return self.torch.empty(max(nreal, 1), dtype=self.torch.float64, device=self.device)[:nreal]

    def close(self):
        if self._h:
            lib().itb_ctx_destroy(self._h)
            self._h = C.c_void_p()


@dataclass
class QTensor:
    """Block-sparse (or dense) tensor resident in HBM: structure on the host, flat reals on the device."""

    ctx: Context
    struct: BlockStruct
    data: object  # torch.float64 tensor of struct.nreal elements on ctx.device

    @property
    def inds(self):
        return self.struct.inds

    @property
    def ptr(self):
        return C.c_void_p(self.data.data_ptr())

    @staticmethod
    def from_host(ctx: Context, struct: BlockStruct, host: np.ndarray) -> "QTensor":
        t = ctx.torch.from_numpy(np.ascontiguousarray(host).view(np.float64).reshape(-1))
        assert t.numel() == struct.nreal, (t.numel(), struct.nreal)
        return QTensor(ctx, struct, t.to(ctx.device))

    def to_host(self) -> np.ndarray:
        a = self.data.cpu().numpy()
        return a.view(np.complex128) if self.struct.is_complex else a

    def clone(self) -> "QTensor":
        return QTensor(self.ctx, self.struct, self.data.clone())


def contract(A: QTensor, B: QTensor, plan: Optional[ContractPlan] = None) -> QTensor:
    """C = A*B on the device (doTask(Contract,...) for QDense/Dense pairs)."""
    ctx = A.ctx
    if plan is None:
        plan = ContractPlan(A.struct, B.struct)
    out = ctx.empty(plan.C.nreal)
    check(lib().itb_contract_run(ctx.handle, plan._h, A.ptr, B.ptr, C.c_void_p(out.data_ptr())))
    return QTensor(ctx, plan.C, out)


def permuted_struct(S: BlockStruct, new_inds: Sequence[Index], flux: Optional[Sequence[int]] = None, dtype: Optional[int] = None):
    """Destination structure of permute(): index order new_inds; QN tensors get EVERY flux-allowed
    block (permuteQDense builds QDense(Bis,div), itdata/qdense.cc:862), others keep the block list."""
    perm = []
    for ix in S.inds:
        pos = [j for j, nj in enumerate(new_inds) if nj.same(ix)]
        assert len(pos) == 1, "permute: index sets differ"
        perm.append(pos[0])
    ordered = [None] * S.order
    for i, p in enumerate(perm):
        ordered[p] = S.inds[i]
    dt = S.dtype if dtype is None else dtype
    if flux is not None and S.order > 0 and S.inds[0].qns is not None:
        blocks = flux_blocks(ordered, flux)
    else:
        pb = np.zeros_like(S.blocks)
        for i, p in enumerate(perm):
            pb[:, p] = S.blocks[:, i]
        # reference order: last index most significant
        keys = tuple(pb[:, j] for j in range(S.order)) if S.order else ()
        order = np.lexsort(keys) if S.order and S.nblocks else np.arange(S.nblocks)
        blocks = pb[order]
    return BlockStruct(ordered, blocks, dt), perm


def permute(T: QTensor, new_inds: Sequence[Index], flux: Optional[Sequence[int]] = None) -> QTensor:
    D, perm = permuted_struct(T.struct, new_inds, flux)
    plan = PermutePlan(T.struct, D, perm)
    out = T.ctx.empty(D.nreal)
    check(lib().itb_permute_run(T.ctx.handle, plan._h, T.ptr, C.c_void_p(out.data_ptr()), 1.0, 0.0, 0))
    return QTensor(T.ctx, D, out)


def _merge_blocks(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    r = a.shape[1]
    allb = np.unique(np.concatenate([a, b], axis=0), axis=0) if r else a[:1]
    if r and len(allb):
        allb = allb[np.lexsort(tuple(allb[:, j] for j in range(r)))]
    return allb.astype(np.int32)


def add(A: QTensor, alpha: complex, B: QTensor) -> QTensor:
    """A += alpha*B with B's indices in any order; widens A's block list / promotes to complex when
    needed (doTask(PlusEQ,QDense,QDense), itdata/qdense.cc:551-668). Returns the updated tensor."""
    ctx = A.ctx
    alpha = complex(alpha)
    perm = []
    for ix in B.inds:
        pos = [j for j, aj in enumerate(A.inds) if aj.same(ix)]
        assert len(pos) == 1, "add: index sets differ"
        perm.append(pos[0])
    pb = np.zeros_like(B.struct.blocks)
    for i, p in enumerate(perm):
        pb[:, p] = B.struct.blocks[:, i]
    need_cplx = (B.struct.is_complex or alpha.imag != 0.0) and not A.struct.is_complex
    have = {tuple(r) for r in A.struct.blocks.tolist()}
    missing = any(tuple(r) not in have for r in pb.tolist())
    dst = A
    if missing or need_cplx:
        blocks = _merge_blocks(A.struct.blocks, pb) if missing else A.struct.blocks
        ns = BlockStruct(A.inds, blocks, ITB_C64 if (need_cplx or A.struct.is_complex) else ITB_F64)
        buf = ctx.empty(ns.nreal)
        p0 = PermutePlan(A.struct, ns, list(range(A.struct.order)))
        check(lib().itb_permute_run(ctx.handle, p0._h, A.ptr, C.c_void_p(buf.data_ptr()), 1.0, 0.0, 0))
        dst = QTensor(ctx, ns, buf)
    trivial = perm == list(range(len(perm))) and np.array_equal(dst.struct.blocks, B.struct.blocks) and dst.struct.dtype == B.struct.dtype
    if trivial:
        check(lib().itb_axpy(ctx.handle, dst.struct.dtype, dst.struct.nelems, alpha.real, alpha.imag, B.ptr, dst.ptr))
    else:
        plan = PermutePlan(B.struct, dst.struct, perm)
        check(lib().itb_permute_run(ctx.handle, plan._h, B.ptr, dst.ptr, alpha.real, alpha.imag, 1))
    return dst


def norm(T: QTensor) -> float:
    out = C.c_double()
    check(lib().itb_nrm2(T.ctx.handle, T.struct.dtype, T.struct.nelems, T.ptr, C.byref(out)))
    return out.value


def scale(T: QTensor, alpha: complex) -> QTensor:
    alpha = complex(alpha)
    if alpha.imag != 0.0 and not T.struct.is_complex:
        ns = BlockStruct(T.inds, T.struct.blocks, ITB_C64)
        buf = T.ctx.empty(ns.nreal)
        check(lib().itb_real_to_cplx(T.ctx.handle, T.struct.nelems, T.ptr, C.c_void_p(buf.data_ptr())))
        T = QTensor(T.ctx, ns, buf)
    check(lib().itb_scal(T.ctx.handle, T.struct.dtype, T.struct.nelems, T.ptr, alpha.real, alpha.imag))
    return T


def fill(T: QTensor, val: complex) -> QTensor:
    val = complex(val)
    check(lib().itb_fill(T.ctx.handle, T.struct.dtype, T.struct.nelems, T.ptr, val.real, val.imag))
    return T


def dag(T: QTensor) -> QTensor:
    """Reverse arrows and conjugate (a copy; ITensor::dag + doTask(Conj))."""
    ns = BlockStruct([i.dag() for i in T.inds], T.struct.blocks, T.struct.dtype)
    out = QTensor(T.ctx, ns, T.data.clone())
    if T.struct.is_complex:
        check(lib().itb_conj(T.ctx.handle, T.struct.nelems, out.ptr))
    return out


def elt(T: QTensor, offset: int = 0) -> complex:
    out = (C.c_double * 2)()
    check(lib().itb_get_elt(T.ctx.handle, T.struct.dtype, T.ptr, offset, out))
    return complex(out[0], out[1])
