"""Multi-GPU sharding of the H_eff*phi chain by OUTPUT rows.

The primed left link l' is an uncontracted index of every intermediate of LocalOp::product
(phi*L -> *W1 -> *W2 -> *R, itensor/mps/localop.h:346-362), so partitioning the RANGE of l' makes every
intermediate local to its rank: rank g computes, in every step, exactly the rows of l' it owns — the same "one
owner per piece of C" rule as the reference's OpenMP path (itensor/itdata/qutil.h:285-348), refined from whole C
blocks to row ranges inside a QN sector so that the ranks carry equal flops whatever the sector sizes are (a whole
sector can hold 46 % of the work). The unit handed to the planner is itb_contract_plan_set_index_slices.

The only communication is re-replicating H*phi once per product: every rank packs the rows it owns (strided boxes of
the blocks of H*phi) into one contiguous segment, ONE all-gather over NVLink moves the segments (each element crosses
the wire once, half the bytes of an all-reduce of the zero-padded buffer), and the segments of the other ranks are
scattered back into place. Pack and scatter are launches of the block-copy kernel (itb_blockcopy_plan_create).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np


def sector_assignment(weights: Sequence[float], world: int) -> np.ndarray:
    """LPT (longest processing time first) assignment of whole sectors to ranks; returns rank of every sector.
    (coarse fallback when a plan cannot be row-sliced, see ChainShard)"""
    order = np.argsort(-np.asarray(weights, float), kind="stable")
    load = np.zeros(world)
    owner = np.zeros(len(weights), np.int64)
    for s in order:
        g = int(np.argmin(load))
        owner[s] = g
        load[g] += weights[s]
    return owner


def row_partition(sizes: Sequence[int], weights: Sequence[float], world: int, align: int = 8, shares: Optional[Sequence[float]] = None,
                  return_cuts: bool = False):
    """Cut the concatenated rows of all sectors into `world` contiguous segments of equal weight (or of the given
    `shares` of the total weight, one per rank: the measured re-balancing of bench.py hands a slower rank less).

    sizes[s] rows in sector s carrying weights[s] in total (uniform inside a sector). Returns lo, hi of shape
    (world, nsect): rank g owns rows [lo[g,s], hi[g,s]) of sector s. Cut points inside a sector are rounded to a
    multiple of `align` rows (DMMA fragments are 8 rows) when that keeps them inside the sector."""
    sizes = np.asarray(sizes, np.int64)
    w = np.asarray(weights, float)
    per_row = np.where(sizes > 0, w / np.maximum(sizes, 1), 0.0)
    start = np.concatenate([[0], np.cumsum(sizes)])          # first global row of every sector
    cumw = np.concatenate([[0.0], np.cumsum(w)])
    total = cumw[-1]
    cuts = [0]
    cum_share = np.arange(world + 1) / world if shares is None else np.concatenate([[0.0], np.cumsum(np.asarray(shares, float) / np.sum(shares))])
    for g in range(1, world):
        target = total * cum_share[g]
        s = int(np.searchsorted(cumw, target, side="right") - 1)
        s = min(max(s, 0), len(sizes) - 1)
        r = int(round((target - cumw[s]) / per_row[s])) if per_row[s] > 0 else 0
        if align > 1 and 0 < r < sizes[s]:
            ra = int(round(r / align)) * align
            if 0 < ra < sizes[s]:
                r = ra
        r = min(max(r, 0), int(sizes[s]))
        cuts.append(max(int(start[s]) + r, cuts[-1]))
    cuts.append(int(start[-1]))
    lo = np.zeros((world, len(sizes)), np.int64)
    hi = np.zeros((world, len(sizes)), np.int64)
    for g in range(world):
        a, b = cuts[g], cuts[g + 1]
        lo[g] = np.clip(a - start[:-1], 0, sizes)
        hi[g] = np.clip(b - start[:-1], 0, sizes)
    if return_cuts:
        return lo, hi, cuts
    return lo, hi


def ranges_from_cuts(sizes: Sequence[int], cuts: Sequence[int], world: int):
    """lo, hi (world x nsect) of the contiguous global row segments [cuts[g], cuts[g+1])"""
    sizes = np.asarray(sizes, np.int64)
    start = np.concatenate([[0], np.cumsum(sizes)])
    lo = np.zeros((world, len(sizes)), np.int64)
    hi = np.zeros((world, len(sizes)), np.int64)
    for g in range(world):
        lo[g] = np.clip(cuts[g] - start[:-1], 0, sizes)
        hi[g] = np.clip(cuts[g + 1] - start[:-1], 0, sizes)
    return lo, hi


def optimise_cuts(plans, pos, sizes, cuts, world, factors=None, sweeps: int = 2):
    """Move the cut points of a row partition to where the planner's CYCLE MODEL (itb_contract_plan_model_work) says the
    slowest rank is fastest. Equal flops is only the starting point: a cut inside a sector leaves each side with a
    remainder tile row that costs almost a full 128-row tile per K-chunk whatever its height, so a cut on a multiple of the
    tile height from the sector start is free while one 8 rows further costs a tile row on both sides. Candidates per cut:
    where it is, the neighbouring multiples of 128 / 64 rows inside its sector, and small shifts; a move is accepted when
    it lowers the larger of the two adjacent ranks' modelled times (factors[r]: measured / modelled time of rank r)."""
    sizes = np.asarray(sizes, np.int64)
    start = np.concatenate([[0], np.cumsum(sizes)])
    total = int(start[-1])
    f = np.ones(world) if factors is None else np.asarray(factors, float)
    cuts = [int(c) for c in cuts]
    cache = {}

    def cost(r, a, b):
        key = (r, a, b)
        if key not in cache:
            lo = np.clip(a - start[:-1], 0, sizes)
            hi = np.clip(b - start[:-1], 0, sizes)
            t = 0.0
            for p, j in zip(plans, pos):
                p.set_index_slices(j, lo, hi)
                t += p.model_cycles()
            cache[key] = t
        return cache[key] * f[r]

    for _ in range(sweeps):
        moved = False
        for g in range(1, world):
            c = cuts[g]
            s = int(np.searchsorted(start, c, side="right") - 1)
            s = min(max(s, 0), len(sizes) - 1)
            cands = {c}
            for snap in (128, 64):
                k = (c - int(start[s])) // snap
                for kk in (k, k + 1):
                    cands.add(int(start[s]) + kk * snap)
            for d in (-32, -16, -8, 8, 16, 32):
                cands.add(c + d)
            cands.add(int(start[s]))
            cands.add(int(start[s + 1]))
            best, best_key = c, None
            for x in sorted(cands):
                if not (cuts[g - 1] < x < cuts[g + 1]) or not (0 < x < total):
                    continue
                a, b = cost(g - 1, cuts[g - 1], x), cost(g, x, cuts[g + 1])
                key = (max(a, b), a + b)
                if best_key is None or key < best_key:
                    best, best_key = x, key
            if best != c:
                cuts[g] = best
                moved = True
        if not moved:
            break
    times = [cost(r, cuts[r], cuts[r + 1]) / f[r] for r in range(world)]
    return cuts, times


def pair_flops(plan) -> np.ndarray:
    """flops of every C block of a plan (2*M*N*K summed over its pairs, complex multipliers included)"""
    A, B = plan.A, plan.B
    keyb = {(ix.id, ix.plev) for ix in B.inds}
    cont = [i for i, ix in enumerate(A.inds) if (ix.id, ix.plev) in keyb]
    out = np.zeros(plan.C.nblocks)
    csize = plan.C.block_sizes().astype(float)
    mult = (2.0 if A.is_complex else 1.0) * (2.0 if B.is_complex else 1.0)
    for ia, ib, ic in plan.pairs():
        k = 1.0
        for i in cont:
            k *= A.inds[i].sizes[A.blocks[ia, i]]
        out[ic] += 2.0 * csize[ic] * k * mult
    return out


class CopyItem(C.Structure):  # itb_copy_item (include/itb200.h)
    _fields_ = [("s_off", C.c_int64), ("d_off", C.c_int64), ("n", C.c_int32), ("pad_", C.c_int32),
                ("ext", C.c_int64 * 12), ("sstr", C.c_int64 * 12), ("dstr", C.c_int64 * 12)]


def _owned_boxes(struct, j: int, lo: np.ndarray, hi: np.ndarray):
    """(block offset shift, box shape, full strides) of the rows [lo[s],hi[s]) of index j in every block, in block order"""
    out = []
    for b in range(struct.nblocks):
        shape = struct.block_shape(b)
        s = int(struct.blocks[b, j])
        a, e = int(lo[s]), int(hi[s])
        if e <= a:
            continue
        strides, st = [], 1
        for d in shape:
            strides.append(st)
            st *= d
        box = list(shape)
        box[j] = e - a
        out.append((int(struct.offsets[b]) + a * strides[j], box, strides))
    return out


class ChainShard:
    """Row partition of l' over `world` ranks for a chain of plans; slices every plan to this rank's rows."""

    def __init__(self, plans: List, world: int, rank: int, shard_index_id: int = 1, shard_index_plev: int = 1, align: int = 8,
                 shares: Optional[Sequence[float]] = None, cuts: Optional[Sequence[int]] = None, optimise: bool = False,
                 factors: Optional[Sequence[float]] = None):
        self.world, self.rank, self.plans = world, rank, plans
        pos, flops = [], []
        sizes = None
        for p in plans:
            j = [t for t, ix in enumerate(p.C.inds) if ix.id == shard_index_id and ix.plev == shard_index_plev]
            assert len(j) == 1, "the sharding index must be an uncontracted index of every step"
            pos.append(j[0])
            sizes = p.C.inds[j[0]].sizes
            flops.append(pair_flops(p))
        self.pos = pos
        w = np.zeros(len(sizes))
        for p, j, f in zip(plans, pos, flops):
            np.add.at(w, p.C.blocks[:, j], f)
        self.sector_work = w
        self.total_flops = float(sum(f.sum() for f in flops))
        self.shares = None if shares is None else [float(x) for x in shares]
        self.lo, self.hi, flop_cuts = row_partition(sizes, w, world, align, shares, return_cuts=True)
        self.mode = "rows"
        self.model_ms = None
        try:
            if cuts is None:
                cuts = flop_cuts
            if optimise:
                cuts, cyc = optimise_cuts(plans, pos, sizes, cuts, world, factors)
                self.model_ms = [c / 1.965e6 for c in cyc]
            self.cuts = [int(c) for c in cuts]
            self.lo, self.hi = ranges_from_cuts(sizes, self.cuts, world)
            for p, j in zip(plans, pos):
                p.set_index_slices(j, self.lo[rank], self.hi[rank])
        except Exception:
            # a step whose sliced index is not the slowest non-unit uncontracted index of its operand (dense site
            # indices): fall back to whole sectors for the whole chain
            self.mode = "sectors"
            owner = sector_assignment(w, world)
            sz = np.asarray(sizes, np.int64)
            self.lo = np.zeros((world, len(sz)), np.int64)
            self.hi = np.stack([np.where(owner == g, sz, 0) for g in range(world)])
            for p, j in zip(plans, pos):
                p.set_index_slices(j, self.lo[rank], self.hi[rank])
        self.my_flops = float(sum(p.executed_flops() for p in plans))
        # packed segments of H*phi (the last plan's C): rank r's rows, block after block, each box contiguous
        out = plans[-1].C
        self.out_struct = out
        cs = 2 if out.is_complex else 1
        self.seg_elems = []
        self._boxes = []
        for r in range(world):
            boxes = _owned_boxes(out, pos[-1], self.lo[r], self.hi[r])
            self._boxes.append(boxes)
            self.seg_elems.append(int(sum(int(np.prod(b[1])) for b in boxes)))
        assert sum(self.seg_elems) == out.nelems
        self.seg_max = max(self.seg_elems)          # elements (of the tensor's dtype) per padded segment
        self.seg_reals = self.seg_max * cs
        self._pack = self._unpack = None
        self.send = self.recv = None

    # ---- device side -----------------------------------------------------------------------------------------
    def _items(self, ranks, to_packed: bool):
        items = []
        for r in ranks:
            run = r * self.seg_max if not to_packed else 0
            for off, box, strides in self._boxes[r]:
                it = CopyItem()
                it.n = len(box)
                packed, st = [], 1
                for d in box:
                    packed.append(st)
                    st *= d
                for d in range(len(box)):
                    it.ext[d] = box[d]
                    it.sstr[d] = strides[d] if to_packed else packed[d]
                    it.dstr[d] = packed[d] if to_packed else strides[d]
                it.s_off = off if to_packed else run
                it.d_off = run if to_packed else off
                run += st
                items.append(it)
        return items

    def _plan(self, items):
        from ._lib import check, lib

        arr = (CopyItem * max(len(items), 1))(*items)
        h = C.c_void_p()
        dt = self.out_struct.dtype
        check(lib().itb_blockcopy_plan_create(len(items), C.cast(arr, C.c_void_p), dt, dt, C.byref(h)))
        return h

    def prepare(self, alloc):
        """build the pack / scatter plans and the segment buffers; alloc(nreal) returns a flat float64 device tensor"""
        self._pack = self._plan(self._items([self.rank], True))
        self._unpack = self._plan(self._items([r for r in range(self.world) if r != self.rank], False))
        self.recv = alloc(self.seg_reals * self.world)
        self.send = self.recv[self.rank * self.seg_reals:(self.rank + 1) * self.seg_reals]  # in-place all-gather layout
        return self

    def allgather(self, ctx_handle, flat, events=None) -> None:
        """re-replicate H*phi: pack own rows -> one all-gather of the segments -> scatter the other ranks' rows into place
        (events: optional list of 4 CUDA events recorded around the three phases, for the bench's phase breakdown)"""
        import torch.distributed as dist

        from ._lib import check, lib

        p = lambda t: C.c_void_p(t.data_ptr())
        if events:
            events[0].record()
        check(lib().itb_permute_run(ctx_handle, self._pack, p(flat), p(self.send), 1.0, 0.0, 0))
        if events:
            events[1].record()
        if dist.get_backend() == "nccl":
            dist.all_gather_into_tensor(self.recv, self.send)
        else:  # gloo (CPU tests)
            parts = [self.recv[r * self.seg_reals:(r + 1) * self.seg_reals] for r in range(self.world)]
            dist.all_gather(parts, self.send.clone())
        if events:
            events[2].record()
        check(lib().itb_permute_run(ctx_handle, self._unpack, p(self.recv), p(flat), 1.0, 0.0, 0))
        if events:
            events[3].record()

    # ---- direct exchange over peer memory (NVLink stores, no NCCL on the data path) -----------------------------------
    def prepare_p2p(self, ctx, nbuf: int = 2):
        """H*phi lives in `nbuf` peer-mapped buffers per rank (itb_p2p_alloc / itb_p2p_open); returns the list of local
        output tensors (flat float64, one per buffer) or None when peer memory is unavailable (caller keeps the all-gather).

        push(k) then stores the rows this rank owns straight into buffer k of EVERY peer — one launch of the block-copy
        kernel whose items carry, per peer, the owned boxes with the destination offset shifted into that peer's mapping —
        followed by a one-element all-reduce that tells every rank all pushes into its buffer have landed. Two buffers
        used alternately make a single barrier per product sufficient: a rank can run at most one barrier ahead, so it
        writes buffer k+1 while a slower peer may still be reading buffer k, never the buffer being read."""
        import torch
        import torch.distributed as dist

        from ._lib import check, lib

        nreal = self.out_struct.nreal
        self._p2p_ctx = ctx
        self._p2p_local, self._p2p_peers, self._p2p_push, self._p2p_tensors = [], [], [], []
        self._p2p_flags = None
        try:
            handles = []
            for _ in range(nbuf):
                ptr = C.c_void_p()
                h = C.create_string_buffer(64)
                check(lib().itb_p2p_alloc(ctx.handle, max(nreal, 1) * 8, C.byref(ptr), h))
                self._p2p_local.append(ptr.value)
                handles.append(h.raw)
            # flag block of the arrival barrier (itb_p2p_barrier)
            ptr = C.c_void_p()
            h = C.create_string_buffer(64)
            check(lib().itb_p2p_alloc(ctx.handle, 256, C.byref(ptr), h))
            check(lib().itb_p2p_barrier_init(ctx.handle, ptr))
            self._p2p_flags = ptr.value
            handles.append(h.raw)
            ok = 1
        except Exception:  # noqa: BLE001
            handles, ok = [], 0
        flag = torch.tensor([ok], dtype=torch.int32, device=ctx.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        allh = [None] * self.world
        dist.all_gather_object(allh, handles)
        if int(flag.item()) == 0:
            self.close_p2p()
            return None
        try:
            for b in range(nbuf):
                peers = {}
                for r in range(self.world):
                    if r == self.rank:
                        continue
                    pp = C.c_void_p()
                    check(lib().itb_p2p_open(ctx.handle, allh[r][b], C.byref(pp)))
                    peers[r] = pp.value
                self._p2p_peers.append(peers)
            self._p2p_flag_peers = (C.c_void_p * self.world)()
            for r in range(self.world):
                if r == self.rank:
                    continue
                pp = C.c_void_p()
                check(lib().itb_p2p_open(ctx.handle, allh[r][nbuf], C.byref(pp)))
                self._p2p_flag_peers[r] = pp.value
            ok = 1
        except Exception:  # noqa: BLE001
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=ctx.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            self.close_p2p()
            return None
        cs = 2 if self.out_struct.is_complex else 1
        for b in range(nbuf):
            items = []
            for r, pptr in self._p2p_peers[b].items():
                delta = pptr - self._p2p_local[b]
                assert delta % (8 * cs) == 0
                for off, box, strides in self._boxes[self.rank]:
                    it = CopyItem()
                    it.n = len(box)
                    for d in range(len(box)):
                        it.ext[d] = box[d]
                        it.sstr[d] = strides[d]
                        it.dstr[d] = strides[d]
                    it.s_off = off
                    it.d_off = off + delta // (8 * cs)   # (offsets are in elements of the tensor's dtype)
                    items.append(it)
            self._p2p_push.append(self._plan(items))

            class _Raw:  # zero-copy torch view of the peer-mappable buffer
                pass

            raw = _Raw()
            raw.__cuda_array_interface__ = {"shape": (max(nreal, 1),), "typestr": "<f8", "data": (self._p2p_local[b], False), "version": 3, "strides": None}
            self._p2p_tensors.append(torch.as_tensor(raw, device=ctx.device)[:nreal])
        self._p2p_token = torch.zeros(1, dtype=torch.float32, device=ctx.device)
        self.flag_barrier = True   # False: one-element NCCL all-reduce instead of the flag kernel
        dist.barrier()             # every rank has zeroed its flag block and mapped its peers' before the first push
        return self._p2p_tensors

    def barrier_status(self):
        """(epochs completed, epoch at which a bounded wait expired or 0) of this rank's flag block"""
        from ._lib import check, lib

        e, err = C.c_int64(), C.c_int64()
        check(lib().itb_p2p_barrier_status(self._p2p_ctx.handle, C.c_void_p(self._p2p_flags), C.byref(e), C.byref(err)))
        return int(e.value), int(err.value)

    def push(self, ctx_handle, b: int, events=None) -> None:
        """store this rank's rows of H*phi (local buffer b) into buffer b of every peer, then the arrival barrier"""
        import torch.distributed as dist

        from ._lib import check, lib

        base = C.c_void_p(self._p2p_local[b])
        if events:
            events[0].record()
        check(lib().itb_permute_run(ctx_handle, self._p2p_push[b], base, base, 1.0, 0.0, 0))
        if events:
            events[1].record()
        if self.flag_barrier:
            check(lib().itb_p2p_barrier(ctx_handle, C.c_void_p(self._p2p_flags), self._p2p_flag_peers, self.world, self.rank))
        else:
            dist.all_reduce(self._p2p_token)
        if events:
            events[2].record()

    def run_last_mirrored(self, ctx_handle, plan, a_ptr, b_ptr, b: int, barrier: bool = True) -> None:
        """the chain's last contraction writing local buffer b AND, from the kernel's epilogue, every peer's buffer b
        (itb_contract_run_mirrored), followed by the arrival barrier: no separate push. Raises when the plan has work
        outside the static tile class (use push() then)."""
        from ._lib import check, lib

        peers = self._p2p_peers[b]
        arr = (C.c_void_p * max(len(peers), 1))(*[peers[r] for r in sorted(peers)])
        check(lib().itb_contract_run_mirrored(ctx_handle, plan._h, a_ptr, b_ptr, C.c_void_p(self._p2p_local[b]), len(peers), arr))
        if barrier:
            self.barrier(ctx_handle)

    def barrier(self, ctx_handle) -> None:
        import torch.distributed as dist

        from ._lib import check, lib

        if self.flag_barrier:
            check(lib().itb_p2p_barrier(ctx_handle, C.c_void_p(self._p2p_flags), self._p2p_flag_peers, self.world, self.rank))
        else:
            dist.all_reduce(self._p2p_token)

    def pack_own(self, ctx_handle, flat) -> None:
        """this rank's rows of `flat` -> its packed segment (self.send): what an end-to-end caller reads back"""
        from ._lib import check, lib

        p = lambda t: C.c_void_p(t.data_ptr())
        check(lib().itb_permute_run(ctx_handle, self._pack, p(flat), p(self.send), 1.0, 0.0, 0))

    def close_p2p(self):
        from ._lib import lib

        ctx = getattr(self, "_p2p_ctx", None)
        for h in getattr(self, "_p2p_push", []):
            lib().itb_permute_plan_destroy(h)
        for peers in getattr(self, "_p2p_peers", []):
            for pp in peers.values():
                lib().itb_p2p_close(ctx.handle if ctx else None, C.c_void_p(pp))
        self._p2p_tensors = []
        fp = getattr(self, "_p2p_flag_peers", None)
        if fp is not None:
            for r in range(self.world):
                if fp[r]:
                    lib().itb_p2p_close(ctx.handle if ctx else None, C.c_void_p(fp[r]))
            self._p2p_flag_peers = None
        if getattr(self, "_p2p_flags", None):
            lib().itb_p2p_free(ctx.handle if ctx else None, C.c_void_p(self._p2p_flags))
            self._p2p_flags = None
        for p in getattr(self, "_p2p_local", []):
            lib().itb_p2p_free(ctx.handle if ctx else None, C.c_void_p(p))
        self._p2p_push, self._p2p_peers, self._p2p_local = [], [], []

    def close(self):
        from ._lib import lib

        for h in (self._pack, self._unpack):
            if h:
                lib().itb_permute_plan_destroy(h)
        self._pack = self._unpack = None


def shard_chain(plans, world, rank, shares=None, cuts=None, optimise=False, factors=None) -> ChainShard:
    return ChainShard(plans, world, rank, shares=shares, cuts=cuts, optimise=optimise, factors=factors)
