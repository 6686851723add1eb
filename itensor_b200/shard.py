"""Multi-GPU sharding of the H_eff*phi chain by OUTPUT blocks.

The primed left link l' is an uncontracted index of every intermediate of LocalOp::product
(phi*L -> *W1 -> *W2 -> *R, itensor/mps/localop.h:346-362), so assigning the QN sectors of l' to ranks makes
every intermediate local to its rank: rank g computes exactly the C blocks whose l' coordinate is in its
sector set (the same "all pairs of one C block go to one worker" rule as the reference's OpenMP path,
itensor/itdata/qutil.h:285-348). The only communication is re-replicating H*phi once per product.

Each element of H*phi is produced by exactly one rank, the others hold zeros there, so the replication is an
exact (order-independent, bit-reproducible) sum: one NCCL all-reduce over NVLink on the flat buffer. (A packed
all-gather would move half the bytes)
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def sector_assignment(weights: Sequence[float], world: int) -> np.ndarray:
    """LPT (longest processing time first) assignment of sectors to ranks; returns rank of every sector."""
    order = np.argsort(-np.asarray(weights, float), kind="stable")
    load = np.zeros(world)
    owner = np.zeros(len(weights), np.int64)
    for s in order:
        g = int(np.argmin(load))
        owner[s] = g
        load[g] += weights[s]
    return owner


def pair_flops(plan) -> np.ndarray:
    """flops of every C block of a plan (2*M*N*K summed over its pairs, complex multipliers included)"""
    A, B = plan.A, plan.B
    lab_b = set(int(x) for x in B.labels)
    cont = [i for i, l in enumerate(A.labels) if int(l) in lab_b]
    out = np.zeros(plan.C.nblocks)
    csize = plan.C.block_sizes().astype(float)
    mult = (2.0 if A.is_complex else 1.0) * (2.0 if B.is_complex else 1.0)
    for ia, ib, ic in plan.pairs():
        k = 1.0
        for i in cont:
            k *= A.inds[i].sizes[A.blocks[ia, i]]
        out[ic] += 2.0 * csize[ic] * k * mult
    return out


class ChainShard:
    def __init__(self, plans: List, world: int, rank: int, shard_index_id: int = 1, shard_index_plev: int = 1):
        self.world, self.rank = world, rank
        # position of l' in every step's C and the per-sector work over the whole chain
        pos, flops = [], []
        nsect = None
        for p in plans:
            j = [t for t, ix in enumerate(p.C.inds) if ix.id == shard_index_id and ix.plev == shard_index_plev]
            assert len(j) == 1, "the sharding index must be an uncontracted index of every step"
            pos.append(j[0])
            nsect = p.C.inds[j[0]].nsect
            flops.append(pair_flops(p))
        w = np.zeros(nsect)
        for p, j, f in zip(plans, pos, flops):
            np.add.at(w, p.C.blocks[:, j], f)
        self.owner = sector_assignment(w, world)
        self.sector_work = w
        self.masks = [(self.owner[p.C.blocks[:, j]] == rank).astype(np.uint8) for p, j in zip(plans, pos)]
        self.my_flops = float(sum(f[m.astype(bool)].sum() for f, m in zip(flops, self.masks)))
        self.total_flops = float(sum(f.sum() for f in flops))
        for p, m in zip(plans, self.masks):
            p.set_cblock_mask(m)

    def zero_unowned(self, out_tensor) -> None:
        """zero H*phi before the last step so that blocks owned by other ranks contribute exact zeros"""
        out_tensor.zero_()

    def allgather(self, flat) -> None:
        import torch.distributed as dist

        dist.all_reduce(flat, op=dist.ReduceOp.SUM)


def shard_chain(plans, world, rank) -> ChainShard:
    return ChainShard(plans, world, rank)
