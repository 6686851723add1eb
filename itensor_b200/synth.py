"""Synthetic QN-structured inputs shaped like the reference's DMRG tensors.

All structure is generated on the host with NumPy from explicit seeds; element values come from a seeded
generator we own (uniform(-1,1)), never from the reference's unseeded quickran (detail/algs.h:87-111).

heff_chain() builds the operands of one H_eff*phi = LocalOp::product (itensor/mps/localop.h:324-365):
    phi(l, s1, s2, r) * L(l+, k0, l') * W1(k0+, s1+, s1', k1) * W2(k1+, s2+, s2', k2) * R(r+, k2+, r')
with Sz-like U(1) sectors: site index d=2 sectors (+1,-1) of size 1, MPO link sectors (0,-2,+2) of
sizes (3,1,1) (Heisenberg AutoMPO, probe in SURVEY §8a), MPS link sectors on an even-spaced ladder.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from ._lib import ITB_C64, ITB_F64
from .tensor import BlockStruct, Index, flux_blocks


def gaussian_sectors(m: int, nsect: int, sigma: float = 1.25, tilt: float = 0.15) -> List[int]:
    """Sector sizes summing to ~m with the bell shape measured on Heisenberg chains
    (m=400: 12,83,158,113,29,5 — SURVEY §8a)."""
    x = np.arange(nsect) - (nsect - 1) / 2.0 + tilt
    w = np.exp(-(x**2) / (2 * sigma**2))
    s = np.maximum(1, np.rint(m * w / w.sum())).astype(int)
    return [int(v) for v in s]


def equal_sectors(m: int, nsect: int) -> List[int]:
    return [m // nsect] * nsect


def link_index(idn: int, sizes: Sequence[int], dirn: int, q0: int = 0) -> Index:
    n = len(sizes)
    qns = tuple((q0 + 2 * (i - n // 2),) for i in range(n))
    return Index(idn, tuple(int(s) for s in sizes), qns, dirn, (1,))


def site_index(idn: int, dirn: int = 1, d: int = 2) -> Index:
    if d == 2:
        qns = ((1,), (-1,))
    elif d == 3:
        qns = ((2,), (0,), (-2,))
    else:
        raise ValueError(d)
    return Index(idn, (1,) * d, qns, dirn, (1,))


def mpo_link_index(idn: int, dirn: int, sizes=(3, 1, 1)) -> Index:
    return Index(idn, tuple(sizes), ((0,), (-2,), (2,)), dirn, (1,))


def random_values(struct: BlockStruct, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n = struct.nelems
    if struct.is_complex:
        return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex128)
    return rng.uniform(-1, 1, n)


def heff_chain(left_sizes: Sequence[int], right_sizes: Sequence[int] | None = None, d: int = 2,
               k_sizes=(3, 1, 1), dtype: int = ITB_F64, phi_order: str = "l s1 s2 r"):
    """Structures of (phi, L, W1, W2, R) for one effective-Hamiltonian product. All have flux 0 and
    every flux-allowed block (what QDense(IndexSet,QN) allocates, itdata/qdense.cc:108-115)."""
    right_sizes = list(left_sizes) if right_sizes is None else list(right_sizes)
    l = link_index(1, left_sizes, -1)       # phi: incoming left link
    r = link_index(2, right_sizes, +1)      # phi: outgoing right link
    s1, s2 = site_index(3, +1, d), site_index(4, +1, d)
    k0, k1, k2 = mpo_link_index(5, +1, k_sizes), mpo_link_index(6, +1, k_sizes), mpo_link_index(7, +1, k_sizes)
    names = {"l": l, "s1": s1, "s2": s2, "r": r}
    phi_inds = [names[t] for t in phi_order.split()]
    zero = (0,)

    def full(inds):
        return BlockStruct(inds, flux_blocks(inds, zero), dtype)

    phi = full(phi_inds)
    # L(l+, k0, l'): environments carry the bra link primed
    L = full([l.dag(), k0, l.prime()])
    W1 = full([k0.dag(), s1.dag(), s1.prime(), k1])
    W2 = full([k1.dag(), s2.dag(), s2.prime(), k2])
    R = full([r.dag(), k2.dag(), r.prime()])
    return phi, L, W1, W2, R


def random_qn_pair(rng: np.random.Generator, order_a: int, order_b: int, ncont: int, max_sect: int = 4,
                   max_size: int = 5, dtype_a: int = ITB_F64, dtype_b: int = ITB_F64, drop: float = 0.25,
                   flux_a: int = 0, flux_b: int = 0):
    """Random pair of QN tensors sharing `ncont` indices (at random positions, random arrows), with a
    random subset of the flux-allowed blocks stored (block-deficient tensors, itensor_test.cc:2734-2822)."""
    nid = [100]

    def rand_index(dirn):
        ns = int(rng.integers(1, max_sect + 1))
        sizes = tuple(int(v) for v in rng.integers(1, max_size + 1, ns))
        qns = tuple((int(rng.integers(-1, 2)),) for _ in range(ns))
        nid[0] += 1
        return Index(nid[0], sizes, qns, dirn, (1,))

    shared = [rand_index(int(rng.choice([-1, 1]))) for _ in range(ncont)]
    a_inds = shared + [rand_index(int(rng.choice([-1, 1]))) for _ in range(order_a - ncont)]
    b_inds = [s.dag() for s in shared] + [rand_index(int(rng.choice([-1, 1]))) for _ in range(order_b - ncont)]
    a_inds = [a_inds[i] for i in rng.permutation(order_a)]
    b_inds = [b_inds[i] for i in rng.permutation(order_b)]

    def struct(inds, flux, dtype):
        bl = flux_blocks(inds, (flux,))
        if len(bl) > 1 and drop > 0:
            keep = rng.uniform(size=len(bl)) >= drop
            if not keep.any():
                keep[int(rng.integers(len(bl)))] = True
            bl = bl[keep]
        return BlockStruct(inds, bl, dtype)

    return struct(a_inds, flux_a, dtype_a), struct(b_inds, flux_b, dtype_b)
