// kernels_permute.cu — batched block permute / permuting accumulate for sm_100a (HBM-bound).
//
// Replaces the generic strided transform() loop (itensor/tensor/ten_impl.h:107-160) behind
// permuteDense (itensor/itdata/dense.cc:417-428), permuteQDense (itensor/itdata/qdense.cc:847-881)
// and the PlusEQ add() (itensor/itdata/qdense.cc:515-549): one launch moves ALL blocks.
//   perm_copy_kernel : src and dst share their fastest (fused) dim -> both sides coalesced directly
//   perm_tile_kernel : fastest dims differ -> 32x32 tile through padded shared memory so that the
//                      global read runs along src-fastest and the global write along dst-fastest
// Elements are moved as 8-byte (real) or 16-byte (complex, one 128-bit access) units.
// Block dims arrive fused/canonical from the planner (plan.cc build_permute_plan).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>

#include "tables.h"

namespace itb {

struct cplx_t { double re, im; };

template <bool CS, bool CD> struct ElemOp;
template <> struct ElemOp<false, false> { // real -> real
    using S = double; using D = double;
    static __device__ __forceinline__ D apply(S v, D old, double ar, double ai, bool accum) { return accum ? fma(ar, v, old) : ar * v; }
};
template <> struct ElemOp<true, true> { // complex -> complex
    using S = double2; using D = double2;
    static __device__ __forceinline__ D apply(S v, D old, double ar, double ai, bool accum) {
        double2 r;
        r.x = ar * v.x - ai * v.y;
        r.y = ar * v.y + ai * v.x;
        if (accum) { r.x += old.x; r.y += old.y; }
        return r;
    }
};
template <> struct ElemOp<false, true> { // real -> complex (promotion)
    using S = double; using D = double2;
    static __device__ __forceinline__ D apply(S v, D old, double ar, double ai, bool accum) {
        double2 r;
        r.x = ar * v; r.y = ai * v;
        if (accum) { r.x += old.x; r.y += old.y; }
        return r;
    }
};

__device__ __forceinline__ int find_block(const ItbPermBlk* __restrict__ blks, int nblk, int64_t item) {
    int lo = 0, hi = nblk - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (blks[mid].item_begin <= item) lo = mid; else hi = mid - 1;
    }
    return lo;
}

constexpr int PC_NT = 256, PC_CHUNK = 4096;

template <bool CS, bool CD>
__global__ void __launch_bounds__(PC_NT) perm_copy_kernel(const ItbPermBlk* __restrict__ blks, const ItbPermChunk* __restrict__ items,
                                                          const void* __restrict__ src_, void* __restrict__ dst_,
                                                          double ar, double ai, int accum) {
    using Op = ElemOp<CS, CD>;
    using S = typename Op::S; using D = typename Op::D;
    __shared__ ItbPermBlk sb;
    const ItbPermChunk it = items[blockIdx.x];
    if (threadIdx.x < (int)(sizeof(ItbPermBlk) / 8)) reinterpret_cast<int64_t*>(&sb)[threadIdx.x] = reinterpret_cast<const int64_t*>(blks + it.blk)[threadIdx.x];
    __syncthreads();
    const S* __restrict__ src = reinterpret_cast<const S*>(src_) + sb.s_off;
    D* __restrict__ dst = reinterpret_cast<D*>(dst_) + sb.d_off;
    const int64_t e0 = it.e0;
    const int n = sb.n;
    constexpr int PER = PC_CHUNK / PC_NT, UB = 8; // UB independent loads in flight per thread
#pragma unroll 1
    for (int i0 = 0; i0 < PER; i0 += UB) {
        S v[UB];
        int64_t eo[UB]; // destination offset (or -1 past the end)
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int64_t e = e0 + threadIdx.x + (int64_t)(i0 + u) * PC_NT;
            eo[u] = -1;
            v[u] = S();
            if (e < sb.nelem) {
                int64_t so = 0, dof = 0, rem = e;
#pragma unroll
                for (int d = 0; d < ITB_MAXG; ++d) {
                    if (d < n) {
                        if (d == n - 1) { so += rem * sb.sstr[d]; dof += rem * sb.dstr[d]; }
                        else {
                            const int64_t q = rem / sb.ext[d], i = rem - q * sb.ext[d];
                            so += i * sb.sstr[d]; dof += i * sb.dstr[d];
                            rem = q;
                        }
                    }
                }
                v[u] = src[so];
                eo[u] = dof;
            }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            if (eo[u] >= 0) {
                D old = D();
                if (accum) old = dst[eo[u]];
                dst[eo[u]] = Op::apply(v[u], old, ar, ai, accum);
            }
        }
    }
}

// PT x PT tile per CTA (64 for 8-byte elements, 32 for 16-byte complex), 256 threads, every thread keeps
// PT*PT/256 independent loads in flight before the tile is turned in shared memory.
constexpr int PTN = 256;

template <bool CS, bool CD, int PT>
__global__ void __launch_bounds__(PTN) perm_tile_kernel(const ItbPermTile* __restrict__ items, const void* __restrict__ src_,
                                                        void* __restrict__ dst_, double ar, double ai, int accum) {
    using Op = ElemOp<CS, CD>;
    using S = typename Op::S; using D = typename Op::D;
    constexpr int ROWS = PTN / PT, PER = PT / ROWS; // rows covered per pass, passes
    __shared__ S tile[PT][PT + 1];
    const ItbPermTile it = items[blockIdx.x]; // one 48-byte record per CTA: no search, no index arithmetic
    const S* __restrict__ src = reinterpret_cast<const S*>(src_) + it.s_base;
    D* __restrict__ dst = reinterpret_cast<D*>(dst_) + it.d_base;
    if (it.nT < 0) { // zero-fill item: n0 contiguous destination elements no source block maps to (permuteQDense fill-in)
        if (!accum) // (an accumulating pass leaves blocks without a source untouched)
            for (int e = threadIdx.x; e < it.n0; e += PTN) dst[e] = D();
        return;
    }
    const int tx = threadIdx.x % PT, ty = threadIdx.x / PT;
    // read: tx runs along the src-fastest dim (stride 1 in src); all PER loads issued before any store
    {
        S v[PER];
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const int i0 = ty + r * ROWS;
            v[r] = S();
            if (tx < it.nT && i0 < it.n0) v[r] = src[(int64_t)tx + (int64_t)i0 * it.ss0];
        }
#pragma unroll
        for (int r = 0; r < PER; ++r) tile[ty + r * ROWS][tx] = v[r];
    }
    __syncthreads();
    // write: tx runs along the dst-fastest dim (stride 1 in dst)
    if (accum) {
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const int iT = ty + r * ROWS;
            if (iT < it.nT && tx < it.n0) {
                const int64_t o = (int64_t)tx + (int64_t)iT * it.dsT;
                dst[o] = Op::apply(tile[tx][iT], dst[o], ar, ai, true);
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const int iT = ty + r * ROWS;
            if (iT < it.nT && tx < it.n0) dst[(int64_t)tx + (int64_t)iT * it.dsT] = Op::apply(tile[tx][iT], D(), ar, ai, false);
        }
    }
}

// ---- persistent, double-buffered variant of the transposing path -------------------------------------------------
// Same per-tile records, but a CTA walks a strided sequence of tiles and the global READ side goes through cp.async
// (LDGSTS: global -> shared without a register round trip, 8-byte units for real data, one 128-bit unit per complex
// element), so the loads of tile i+1 are in flight while tile i is turned and written. Against the one-tile-per-CTA
// kernel above this removes the load/store phase alternation inside a CTA (no loads in flight while it stores), the STS
// half of the shared-memory instruction stream, and the wave-quantisation tail of a ~4000-CTA launch.
__device__ __forceinline__ void cp_async_elem(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_elem(double2* smem_dst, const double2* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
template <bool CS, bool CD, int PT>
__global__ void __launch_bounds__(PTN) perm_tile_pipe_kernel(const ItbPermTile* __restrict__ items, int nitems, const void* __restrict__ src_,
                                                             void* __restrict__ dst_, double ar, double ai, int accum) {
    using Op = ElemOp<CS, CD>;
    using S = typename Op::S; using D = typename Op::D;
    constexpr int ROWS = PTN / PT, PER = PT / ROWS;
    extern __shared__ __align__(16) unsigned char perm_smem[];
    S (*tile)[PT][PT + 1] = reinterpret_cast<S (*)[PT][PT + 1]>(perm_smem); // [2][PT][PT+1]
    const int tx = threadIdx.x % PT, ty = threadIdx.x / PT;
    auto issue = [&](int item, int buf) {
        const ItbPermTile it = items[item];
        if (it.nT >= 0 && tx < it.nT) {
            const S* __restrict__ src = reinterpret_cast<const S*>(src_) + it.s_base + tx;
#pragma unroll
            for (int r = 0; r < PER; ++r) {
                const int i0 = ty + r * ROWS;
                if (i0 < it.n0) cp_async_elem(&tile[buf][i0][tx], src + (int64_t)i0 * it.ss0);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int item = blockIdx.x, buf = 0;
    if (item < nitems) issue(item, 0);
    for (; item < nitems; item += gridDim.x, buf ^= 1) {
        const int next = item + gridDim.x;
        if (next < nitems) { issue(next, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const ItbPermTile it = items[item];
        D* __restrict__ dst = reinterpret_cast<D*>(dst_) + it.d_base;
        if (it.nT < 0) { // zero-fill item (permuteQDense fill-in); an accumulating pass leaves such blocks untouched
            if (!accum)
                for (int e = threadIdx.x; e < it.n0; e += PTN) dst[e] = D();
        } else if (accum) {
            D old[PER];
#pragma unroll
            for (int r = 0; r < PER; ++r) {
                const int iT = ty + r * ROWS;
                old[r] = D();
                if (iT < it.nT && tx < it.n0) old[r] = dst[(int64_t)tx + (int64_t)iT * it.dsT];
            }
#pragma unroll
            for (int r = 0; r < PER; ++r) {
                const int iT = ty + r * ROWS;
                if (iT < it.nT && tx < it.n0) dst[(int64_t)tx + (int64_t)iT * it.dsT] = Op::apply(tile[buf][tx][iT], old[r], ar, ai, true);
            }
        } else {
#pragma unroll
            for (int r = 0; r < PER; ++r) {
                const int iT = ty + r * ROWS;
                if (iT < it.nT && tx < it.n0) dst[(int64_t)tx + (int64_t)iT * it.dsT] = Op::apply(tile[buf][tx][iT], D(), ar, ai, false);
            }
        }
        __syncthreads(); // the loads issued in the next iteration overwrite this buffer
    }
}

template <bool CS, bool CD>
static cudaError_t launch_t(const ItbPermBlk* bc, const ItbPermChunk* chunks, int64_t items_c, const ItbPermTile* tiles, int64_t items_t,
                            const void* src, void* dst, double ar, double ai, int accum, cudaStream_t st, int* launches) {
    if (items_c > 0) {
        perm_copy_kernel<CS, CD><<<(unsigned)items_c, PC_NT, 0, st>>>(bc, chunks, src, dst, ar, ai, accum);
        ++*launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (items_t > 0) {
        // ITB_PERM_PIPE=0 selects the one-tile-per-CTA kernel (measurement / fallback switch); ITB_PERM_CTAS: CTAs per SM
        static int pipe = -1, ctas = 2, sms = 148; // 2 CTAs/SM measured best (0.758 of the copy peak; 3: 0.68, 4: 0.75, 6: 0.757)
        if (pipe < 0) {
            const char* e = getenv("ITB_PERM_PIPE"); pipe = e ? atoi(e) : 1;
            if (const char* c = getenv("ITB_PERM_CTAS")) ctas = atoi(c) > 0 ? atoi(c) : 2;
            int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
        constexpr int PT = CS ? 32 : 64;
        using S = typename ElemOp<CS, CD>::S;
        constexpr size_t smem = 2 * sizeof(S) * PT * (PT + 1);
        if (pipe) {
            static bool configured = false;
            if (!configured) {
                cudaError_t e = cudaFuncSetAttribute(perm_tile_pipe_kernel<CS, CD, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                configured = true;
            }
            const unsigned grid = (unsigned)std::min<int64_t>(items_t, (int64_t)sms * ctas);
            perm_tile_pipe_kernel<CS, CD, PT><<<grid, PTN, smem, st>>>(tiles, (int)items_t, src, dst, ar, ai, accum);
        } else {
            perm_tile_kernel<CS, CD, PT><<<(unsigned)items_t, PTN, 0, st>>>(tiles, src, dst, ar, ai, accum);
        }
        ++*launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_permute(int src_cplx, int dst_cplx, const ItbPermBlk* bc, const ItbPermChunk* chunks, int64_t items_c,
                           const ItbPermTile* tiles, int64_t items_t, const void* src, void* dst, double ar, double ai, int accum,
                           cudaStream_t st, int* launches) {
    if (src_cplx && dst_cplx) return launch_t<true, true>(bc, chunks, items_c, tiles, items_t, src, dst, ar, ai, accum, st, launches);
    if (!src_cplx && dst_cplx) return launch_t<false, true>(bc, chunks, items_c, tiles, items_t, src, dst, ar, ai, accum, st, launches);
    return launch_t<false, false>(bc, chunks, items_c, tiles, items_t, src, dst, ar, ai, accum, st, launches);
}

} // namespace itb
