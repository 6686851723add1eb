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
__global__ void __launch_bounds__(PC_NT) perm_copy_kernel(const ItbPermBlk* __restrict__ blks, int nblk,
                                                          const void* __restrict__ src_, void* __restrict__ dst_,
                                                          double ar, double ai, int accum) {
    using Op = ElemOp<CS, CD>;
    using S = typename Op::S; using D = typename Op::D;
    __shared__ ItbPermBlk sb;
    const int64_t item = blockIdx.x;
    if (threadIdx.x == 0) sb = blks[find_block(blks, nblk, item)];
    __syncthreads();
    const S* __restrict__ src = reinterpret_cast<const S*>(src_) + sb.s_off;
    D* __restrict__ dst = reinterpret_cast<D*>(dst_) + sb.d_off;
    const int64_t e0 = (item - sb.item_begin) * PC_CHUNK;
    const int n = sb.n;
#pragma unroll 4
    for (int i = 0; i < PC_CHUNK / PC_NT; ++i) {
        const int64_t e = e0 + threadIdx.x + i * PC_NT;
        if (e >= sb.nelem) break;
        int64_t so = 0, rem = e;
#pragma unroll
        for (int d = 0; d < ITB_MAXG; ++d) {
            if (d < n) {
                if (d == n - 1) so += rem * sb.sstr[d];
                else { const int64_t q = rem / sb.ext[d]; so += (rem - q * sb.ext[d]) * sb.sstr[d]; rem = q; }
            }
        }
        const S v = src[so];
        D old = D();
        if (accum) old = dst[e];
        dst[e] = Op::apply(v, old, ar, ai, accum);
    }
}

constexpr int PT = 32, PT_ROWS = 8;

template <bool CS, bool CD>
__global__ void __launch_bounds__(PT* PT_ROWS) perm_tile_kernel(const ItbPermBlk* __restrict__ blks, int nblk,
                                                                const void* __restrict__ src_, void* __restrict__ dst_,
                                                                double ar, double ai, int accum) {
    using Op = ElemOp<CS, CD>;
    using S = typename Op::S; using D = typename Op::D;
    __shared__ S tile[PT][PT + 1];
    __shared__ ItbPermBlk sb;
    __shared__ int64_t base_s, base_d;
    __shared__ int t0_s, tT_s;
    const int64_t item = blockIdx.x;
    if (threadIdx.x == 0) {
        sb = blks[find_block(blks, nblk, item)];
        int64_t r = item - sb.item_begin;
        t0_s = (int)(r % sb.tiles0); r /= sb.tiles0;
        tT_s = (int)(r % sb.tilesT); r /= sb.tilesT;
        int64_t bs = 0, bd = 0;
        for (int d = 1; d < sb.n; ++d) {
            if (d == sb.tdim) continue;
            const int64_t i = r % sb.ext[d]; r /= sb.ext[d];
            bs += i * sb.sstr[d]; bd += i * sb.dstr[d];
        }
        base_s = bs; base_d = bd;
    }
    __syncthreads();
    const S* __restrict__ src = reinterpret_cast<const S*>(src_) + sb.s_off + base_s;
    D* __restrict__ dst = reinterpret_cast<D*>(dst_) + sb.d_off + base_d;
    const int tx = threadIdx.x % PT, ty = threadIdx.x / PT;
    const int e0 = sb.ext[0], eT = sb.ext[sb.tdim];
    const int64_t ss0 = sb.sstr[0], dsT = sb.dstr[sb.tdim];
    // read: tx runs along the src-fastest dim (stride 1 in src)
    {
        const int iT = tT_s * PT + tx;
#pragma unroll
        for (int r = 0; r < PT / PT_ROWS; ++r) {
            const int i0 = t0_s * PT + ty + r * PT_ROWS;
            if (iT < eT && i0 < e0) tile[ty + r * PT_ROWS][tx] = src[(int64_t)iT + (int64_t)i0 * ss0];
        }
    }
    __syncthreads();
    // write: tx runs along the dst-fastest dim (stride 1 in dst)
    {
        const int i0 = t0_s * PT + tx;
#pragma unroll
        for (int r = 0; r < PT / PT_ROWS; ++r) {
            const int iT = tT_s * PT + ty + r * PT_ROWS;
            if (iT < eT && i0 < e0) {
                const int64_t o = (int64_t)i0 + (int64_t)iT * dsT;
                D old = D();
                if (accum) old = dst[o];
                dst[o] = Op::apply(tile[tx][ty + r * PT_ROWS], old, ar, ai, accum);
            }
        }
    }
}

template <bool CS, bool CD>
static cudaError_t launch_t(const ItbPermBlk* bc, int nbc, int64_t items_c, const ItbPermBlk* bt, int nbt, int64_t items_t,
                            const void* src, void* dst, double ar, double ai, int accum, cudaStream_t st, int* launches) {
    if (items_c > 0) {
        perm_copy_kernel<CS, CD><<<(unsigned)items_c, PC_NT, 0, st>>>(bc, nbc, src, dst, ar, ai, accum);
        ++*launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (items_t > 0) {
        perm_tile_kernel<CS, CD><<<(unsigned)items_t, PT * PT_ROWS, 0, st>>>(bt, nbt, src, dst, ar, ai, accum);
        ++*launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_permute(int src_cplx, int dst_cplx, const ItbPermBlk* bc, int nbc, int64_t items_c, const ItbPermBlk* bt,
                           int nbt, int64_t items_t, const void* src, void* dst, double ar, double ai, int accum,
                           cudaStream_t st, int* launches) {
    if (src_cplx && dst_cplx) return launch_t<true, true>(bc, nbc, items_c, bt, nbt, items_t, src, dst, ar, ai, accum, st, launches);
    if (!src_cplx && dst_cplx) return launch_t<false, true>(bc, nbc, items_c, bt, nbt, items_t, src, dst, ar, ai, accum, st, launches);
    return launch_t<false, false>(bc, nbc, items_c, bt, nbt, items_t, src, dst, ar, ai, accum, st, launches);
}

} // namespace itb
