// kernels_gemm.cu — grouped block-pair contraction kernels for sm_100a (FP64).
//
// Replaces the per-pair loop of the reference (loopContractedBlocks, itensor/itdata/qutil.h:244-371
// -> contract(), itensor/tensor/contract.cc:733-824 -> dgemm) by ONE launch per shape class over
// ALL block pairs of a contraction:
//
//   bsc_gemm_kernel<BM,BN,WM,WN>  tensor-pipe class: persistent CTAs pull (C block, tile) items from
//                                 a work queue; every item runs one K loop over ALL pairs of its C
//                                 block (so no beta / no atomics) with warp-level DMMA
//                                 (mma.sync.m8n8k4.f64; sm_100 has no FP64 tcgen05 kind and ptxas
//                                 lowers every larger f64 mma shape to DMMA.8x8x4 anyway).
//                                 Operand tiles are GATHERED straight from the strided N-index
//                                 blocks through separable offset tables — no permuted copies.
//   bsc_skinny_kernel             HBM-bound class (min(M,N) <= 8, e.g. the MPO steps of H_eff*phi):
//                                 one thread per long-side row, short operand staged in smem.
//   bsc_dot_kernel(+_finish)      M*N <= 4 (scalar products <V|q>): deterministic split-K.
//
// Complex arithmetic arrives here already folded into a real problem by the planner (plan.cc).
#include <cuda_runtime.h>
#include <stdint.h>

#include "tables.h"

namespace itb {

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// offset of linear index idx within an index group (extents fastest-first)
__device__ __forceinline__ int64_t grp_off(int idx, const int32_t* __restrict__ ext, const int64_t* __restrict__ str, int n) {
    int64_t o = 0;
#pragma unroll
    for (int d = 0; d < ITB_MAXG; ++d) {
        if (d < n) {
            if (d == n - 1) {
                o += (int64_t)idx * str[d];
            } else {
                const int e = ext[d];
                const int q = idx / e;
                o += (int64_t)(idx - q * e) * str[d];
                idx = q;
            }
        }
    }
    return o;
}

template <int BM_, int BN_, int WM_, int WN_>
struct GemmCfg {
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, BK = 16;
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
    static constexpr int NT = WARPS_M * WARPS_N * 32;
    static constexpr int LD_XF = 4;                                   // padding: ld == 4 (mod 16) -> conflict-free frags
    static constexpr int A_ELEMS = (BM * (BK + LD_XF) > BK * (BM + LD_XF)) ? BM * (BK + LD_XF) : BK * (BM + LD_XF);
    static constexpr int B_ELEMS = (BN * (BK + LD_XF) > BK * (BN + LD_XF)) ? BN * (BK + LD_XF) : BK * (BN + LD_XF);
    static constexpr int EA = BM * BK / NT, EB = BN * BK / NT;        // elements staged per thread per k-chunk
    static constexpr int KSLOTS = 4;                                  // ring of per-chunk k-offset tables
    static constexpr size_t SMEM = (size_t)(2 * (A_ELEMS + B_ELEMS)) * 8 + (size_t)(BM + BN + 2 * KSLOTS * BK) * 8 + 16;
    static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile/threads mismatch");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT) bsc_gemm_kernel(const ItbTile* __restrict__ tiles, int ntiles,
                                                           const ItbCBlk* __restrict__ cblks,
                                                           const ItbPair* __restrict__ pairs,
                                                           const double* __restrict__ A, const double* __restrict__ B,
                                                           double* __restrict__ C, int* __restrict__ counter) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, WM = Cfg::WM, WN = Cfg::WN, NT = Cfg::NT;
    constexpr int FM = WM / 8, FN = WN / 8, EA = Cfg::EA, EB = Cfg::EB, KS = Cfg::KSLOTS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);            // [2][A_ELEMS]
    double* Bs = As + 2 * Cfg::A_ELEMS;                          // [2][B_ELEMS]
    int64_t* offM_s = reinterpret_cast<int64_t*>(Bs + 2 * Cfg::B_ELEMS); // [BM]
    int64_t* offN_s = offM_s + BM;                               // [BN]
    int64_t* offKa_s = offN_s + BN;                              // [KS][BK]
    int64_t* offKb_s = offKa_s + KS * BK;                        // [KS][BK]
    int* item_s = reinterpret_cast<int*>(offKb_s + KS * BK);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int wm0 = (warp % Cfg::WARPS_M) * WM, wn0 = (warp / Cfg::WARPS_M) * WN;

    for (;;) {
        __syncthreads();
        if (tid == 0) *item_s = atomicAdd(counter, 1);
        __syncthreads();
        const int item = *item_s;
        if (item >= ntiles) break;
        const ItbTile tile = tiles[item];
        const ItbCBlk* cb = cblks + tile.cblk;
        const int M = cb->M, N = cb->N;
        const int m0 = tile.tm * BM, n0 = tile.tn * BN;

        double acc[FM][FN][2];
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
            for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int p = cb->pair_begin; p < cb->pair_end; ++p) {
            const ItbPair* pr = pairs + p;
            const int K = pr->K, flags = pr->flags;
            const bool cca = flags & ITB_PF_CCA, akf = flags & ITB_PF_A_KFAST, bkf = flags & ITB_PF_B_KFAST;
            const double* __restrict__ Ap = A + pr->a_off;
            const double* __restrict__ Bp = B + pr->b_off;
            const int nk = (K + BK - 1) / BK;
            // smem strides of the two possible tile layouts
            const int sAm = akf ? (BK + 4) : 1, sAk = akf ? 1 : (BM + 4);
            const int sBn = bkf ? (BK + 4) : 1, sBk = bkf ? 1 : (BN + 4);

            __syncthreads(); // previous pair's tables / buffers are dead
            for (int i = tid; i < BM + BN; i += NT) {
                if (i < BM) {
                    const int m = m0 + i;
                    offM_s[i] = (m < M) ? grp_off(m, pr->m_ext, pr->am_str, pr->m_n) : -1;
                } else {
                    const int n = n0 + i - BM;
                    offN_s[i - BM] = (n < N) ? grp_off(n, pr->n_ext, pr->bn_str, pr->n_n) : -1;
                }
            }
            // k-offset tables for chunks 0 and 1
            for (int i = tid; i < 4 * BK; i += NT) {
                const int c = i / (2 * BK), r = i % (2 * BK), kk = r % BK;
                const int k = c * BK + kk;
                if (r < BK) offKa_s[c * BK + kk] = (k < K) ? grp_off(k, pr->k_ext, pr->ak_str, pr->k_n) : -1;
                else offKb_s[c * BK + kk] = (k < K) ? grp_off(k, pr->k_ext, pr->bk_str, pr->k_n) : -1;
            }
            __syncthreads();

            double ra[EA], rb[EB];
            auto ldg_chunk = [&](int kc) {
                const int64_t* oka = offKa_s + (kc % KS) * BK;
                const int64_t* okb = offKb_s + (kc % KS) * BK;
#pragma unroll
                for (int e = 0; e < EA; ++e) {
                    const int idx = tid + e * NT;
                    const int mm = akf ? idx / BK : idx % BM, kk = akf ? idx % BK : idx / BM;
                    const int64_t om = offM_s[mm], ok = oka[kk];
                    double v = 0.0;
                    if ((om | ok) >= 0) {
                        int64_t off = om + ok;
                        if (cca) {
                            const int pq = (mm & 1) | ((kk & 1) << 1); // m0, k0 are even
                            if (pq == 3) off -= 2;
                            v = Ap[off];
                            if (pq == 2) v = -v;
                        } else {
                            v = Ap[off];
                        }
                    }
                    ra[e] = v;
                }
#pragma unroll
                for (int e = 0; e < EB; ++e) {
                    const int idx = tid + e * NT;
                    const int nn = bkf ? idx / BK : idx % BN, kk = bkf ? idx % BK : idx / BN;
                    const int64_t on = offN_s[nn], ok = okb[kk];
                    rb[e] = ((on | ok) >= 0) ? Bp[on + ok] : 0.0;
                }
            };
            auto sts_chunk = [&](int buf) {
                double* as = As + buf * Cfg::A_ELEMS;
                double* bs = Bs + buf * Cfg::B_ELEMS;
#pragma unroll
                for (int e = 0; e < EA; ++e) {
                    const int idx = tid + e * NT;
                    const int mm = akf ? idx / BK : idx % BM, kk = akf ? idx % BK : idx / BM;
                    as[mm * sAm + kk * sAk] = ra[e];
                }
#pragma unroll
                for (int e = 0; e < EB; ++e) {
                    const int idx = tid + e * NT;
                    const int nn = bkf ? idx / BK : idx % BN, kk = bkf ? idx % BK : idx / BN;
                    bs[nn * sBn + kk * sBk] = rb[e];
                }
            };

            ldg_chunk(0);
            sts_chunk(0);
            __syncthreads();
            for (int kc = 0; kc < nk; ++kc) {
                const int buf = kc & 1;
                if (kc + 1 < nk) ldg_chunk(kc + 1);
                // k-offsets for chunk kc+2 (read after the sync that ends iteration kc)
                if (kc + 2 < nk && tid < 2 * BK) {
                    const int kk = tid % BK, k = (kc + 2) * BK + kk;
                    if (tid < BK) offKa_s[((kc + 2) % KS) * BK + kk] = (k < K) ? grp_off(k, pr->k_ext, pr->ak_str, pr->k_n) : -1;
                    else offKb_s[((kc + 2) % KS) * BK + kk] = (k < K) ? grp_off(k, pr->k_ext, pr->bk_str, pr->k_n) : -1;
                }
                const double* as = As + buf * Cfg::A_ELEMS;
                const double* bs = Bs + buf * Cfg::B_ELEMS;
#pragma unroll
                for (int ks = 0; ks < BK / 4; ++ks) {
                    double fa[FM], fb[FN];
#pragma unroll
                    for (int i = 0; i < FM; ++i) fa[i] = as[(wm0 + i * 8 + g) * sAm + (ks * 4 + t4) * sAk];
#pragma unroll
                    for (int j = 0; j < FN; ++j) fb[j] = bs[(wn0 + j * 8 + g) * sBn + (ks * 4 + t4) * sBk];
#pragma unroll
                    for (int i = 0; i < FM; ++i)
#pragma unroll
                        for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
                }
                if (kc + 1 < nk) sts_chunk(buf ^ 1);
                __syncthreads();
            }
        }
        // ---- epilogue: C is written exactly once (all pairs of the block were accumulated above) ----
        double* __restrict__ Cp = C + cb->c_off;
        const int64_t cms = cb->c_ms, cns = cb->c_ns;
        const int nmask = cb->c_nmask, nshift = cb->c_nshift;
#pragma unroll
        for (int i = 0; i < FM; ++i) {
            const int m = m0 + wm0 + i * 8 + g;
#pragma unroll
            for (int j = 0; j < FN; ++j) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n = n0 + wn0 + j * 8 + 2 * t4 + h;
                    if (m < M && n < N) Cp[(int64_t)m * cms + (n & nmask) + (int64_t)(n >> nshift) * cns] = acc[i][j][h];
                }
            }
        }
    }
}

// ---- streaming kernel: short side S <= 8 -------------------------------------------------------------
constexpr int SK_NT = 256, SK_KC = 64, SK_S = 8;

__global__ void __launch_bounds__(SK_NT) bsc_skinny_kernel(const ItbSkinny* __restrict__ items, const ItbCBlk* __restrict__ cblks,
                                                           const ItbPair* __restrict__ pairs, const double* __restrict__ A,
                                                           const double* __restrict__ B, double* __restrict__ C) {
    __shared__ double Ss[SK_KC][SK_S];
    __shared__ int64_t offKl_s[SK_KC];
    const ItbSkinny it = items[blockIdx.x];
    const ItbCBlk* cb = cblks + it.cblk;
    const int tid = threadIdx.x;
    const bool lin = it.long_is_n; // long side is n (B is the long operand); else A is
    const int S = lin ? cb->M : cb->N;
    const int l = it.row0 + tid;
    const bool active = tid < it.rows;
    double acc[SK_S];
#pragma unroll
    for (int s = 0; s < SK_S; ++s) acc[s] = 0.0;

    for (int p = cb->pair_begin; p < cb->pair_end; ++p) {
        const ItbPair* pr = pairs + p;
        const int K = pr->K;
        const bool cca = pr->flags & ITB_PF_CCA;
        const double* __restrict__ Lp = lin ? B + pr->b_off : A + pr->a_off;
        const double* __restrict__ Sp = lin ? A + pr->a_off : B + pr->b_off;
        int64_t offL = 0;
        if (active) offL = lin ? grp_off(l, pr->n_ext, pr->bn_str, pr->n_n) : grp_off(l, pr->m_ext, pr->am_str, pr->m_n);
        for (int k0 = 0; k0 < K; k0 += SK_KC) {
            const int kn = min(SK_KC, K - k0);
            __syncthreads();
            for (int i = tid; i < SK_KC * SK_S; i += SK_NT) {
                const int kk = i / SK_S, s = i % SK_S;
                double v = 0.0;
                if (kk < kn && s < S) {
                    const int k = k0 + kk;
                    if (lin) { // short operand is A: element (m=s, k)
                        int64_t off = grp_off(s, pr->m_ext, pr->am_str, pr->m_n) + grp_off(k, pr->k_ext, pr->ak_str, pr->k_n);
                        if (cca) {
                            const int pq = (s & 1) | ((k & 1) << 1);
                            if (pq == 3) off -= 2;
                            v = Sp[off];
                            if (pq == 2) v = -v;
                        } else v = Sp[off];
                    } else { // short operand is B: element (k, n=s)
                        v = Sp[grp_off(k, pr->k_ext, pr->bk_str, pr->k_n) + grp_off(s, pr->n_ext, pr->bn_str, pr->n_n)];
                    }
                }
                Ss[kk][s] = v;
            }
            if (tid < kn) offKl_s[tid] = grp_off(k0 + tid, pr->k_ext, lin ? pr->bk_str : pr->ak_str, pr->k_n);
            __syncthreads();
            if (active) {
                const bool fix = cca && !lin; // long operand is the complex*complex A
#pragma unroll 4
                for (int kk = 0; kk < kn; ++kk) {
                    int64_t off = offL + offKl_s[kk];
                    double x;
                    if (fix) {
                        const int pq = (l & 1) | (((k0 + kk) & 1) << 1);
                        if (pq == 3) off -= 2;
                        x = Lp[off];
                        if (pq == 2) x = -x;
                    } else x = Lp[off];
                    const double4 s0 = *reinterpret_cast<const double4*>(&Ss[kk][0]);
                    const double4 s1 = *reinterpret_cast<const double4*>(&Ss[kk][4]);
                    acc[0] += x * s0.x; acc[1] += x * s0.y; acc[2] += x * s0.z; acc[3] += x * s0.w;
                    acc[4] += x * s1.x; acc[5] += x * s1.y; acc[6] += x * s1.z; acc[7] += x * s1.w;
                }
            }
        }
    }
    if (active) {
        double* __restrict__ Cp = C + cb->c_off;
#pragma unroll
        for (int s = 0; s < SK_S; ++s) {
            if (s < S) {
                const int m = lin ? s : l, n = lin ? l : s;
                Cp[(int64_t)m * cb->c_ms + (n & cb->c_nmask) + (int64_t)(n >> cb->c_nshift) * cb->c_ns] = acc[s];
            }
        }
    }
}

// ---- split-K reduction kernel: M*N <= 4 ---------------------------------------------------------------
constexpr int DOT_NT = 256;

__global__ void __launch_bounds__(DOT_NT) bsc_dot_kernel(const ItbDot* __restrict__ items, const ItbCBlk* __restrict__ cblks,
                                                         const ItbPair* __restrict__ pairs, const double* __restrict__ A,
                                                         const double* __restrict__ B, double* __restrict__ partial) {
    __shared__ double red[DOT_NT / 32][4];
    const ItbDot it = items[blockIdx.x];
    const ItbCBlk* cb = cblks + it.cblk;
    const ItbPair* pr = pairs + it.pair;
    const int M = cb->M, N = cb->N;
    const bool cca = pr->flags & ITB_PF_CCA;
    const double* __restrict__ Ap = A + pr->a_off;
    const double* __restrict__ Bp = B + pr->b_off;
    int64_t om[4], on[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        om[i] = (i < M) ? grp_off(i, pr->m_ext, pr->am_str, pr->m_n) : 0;
        on[i] = (i < N) ? grp_off(i, pr->n_ext, pr->bn_str, pr->n_n) : 0;
    }
    double acc[4] = {0.0, 0.0, 0.0, 0.0}; // acc[m + M*n]
    for (int kk = threadIdx.x; kk < it.klen; kk += DOT_NT) {
        const int k = it.k0 + kk;
        const int64_t oka = grp_off(k, pr->k_ext, pr->ak_str, pr->k_n);
        const int64_t okb = grp_off(k, pr->k_ext, pr->bk_str, pr->k_n);
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = 0.0; b[i] = 0.0;
            if (i < M) {
                int64_t off = om[i] + oka;
                if (cca) {
                    const int pq = (i & 1) | ((k & 1) << 1);
                    if (pq == 3) off -= 2;
                    a[i] = Ap[off];
                    if (pq == 2) a[i] = -a[i];
                } else a[i] = Ap[off];
            }
            if (i < N) b[i] = Bp[okb + on[i]];
        }
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int m = 0; m < 4; ++m)
                if (m < M && n < N && m + M * n < 4) acc[m + M * n] += a[m] * b[n];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) red[warp][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int w = 0; w < DOT_NT / 32; ++w) s += red[w][threadIdx.x];
        partial[(int64_t)it.slot * 4 + threadIdx.x] = s;
    }
}

__global__ void bsc_dot_finish_kernel(const ItbDotOut* __restrict__ outs, int nouts, const ItbCBlk* __restrict__ cblks,
                                      const double* __restrict__ partial, double* __restrict__ C) {
    const int o = blockIdx.x * (blockDim.x / 4) + threadIdx.x / 4, i = threadIdx.x & 3;
    if (o >= nouts) return;
    const ItbDotOut out = outs[o];
    const ItbCBlk* cb = cblks + out.cblk;
    const int M = cb->M, N = cb->N;
    if (i >= M * N) return;
    double s = 0.0;
    for (int q = 0; q < out.nslots; ++q) s += partial[(int64_t)(out.slot0 + q) * 4 + i];
    const int m = i % M, n = i / M;
    C[cb->c_off + (int64_t)m * cb->c_ms + (n & cb->c_nmask) + (int64_t)(n >> cb->c_nshift) * cb->c_ns] = s;
}

// ---- peak probes (no memory traffic) ------------------------------------------------------------------
__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}

// ---- launchers (called from api.cu) ---------------------------------------------------------------------
using CfgBig = GemmCfg<128, 128, 32, 32>;
using CfgMed = GemmCfg<64, 64, 32, 32>;
using CfgSmall = GemmCfg<32, 32, 16, 16>;

template <class Cfg>
static cudaError_t launch_gemm_cfg(const ItbTile* tiles, int ntiles, const ItbCBlk* cblks, const ItbPair* pairs,
                                   const double* A, const double* B, double* C, int* counter, int num_sms,
                                   cudaStream_t st) {
    static bool configured = false;
    static int occ = 1;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(bsc_gemm_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bsc_gemm_kernel<Cfg>, Cfg::NT, Cfg::SMEM);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
        configured = true;
    }
    int grid = num_sms * occ;
    if (grid > ntiles) grid = ntiles;
    bsc_gemm_kernel<Cfg><<<grid, Cfg::NT, Cfg::SMEM, st>>>(tiles, ntiles, cblks, pairs, A, B, C, counter);
    return cudaGetLastError();
}

cudaError_t launch_gemm(int cfg, const ItbTile* tiles, int ntiles, const ItbCBlk* cblks, const ItbPair* pairs,
                        const double* A, const double* B, double* C, int* counter, int num_sms, cudaStream_t st) {
    switch (cfg) {
        case 0: return launch_gemm_cfg<CfgBig>(tiles, ntiles, cblks, pairs, A, B, C, counter, num_sms, st);
        case 1: return launch_gemm_cfg<CfgMed>(tiles, ntiles, cblks, pairs, A, B, C, counter, num_sms, st);
        default: return launch_gemm_cfg<CfgSmall>(tiles, ntiles, cblks, pairs, A, B, C, counter, num_sms, st);
    }
}

cudaError_t launch_skinny(const ItbSkinny* items, int n, const ItbCBlk* cblks, const ItbPair* pairs, const double* A,
                          const double* B, double* C, cudaStream_t st) {
    bsc_skinny_kernel<<<n, SK_NT, 0, st>>>(items, cblks, pairs, A, B, C);
    return cudaGetLastError();
}

cudaError_t launch_dot(const ItbDot* items, int n, const ItbDotOut* outs, int nouts, const ItbCBlk* cblks,
                       const ItbPair* pairs, const double* A, const double* B, double* partial, double* C,
                       cudaStream_t st) {
    bsc_dot_kernel<<<n, DOT_NT, 0, st>>>(items, cblks, pairs, A, B, partial);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int per = 256 / 4;
    bsc_dot_finish_kernel<<<(nouts + per - 1) / per, 256, 0, st>>>(outs, nouts, cblks, partial, C);
    return cudaGetLastError();
}

cudaError_t launch_peak(int which, int iters, double* out, int num_sms, cudaStream_t st) {
    if (which == 1) peak_dfma_kernel<<<num_sms * 8, 256, 0, st>>>(out, iters);
    else peak_dmma_kernel<<<num_sms * 8, 256, 0, st>>>(out, iters);
    return cudaGetLastError();
}

} // namespace itb
