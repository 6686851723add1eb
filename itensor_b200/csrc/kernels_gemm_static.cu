// kernels_gemm_static.cu — the tile kernel for purely STATIC schedules (stream-K partition, plan.cc): every CTA walks its
// own contiguous range of the item list, cta_begin[b]..cta_begin[b+1]; no work queue, no item ring. Same warp-specialised
// producer / consumer loops as kernels_gemm.cu (which adds the dynamic queue); both roles read the tile records straight
// from global memory. On the bench workload this kernel + the fitted static partition is the fastest combination
// measured (same-box A/B, profiles/r03_tile_schedule_ab.txt), so it is the default; ITB_SCHED=guided selects the queue.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "tables.h"

namespace itb {
namespace r1 {

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma884_if(double& d0, double& d1, double a, double b, int on) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %4, 0;\n\t"
        "@p mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n\t}"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b), "r"(on));
}
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

// ---- persistent warp-specialised DMMA tile kernel ---------------------------------------------------------
// One launch serves every tile class: items carry their configuration (128x128 / 64x64 / 32x32) and a
// range of K-chunks (split-K for C blocks with few tiles but long K loops).
//   warps 0-15 CONSUMERS  4 (m) x 4 (n) warp grid; per K-chunk: wait full[stage] -> 4 x (LDS fragments,
//                         DMMA.8x8x4) -> arrive empty[stage]. They never touch global operands or tables.
//   warps 16-19 PRODUCERS  gather the A/B chunk straight from the strided N-index blocks with 8-byte
//                         cp.async (zero-fill past the edges) into a 4-stage ring; completion is tracked
//                         by cp.async.mbarrier.arrive on full[stage].
// Producers and consumers walk the same deterministic (tile, pair, chunk) sequence, so there is no
// CTA-wide barrier anywhere in the main loop: DMMA stretches of the two consumer warps of an SMSP
// interleave freely and the epilogue of one tile overlaps the loads of the next.
constexpr int G_NCONS = 512, G_NPROD = 128, G_NT = G_NCONS + G_NPROD;
constexpr int G_STAGES = 4, G_BK = ITB_BK, G_PAD = 4, G_MAXT = 128;
constexpr int G_KT = 1024; // k offsets per shared table fill (per operand)
constexpr int G_STAGE_ELEMS = G_MAXT * (G_BK + G_PAD); // per operand per stage (covers both layouts)
constexpr size_t G_SMEM = (size_t)(2 * G_STAGES * G_STAGE_ELEMS) * 8 + (size_t)(2 * G_MAXT) * 8 + (size_t)(2 * G_KT) * 4 + 2 * G_STAGES * 8 + 16;

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cp_async(uint64_t* bar) { // arrives once this thread's prior cp.asyncs landed
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, int parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void producer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(G_NPROD) : "memory"); }

struct PipeState { // position in the stage ring; identical sequence on both sides
    int stage = 0, phase = 0;
    __device__ __forceinline__ void advance() {
        if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
    }
};

// inner K loop of one block pair with the shared-memory layout of both operands fixed at compile time
// (all fragment addresses become immediates off one base register per operand)
// EDGE: the warp owns fewer than FM x FN valid 8x8 fragments (tile overhanging the C block): fragments that lie
// entirely outside are neither loaded nor multiplied. Rows/columns of a partially valid fragment that fall
// outside only ever see their own (unwritten) C rows/columns, so the operands need no zero fill along m and n.
template <int BM, int BN, bool AKF, bool BKF, bool EDGE>
__device__ __forceinline__ void consume_pair(double (&acc)[BM / 32][BN / 32][2], int nchunks, int sgn, int wm0, int wn0, int g, int t4,
                                             int lane, int fmv, int fnv, const double* As, const double* Bs, uint64_t* full,
                                             uint64_t* empty, PipeState& ps) {
    constexpr int BK = G_BK, FM = BM / 32, FN = BN / 32;
    constexpr int sAm = AKF ? (BK + G_PAD) : 1, sAk = AKF ? 1 : (BM + G_PAD);
    constexpr int sBn = BKF ? (BK + G_PAD) : 1, sBk = BKF ? 1 : (BN + G_PAD);
    const int a_base = (wm0 + g) * sAm + t4 * sAk, b_base = (wn0 + g) * sBn + t4 * sBk;
    for (int kc = 0; kc < nchunks; ++kc) {
        mbar_wait(&full[ps.stage], ps.phase);
        const double* as = As + ps.stage * G_STAGE_ELEMS + a_base;
        const double* bs = Bs + ps.stage * G_STAGE_ELEMS + b_base;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double fa[FM], fb[FN]; // (fragments outside an edge tile are loaded anyway: the addresses stay inside the stage)
#pragma unroll
            for (int i = 0; i < FM; ++i) {
                const double v = as[i * 8 * sAm + ks * 4 * sAk];
                fa[i] = __hiloint2double(__double2hiint(v) ^ sgn, __double2loint(v));
            }
#pragma unroll
            for (int j = 0; j < FN; ++j) fb[j] = bs[j * 8 * sBn + ks * 4 * sBk];
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j) {
                    if (EDGE) dmma884_if(acc[i][j][0], acc[i][j][1], fa[i], fb[j], (i < fmv) & (j < fnv));
                    else dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
                }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[ps.stage]);
        ps.advance();
    }
}

template <int BM, int BN, bool MIRROR>
__device__ __forceinline__ void consume_tile(const ItbTile& tile, const ItbCBlk* __restrict__ cb, const ItbPair* __restrict__ pairs,
                                             double* __restrict__ C, double* __restrict__ ws, const double* As, const double* Bs,
                                             uint64_t* full, uint64_t* empty, PipeState& ps, int dbg_nocompute, const ItbMirrors& mir) {
    constexpr int BK = G_BK;
    constexpr int WM = BM / 4, WN = BN / 4, FM = WM / 8, FN = WN / 8;
    static_assert(FM >= 1 && FN >= 1, "tile too small for a 4x4 warp grid");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    // warp w runs on SMSP w&3: (wm,wn) = ((w ^ (w>>2)) & 3, w>>2) puts one warp of every warp-row and of every
    // warp-column on each SMSP, so fragment skipping on edge tiles unloads all four tensor pipes evenly
    const int wm0 = ((warp ^ (warp >> 2)) & 3) * WM, wn0 = (warp >> 2) * WN;
    const int M = cb->M, N = cb->N;
    const int m0 = tile.m0, n0 = tile.n0;
    const int fmv = min(FM, max(0, (M - m0 - wm0 + 7) >> 3)), fnv = min(FN, max(0, (N - n0 - wn0 + 7) >> 3));
    const bool edge = fmv < FM || fnv < FN;

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    int gchunk = 0;
    for (int p = cb->pair_begin; p < cb->pair_end; ++p) {
        const ItbPair* pr = pairs + p;
        const int nk = (pr->K + BK - 1) / BK;
        const int c0 = max(tile.chunk_begin - gchunk, 0), c1 = min(tile.chunk_end - gchunk, nk);
        gchunk += nk;
        if (c0 >= c1) continue;
        const int flags = pr->flags;
        // sign of the A' = [[Ar,-Ai],[Ai,Ar]] expansion, applied when the fragment is read (row even, col odd)
        const int sgn = ((flags & ITB_PF_CCA) && !(g & 1) && (t4 & 1)) ? (int)0x80000000 : 0;
        if (fmv == 0 || fnv == 0 || dbg_nocompute) { // nothing of this warp's sub-tile is inside the C block: keep the ring moving
            // (dbg_nocompute: producer-rate measurement, tools/tile_calib.py --nocompute; results are garbage)
            for (int kc = c0; kc < c1; ++kc) {
                mbar_wait(&full[ps.stage], ps.phase);
                if (lane == 0) mbar_arrive(&empty[ps.stage]);
                ps.advance();
            }
            continue;
        }
#define ITB_CONSUME(AKF, BKF)                                                                                                     \
    do {                                                                                                                          \
        if (edge) consume_pair<BM, BN, AKF, BKF, true>(acc, c1 - c0, sgn, wm0, wn0, g, t4, lane, fmv, fnv, As, Bs, full, empty, ps); \
        else consume_pair<BM, BN, AKF, BKF, false>(acc, c1 - c0, sgn, wm0, wn0, g, t4, lane, fmv, fnv, As, Bs, full, empty, ps);     \
    } while (0)
        switch (flags & (ITB_PF_A_KFAST | ITB_PF_B_KFAST)) {
            case 0: ITB_CONSUME(false, false); break;
            case ITB_PF_A_KFAST: ITB_CONSUME(true, false); break;
            case ITB_PF_B_KFAST: ITB_CONSUME(false, true); break;
            default: ITB_CONSUME(true, true); break;
        }
#undef ITB_CONSUME
    }
    // ---- epilogue: each C element is written exactly once (or one partial per split) ----------------------
    if (tile.ws_slot < 0) {
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
                    if (m < M && n < N) {
                        double* dst = Cp + ((int64_t)m * cms + (n & nmask) + (int64_t)(n >> nshift) * cns);
                        *dst = acc[i][j][h];
                        if (MIRROR) { // the same element into every peer's copy of C (stores over NVLink, posted)
                            for (int q = 0; q < mir.n; ++q) dst[mir.delta[q]] = acc[i][j][h];
                        }
                    }
                }
            }
        }
    } else {
        double* __restrict__ W = ws + (int64_t)tile.ws_slot * ITB_WS_TILE;
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
            for (int j = 0; j < FN; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) W[(wm0 + i * 8 + g) + BM * (wn0 + j * 8 + 2 * t4 + h)] = acc[i][j][h];
    }
}

// Producer side of one block pair with both operand layouts fixed at compile time. Everything that does not change
// along K lives in registers for the whole pair: the row (m / n) offsets of the elements this thread gathers — one
// offset when the operand's fastest index is m/n (thread = one row, walks k), ROWS/8 offsets when it is k
// (thread = one k column, walks rows). The k offsets of up to G_KT consecutive k sit in a shared table that is
// refilled every G_KT/BK chunks, so a chunk costs one broadcast LDS + one 64-bit add + one LDGSTS per element and
// no barrier among the producers.
template <int ROWS, bool KF>
struct OperandGather {
    static constexpr int BK = G_BK, NP = G_NPROD;
    static constexpr int E = ROWS * BK / NP;                     // elements per thread per chunk
    static constexpr int NR = KF ? E : 1;                        // row offsets held in registers
    const double* base;                                          // block base
    int roff[NR];                                                // row offset inside the block (< 2^31, planner-checked); -1: row outside the C block
    int soff;                                                    // shared-memory element offset of this thread's first element in a stage
    int k0;                                                      // first k column (within a chunk) of this thread
    bool odd_row;                                                // parity of the row(s): complex*complex fix-up
    __device__ __forceinline__ void init(int pt, const double* b, const int64_t* off_s) {
        base = b;
        if (KF) { // k column fixed, rows pt/BK + e*(NP/BK)
            k0 = pt % BK;
            odd_row = (pt / BK) & 1; // NP/BK is even
            soff = (pt / BK) * (BK + G_PAD) + k0;
#pragma unroll
            for (int e = 0; e < NR; ++e) roff[e] = (int)off_s[pt / BK + e * (NP / BK)];
        } else { // row fixed, k columns pt/ROWS + e*(NP/ROWS)
            const int r = pt % ROWS;
            k0 = pt / ROWS;
            odd_row = r & 1;
            roff[0] = (int)off_s[r];
            soff = r + k0 * (ROWS + G_PAD);
        }
    }
    // issue this thread's share of one chunk; ktab: k offsets of the chunk's BK columns (0 past K: the address stays
    // valid and the copy zero-fills; kleft = K - first k of the chunk). All table reads happen BEFORE the first copy
    // is issued: the cp.async statements are ordered memory operations for the compiler, so interleaving them with
    // the table loads would serialise one shared-memory round trip per element.
    template <bool CCA>
    __device__ __forceinline__ void issue(double* stage, const int* ktab, int kleft) const {
        if (KF) {
            const int ok = ktab[k0];
            const bool v = k0 < kleft;
            const double* col = base + ok - ((CCA && v && (k0 & 1) && odd_row) ? 2 : 0);
#pragma unroll
            for (int e = 0; e < NR; ++e)
                if (roff[e] >= 0) cp_async8(stage + soff + e * (NP / BK) * (BK + G_PAD), col + roff[e], v);
        } else {
            constexpr int STEP = NP / ROWS; // k columns between two elements of this thread (1, 2 or 4)
            int ok[E];
            if (STEP == 1) { // 16 consecutive table entries: four 128-bit loads
#pragma unroll
                for (int q = 0; q < E / 4; ++q) {
                    const int4 t = reinterpret_cast<const int4*>(ktab)[q];
                    ok[4 * q] = t.x; ok[4 * q + 1] = t.y; ok[4 * q + 2] = t.z; ok[4 * q + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) ok[e] = ktab[k0 + e * STEP];
            }
            if (roff[0] >= 0) {
                const double* row = base + roff[0];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int kk = k0 + e * STEP;
                    const bool v = kk < kleft;
                    const int adj = (CCA && v && (kk & 1) && odd_row) ? 2 : 0;
                    cp_async8(stage + soff + e * STEP * (ROWS + G_PAD), row + ok[e] - adj, v);
                }
            }
        }
    }
};

template <int BM, int BN, bool AKF, bool BKF>
__device__ __forceinline__ void produce_pair(const ItbPair* __restrict__ pr, int c0, int c1, const double* __restrict__ Ap,
                                             const double* __restrict__ Bp, double* As, double* Bs, const int64_t* offM_s,
                                             const int64_t* offN_s, int* ktabA, int* ktabB, uint64_t* full, uint64_t* empty,
                                             PipeState& ps) {
    constexpr int BK = G_BK, NP = G_NPROD;
    const int pt = threadIdx.x - G_NCONS;
    const int K = pr->K;
    const bool cca = pr->flags & ITB_PF_CCA;
    OperandGather<BM, AKF> ga;
    OperandGather<BN, BKF> gb;
    ga.init(pt, Ap, offM_s);
    gb.init(pt, Bp, offN_s);
    for (int kb = c0; kb < c1; kb += G_KT / BK) {
        const int ke = min(c1, kb + G_KT / BK);
        if (kb > c0) producer_sync(); // everyone is done with the previous table
        for (int i = pt; i < (ke - kb) * BK; i += NP) {
            const int k = kb * BK + i;
            ktabA[i] = (k < K) ? (int)grp_off(k, pr->k_ext, pr->ak_str, pr->k_n) : 0; // < 2^31: planner-checked block size
            ktabB[i] = (k < K) ? (int)grp_off(k, pr->k_ext, pr->bk_str, pr->k_n) : 0;
        }
        producer_sync();
        for (int kc = kb; kc < ke; ++kc) {
            mbar_wait(&empty[ps.stage], ps.phase ^ 1);
            double* as = As + ps.stage * G_STAGE_ELEMS;
            double* bs = Bs + ps.stage * G_STAGE_ELEMS;
            const int* ta = ktabA + (kc - kb) * BK;
            const int* tb = ktabB + (kc - kb) * BK;
            const int kleft = K - kc * BK;
            if (cca) ga.template issue<true>(as, ta, kleft);
            else ga.template issue<false>(as, ta, kleft);
            gb.template issue<false>(bs, tb, kleft);
            mbar_arrive_cp_async(&full[ps.stage]);
            ps.advance();
        }
    }
}

template <int BM, int BN>
__device__ __forceinline__ void produce_tile(const ItbTile& tile, const ItbCBlk* __restrict__ cb, const ItbPair* __restrict__ pairs,
                                             const double* __restrict__ A, const double* __restrict__ B, double* As, double* Bs,
                                             int64_t* offM_s, int64_t* offN_s, int* ktabA, int* ktabB, uint64_t* full,
                                             uint64_t* empty, PipeState& ps) {
    constexpr int BK = G_BK, NP = G_NPROD;
    const int pt = threadIdx.x - G_NCONS; // 0..G_NPROD-1
    const int M = cb->M, N = cb->N;
    const int m0 = tile.m0, n0 = tile.n0;
    int gchunk = 0;
    for (int p = cb->pair_begin; p < cb->pair_end; ++p) {
        const ItbPair* pr = pairs + p;
        const int nk = (pr->K + BK - 1) / BK;
        const int c0 = max(tile.chunk_begin - gchunk, 0), c1 = min(tile.chunk_end - gchunk, nk);
        gchunk += nk;
        if (c0 >= c1) continue;
        const int flags = pr->flags;
        producer_sync(); // every producer is done reading the previous pair's tables
        for (int i = pt; i < BM + BN; i += NP) {
            if (i < BM) {
                const int m = m0 + i;
                offM_s[i] = (m < M) ? grp_off(m, pr->m_ext, pr->am_str, pr->m_n) : -1;
            } else {
                const int n = n0 + i - BM;
                offN_s[i - BM] = (n < N) ? grp_off(n, pr->n_ext, pr->bn_str, pr->n_n) : -1;
            }
        }
        producer_sync();
        const double* __restrict__ Ap = A + pr->a_off;
        const double* __restrict__ Bp = B + pr->b_off;
        switch (flags & (ITB_PF_A_KFAST | ITB_PF_B_KFAST)) {
            case 0: produce_pair<BM, BN, false, false>(pr, c0, c1, Ap, Bp, As, Bs, offM_s, offN_s, ktabA, ktabB, full, empty, ps); break;
            case ITB_PF_A_KFAST: produce_pair<BM, BN, true, false>(pr, c0, c1, Ap, Bp, As, Bs, offM_s, offN_s, ktabA, ktabB, full, empty, ps); break;
            case ITB_PF_B_KFAST: produce_pair<BM, BN, false, true>(pr, c0, c1, Ap, Bp, As, Bs, offM_s, offN_s, ktabA, ktabB, full, empty, ps); break;
            default: produce_pair<BM, BN, true, true>(pr, c0, c1, Ap, Bp, As, Bs, offM_s, offN_s, ktabA, ktabB, full, empty, ps); break;
        }
    }
}

template <bool MIRROR>
__global__ void __launch_bounds__(G_NT, 1) bsc_gemm_static_kernel(const ItbTile* __restrict__ tiles, const int32_t* __restrict__ cta_begin,
                                                            const ItbCBlk* __restrict__ cblks, const ItbPair* __restrict__ pairs,
                                                            const double* __restrict__ A, const double* __restrict__ B,
                                                            double* __restrict__ C, double* __restrict__ ws,
                                                            long long* __restrict__ cta_cycles, int dbg_nocompute, const ItbMirrors mir) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const long long t_begin = cta_cycles ? clock64() : 0;
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + G_STAGES * G_STAGE_ELEMS;
    int64_t* offM_s = reinterpret_cast<int64_t*>(Bs + G_STAGES * G_STAGE_ELEMS);
    int64_t* offN_s = offM_s + G_MAXT;
    int* offKa_s = reinterpret_cast<int*>(offN_s + G_MAXT); // 16-byte aligned: read as int4 by the producers
    int* offKb_s = offKa_s + G_KT;
    uint64_t* full = reinterpret_cast<uint64_t*>(offKb_s + G_KT);
    uint64_t* empty = full + G_STAGES;
    if (threadIdx.x == 0) {
        for (int s = 0; s < G_STAGES; ++s) {
            mbar_init(&full[s], G_NPROD);     // one cp.async-completion arrive per producer thread
            mbar_init(&empty[s], G_NCONS / 32); // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool producer = threadIdx.x >= G_NCONS;
    // register re-balancing (warpgroup granular): the kernel launches with 96 regs/thread (640 threads);
    // the producer warpgroup shrinks, the four consumer warpgroups grow (4*112 + 56 per SMSP fits 16K).
    // (setmaxnreg variants measured slower or spilling: see DESIGN.md) if (producer) setmaxnreg.dec 56
    // else setmaxnreg.inc 112
    PipeState ps;
    // the host planner hands every CTA a contiguous range of (tile, K-chunk range) items of equal modelled cost
    // (stream-K partition, plan.cc); both roles walk it in the same order
    const int item_end = cta_begin[blockIdx.x + 1];
    for (int item = cta_begin[blockIdx.x]; item < item_end; ++item) {
        const ItbTile tile = tiles[item];
        const ItbCBlk* cb = cblks + tile.cblk;
        if (producer) {
            if (tile.cfg == 0) produce_tile<128, 128>(tile, cb, pairs, A, B, As, Bs, offM_s, offN_s, offKa_s, offKb_s, full, empty, ps);
            else if (tile.cfg == 1) produce_tile<64, 64>(tile, cb, pairs, A, B, As, Bs, offM_s, offN_s, offKa_s, offKb_s, full, empty, ps);
            else produce_tile<32, 32>(tile, cb, pairs, A, B, As, Bs, offM_s, offN_s, offKa_s, offKb_s, full, empty, ps);
        } else {
            if (tile.cfg == 0) consume_tile<128, 128, MIRROR>(tile, cb, pairs, C, ws, As, Bs, full, empty, ps, dbg_nocompute, mir);
            else if (tile.cfg == 1) consume_tile<64, 64, MIRROR>(tile, cb, pairs, C, ws, As, Bs, full, empty, ps, dbg_nocompute, mir);
            else consume_tile<32, 32, MIRROR>(tile, cb, pairs, C, ws, As, Bs, full, empty, ps, dbg_nocompute, mir);
        }
    }
    if (cta_cycles && threadIdx.x == 0) cta_cycles[blockIdx.x] = clock64() - t_begin; // schedule calibration (profile mode)
}


} // namespace r1

cudaError_t launch_splitk_reduce(const ItbSplitOut* souts, int nsouts, const ItbCBlk* cblks, const double* ws, double* C, const ItbMirrors* mir,
                                 cudaStream_t st); // kernels_gemm.cu

// mir != nullptr (multi-GPU): every element of C is also stored into the peers' copies by the epilogue / the split-K reduction
cudaError_t launch_gemm_static(const ItbTile* items, const int32_t* cta_begin, int grid, const ItbSplitOut* souts, int nsouts,
                               const ItbCBlk* cblks, const ItbPair* pairs, const double* A, const double* B, double* C, double* ws,
                               long long* cta_cycles, const ItbMirrors* mir, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(r1::bsc_gemm_static_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r1::G_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(r1::bsc_gemm_static_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r1::G_SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    ItbMirrors none;
    none.n = 0;
    if (mir && mir->n > 0) r1::bsc_gemm_static_kernel<true><<<grid, r1::G_NT, r1::G_SMEM, st>>>(items, cta_begin, cblks, pairs, A, B, C, ws, cta_cycles, 0, *mir);
    else r1::bsc_gemm_static_kernel<false><<<grid, r1::G_NT, r1::G_SMEM, st>>>(items, cta_begin, cblks, pairs, A, B, C, ws, cta_cycles, 0, none);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || nsouts == 0) return e;
    return launch_splitk_reduce(souts, nsouts, cblks, ws, C, mir, st);
}

} // namespace itb
