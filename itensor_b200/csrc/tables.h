// tables.h — compact POD tables shared by the host planner (plan.cc) and the sm_100a kernels.
//
// The host turns IndexSet/QN block bookkeeping into these tables (the role of
// getContractedOffsets + CProps::compute in the reference, itensor/itdata/qutil.h:93-242 and
// itensor/tensor/contract.cc:240-548); the device only ever sees flat data pointers plus
// {offsets, extents, strides}. All offsets/strides are in units of REAL doubles: complex
// tensors are addressed through their interleaved (re,im) storage.
#pragma once
#include <stdint.h>

#define ITB_MAXG 6 // fused dims per index group (M / K / N); more -> ITB_ERR_UNSUPPORTED

// ---- contraction -------------------------------------------------------------------------------
// A block pair contributes  C(m,n) += sum_k A[a_off + offM(m) + offK_a(k)] * B[b_off + offK_b(k) + offN(n)]
// where offX(i) decomposes i over the group's extents (fastest first) and dots with strides.
// Complex operands are folded into this REAL problem by the planner (pseudo-dim of extent 2,
// see plan.cc "complex folding"); ITB_PF_CCA marks the A operand of a complex*complex pair:
// element (m',k') with p=m'&1, q=k'&1 is  (q&!p ? -1 : 1) * A[off - 2*(p&q)].
enum { ITB_PF_CCA = 1, ITB_PF_A_KFAST = 2, ITB_PF_B_KFAST = 4 };

struct ItbPair {
    int64_t a_off, b_off;
    int64_t am_str[ITB_MAXG];
    int64_t bn_str[ITB_MAXG];
    int64_t ak_str[ITB_MAXG];
    int64_t bk_str[ITB_MAXG];
    int32_t m_ext[ITB_MAXG]; // fusion pattern can differ between pairs of one C block
    int32_t n_ext[ITB_MAXG];
    int32_t k_ext[ITB_MAXG];
    int32_t m_n, n_n, k_n;
    int32_t K;     // prod(k_ext) (real-expanded)
    int32_t flags; // ITB_PF_*
    int32_t pad_;
};

struct ItbCBlk {
    int64_t c_off;
    int64_t c_ms;     // C address = c_off + m*c_ms + (n & c_nmask) + (n >> c_nshift)*c_ns
    int64_t c_ns;
    int32_t c_nmask;  // 1 for real(A)*cplx(B) (interleaved re/im along n'), else 0
    int32_t c_nshift; // 1 for real*cplx, else 0
    int32_t M, N;     // real-expanded block dims
    int32_t pair_begin, pair_end; // pairs of this C block, in reference enumeration order
    int64_t ksum;     // sum of K over the pairs (work estimate)
};

struct ItbTile { // work item of the persistent DMMA tile kernel
    int32_t cblk, m0, n0; // C block and origin of the tile inside it (real-expanded rows / columns)
    int32_t cfg;         // tile configuration (ITB_CFG_*: 128x128 / 64x64 / 32x32)
    int32_t chunk_begin; // range of BK-chunks of the C block's concatenated K loop (split-K)
    int32_t chunk_end;
    int32_t ws_slot;     // -1: write C directly; else partial tile goes to workspace slot ws_slot
    int32_t split;       // index of this tile's ItbSplitOut record (arrival counter, first slot, piece count); -1 if uncut
};
// Device record of one queue item: the tile plus everything the kernel needs to know about its C block and the first
// ITB_QPAIRS block pairs (K, flags), flattened by the host so that fetching an item is ONE 160-byte read instead of a
// chain of dependent loads (tile -> C block -> pairs).
#define ITB_QPAIRS 8
struct ItbQItem {
    ItbTile tile;
    ItbCBlk cb;
    int32_t pK[ITB_QPAIRS];
    int32_t pflags[ITB_QPAIRS];
    int32_t pad_[2];
};
struct ItbSplitOut { // split-K tile: C tile = sum of workspace slots [ws_slot0, ws_slot0+nsplit) in order
    int32_t cblk, m0, n0, cfg, ws_slot0, nsplit, pad_[2];
};
#define ITB_BK 16              // K-chunk of the tile kernel
#define ITB_WS_TILE (128 * 128) // doubles per workspace slot
// Peer copies of the result (multi-GPU, see itb_contract_run_mirrored): every C element the tile class writes is also stored at
// the same offset of up to ITB_MAX_MIRRORS other buffers (peer GPUs' copies of C mapped over NVLink); delta = element offset of
// the mirror's base relative to C.
#define ITB_MAX_MIRRORS 7
struct ItbMirrors { int32_t n; int32_t pad_; long long delta[ITB_MAX_MIRRORS]; };

struct ItbSkinny { // work item of the streaming kernel: rows [row0,row0+rows) of the long side
    int32_t cblk, row0, rows, long_is_n; // long_is_n: 1 -> threads run over n, short side is m
};
// ---- row-group streaming class -----------------------------------------------------------------------------
// HBM-bound contractions of a large operand A with tiny operator blocks B (the MPO steps of H_eff*phi): all A blocks
// that share their uncontracted ("long side") block coordinates feed all C blocks with those coordinates. One row
// group = those blocks; for every long-side row l
//     y[o] = sum_j x[j] * W[j][o],   x[j] = A[in_base_j + sum_d i_d*in_str_j[d]],   C[out_base_o + l] = y[o]
// where (i_0..i_{nL-1}) decomposes l over ext[]. Input slot j = (A block, k), output slot o = (C block, n); W is
// assembled in shared memory from B through w entries. Every element of A is read once and every element of C
// written once per row group (the C-stationary kernels re-read an A block once per C block it feeds).
#define ITB_RG_MAXIN 32
#define ITB_RG_MAXOUT 16
#define ITB_RG_MAXL 3
struct ItbRowGroup {
    int32_t nin, nout, nL;
    int32_t ostr;                // stride (in reals) between consecutive long-side rows of an output slot: 1, or 2 when the
                                 // slots are the (re, im) components of a complex C whose rows are complex elements
    int32_t ext[ITB_RG_MAXL];
    int32_t in_begin, out_begin; // into the slot tables
    int32_t w_begin, w_count;    // into the W entry table
    int32_t pad2_;
    int64_t L;                   // rows
};
struct ItbRgIn { int64_t base; int64_t str[ITB_RG_MAXL]; }; // REAL-element offsets into A
struct ItbRgW { int32_t j, o; int64_t b_off; };              // W[j][o] = B[b_off]; b_off < 0: W[j][o] = -B[~b_off] (complex weights)
struct ItbRgItem { int32_t group, row0, rows, pad_; };

struct ItbDot { // work item of the split-K reduction kernel
    int32_t cblk, pair, k0, klen;
    int32_t slot; // partial-sum slot (row in the partial buffer, 16 doubles each)
    int32_t pad_[3];
};
struct ItbDotOut { // per tiny C block: partial slots [slot0, slot0+nslots) summed in order
    int32_t cblk, slot0, nslots, pad_;
};

// ---- permute -----------------------------------------------------------------------------------
// dst[d_off + sum_j i_j*dstr_j] (op)= alpha * src[s_off + sum_j i_j*sstr_j], dims fused/canonical,
// listed in DST order (dim 0 = fastest in dst, dstr[0] == cs). tiled: the smem-transpose path is
// used over (dim 0 = dst-fastest, dim tdim = src-fastest); otherwise src and dst share dim 0.
struct ItbPermBlk {
    int64_t s_off, d_off;
    int64_t sstr[ITB_MAXG];
    int64_t dstr[ITB_MAXG];
    int32_t ext[ITB_MAXG];
    int32_t n;
    int32_t tdim;       // index (in this list) of the src-fastest dim; 0 => copy-like
    int64_t item_begin; // first work item (tile / chunk) of this block in the launch
    int64_t nelem;
    int32_t tiles0, tilesT; // tile counts along dim 0 and dim tdim (tiled path)
    int32_t cs;             // 2 if elements are complex pairs moved as units, else 1
    int32_t promote;        // 1: src real -> dst complex
};
// per-work-item records (one global load per CTA instead of a block search + index arithmetic)
struct ItbPermTile { // transposing path: one PT x PT tile
    int64_t s_base, d_base; // element offsets of the tile's origin in src / dst
    int64_t ss0, dsT;       // src stride of the dst-fastest dim, dst stride of the src-fastest dim
    int32_t n0, nT;         // valid extent of the tile along dst-fastest / src-fastest dim (<= PT);
                            // nT < 0: zero-fill item, n0 (<= 4096) contiguous elements at d_base
};
struct ItbPermChunk { // copy-like path: PC_CHUNK consecutive dst elements of one block
    int32_t blk, pad_;
    int64_t e0;
};
