// plan.h — host-side planner objects (integer bookkeeping only; no CUDA types).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/itb200.h"
#include "tables.h"

namespace itb {

struct TensorStruct { // owning copy of an itb_tensor_desc
    int order = 0;
    int dtype = ITB_F64;
    std::vector<int32_t> nsect;
    std::vector<int64_t> sect;       // concatenated
    std::vector<int64_t> sect_start; // prefix into sect, order+1 entries
    int64_t nblocks = 0;
    std::vector<int32_t> blocks;
    std::vector<int64_t> offsets;
    int64_t nelems = 0;
    int64_t ext(int j, int s) const { return sect[sect_start[j] + s]; }
    const int32_t* block(int64_t b) const { return blocks.data() + b * order; }
};

// Tile configurations of the DMMA kernel (kernels_gemm.cu instantiates the same list).
enum { ITB_CFG_BIG = 0, ITB_CFG_MED = 1, ITB_CFG_SMALL = 2, ITB_NCFG = 3 };
static const int kTileM[ITB_NCFG] = {128, 64, 32};
static const int kTileN[ITB_NCFG] = {128, 64, 32};

struct DeviceTables; // defined in api.cu (device copies of the vectors below)

} // namespace itb

struct itb_contract_plan {
    itb::TensorStruct A, B, C;
    std::vector<int32_t> labA, labB, labC;
    std::vector<int64_t> triples; // (iA,iB,iC) per pair, reference enumeration order
    double flops = 0;
    double class_flops[5] = {0, 0, 0, 0, 0};
    int64_t cb_first = 0, cb_last = -1; // execution range of C blocks (sharding); -1 => all
    std::vector<uint8_t> cb_mask;       // optional per-C-block selection (empty => all)
    // optional row-slice of one index of C (multi-GPU sharding inside QN sectors): C blocks whose coordinate on index
    // slice_index is sector s execute only the rows [slice_lo[s], slice_hi[s]) of that index (empty range: block skipped)
    int32_t slice_index = -1;
    std::vector<int64_t> slice_lo, slice_hi;

    // device-format tables (built by build_tables(), rebuilt when the range changes)
    bool tables_built = false;
    int64_t runs = 0;        // executions so far (itb_contract_run): drives the tiered planning of the streaming class
    bool rg_deferred = false; // the current tables route row-group-eligible C blocks to the C-stationary kernels
    std::vector<ItbPair> pairs;
    std::vector<ItbCBlk> cblks;
    std::vector<ItbTile> tiles;       // every tile class in one list; CTA b of the persistent grid owns [cta_begin[b], cta_begin[b+1])
    std::vector<ItbQItem> qitems;     // device form of `tiles` (flattened per-item records, same order)
    std::vector<int32_t> cta_begin;   // kNumSMs+1 entries (stream-K partition: equal modelled cycles per CTA)
    std::vector<ItbSplitOut> splits;  // split-K tiles to be reduced from the workspace
    int64_t ws_slots = 0;
    // measured refinement of the static partition (itb_contract_plan_refine): tile_scale[t] multiplies the modelled
    // cost of tile t (planner order); item_tile / item_cost describe the items of the current partition
    std::vector<double> tile_scale;
    std::vector<int32_t> item_tile;
    std::vector<double> item_cost;
    std::vector<ItbSkinny> skinny;    // generic streaming items
    std::vector<ItbSkinny> skinny_q4; // small-K fast path, short side <= 4
    std::vector<ItbSkinny> skinny_q8; // small-K fast path, short side <= 8
    std::vector<ItbRowGroup> rgroups; // row-group streaming class (see tables.h)
    std::vector<ItbRgIn> rg_in;
    std::vector<int64_t> rg_out;      // output slot bases (REAL-element offsets into C)
    std::vector<ItbRgW> rg_w;
    std::vector<ItbRgItem> rg_items;
    std::vector<ItbDot> dots;
    std::vector<ItbDotOut> dot_outs;
    int64_t ndot_slots = 0;
    std::vector<int64_t> zero_ranges; // (offset,len) REAL ranges of C blocks to zero: none today (every
                                      // C block has >=1 pair) but kept for range-restricted runs
    int64_t table_bytes = 0;
    itb::DeviceTables* dev = nullptr; // owned by api.cu
    void* dev_ctx = nullptr;
};

struct itb_permute_plan {
    itb::TensorStruct S, D;
    std::vector<int32_t> perm;
    std::vector<ItbPermBlk> blks_copy;  // blocks whose src/dst fastest dims coincide
    std::vector<ItbPermBlk> blks_tiled; // blocks needing a shared-memory transpose
    int64_t items_copy = 0, items_tiled = 0;
    std::vector<ItbPermChunk> chunk_items; // one per copy-path work item
    std::vector<ItbPermTile> tile_items;   // one per transposing-path work item
    bool need_zero = false; // dst has blocks no src block maps to (memset path)
    bool zero_in_items = false; // ... and they are zero-fill items of the tile kernel instead (skipped when accumulating)
    std::vector<int64_t> zero_ranges; // (element offset, element count) of those blocks, merged when adjacent
    int64_t bytes = 0;
    itb::DeviceTables* dev = nullptr;
    void* dev_ctx = nullptr;
};

namespace itb {
void set_error(const std::string& msg);
int parse_desc(const itb_tensor_desc* d, TensorStruct& out, const char* what);
int build_contract_plan(itb_contract_plan& P);
int build_contract_tables(itb_contract_plan& P);
// counts one execution; true when the tables must be rebuilt first (tiered planning: a plan that keeps coming back gets
// its row-group tables now)
bool plan_note_run(itb_contract_plan& P);
int build_permute_plan(itb_permute_plan& P);

constexpr int kPermCopyChunk = 4096; // elements per work item, copy-like path
constexpr int kPermTile = 32;        // tile edge, transposing path
constexpr int kDotChunk = 8192;      // k-range per split-K item
constexpr int kSkinnyRows = 1024;    // long-side rows per streaming item (4 per thread)
constexpr int kSkinnyQRows = 8192;   // rows per item of the small-K fast path (8 batches of 4 rows/thread)
constexpr int kSkinnyQMaxK = 64;     // sum of K over the pairs of a C block / max pairs for the fast path
constexpr int kSkinnyQMaxPairs = 8;
constexpr int64_t kSkinnyMinElems = 0;
constexpr int64_t kStreamMinTotal = 262144;  // when a contraction also has tile work and its skinny C blocks hold fewer elements than this in
                                             // total, they ride the tile queue (a few thousand cycles each, spread over all CTAs by the stream-K
                                             // partition) instead of paying latency-bound launches of their own
constexpr int kNumSMs = 148;         // B200: persistent grid size the split-K heuristic balances for
constexpr int kSkinnyMax = 8;        // short side <= this -> streaming kernel
constexpr int kRowGroupRows = 4096;  // long-side rows per work item of the row-group streaming kernel
constexpr int kDotMaxMN = 4;        // M*N <= this -> reduction kernel (scalar results, re/im pairs)
} // namespace itb
