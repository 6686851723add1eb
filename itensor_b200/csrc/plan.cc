// plan.cc — host-side integer planning for the block-sparse contraction and permute paths.
//
// What the reference does per call on the CPU, restated for a table-driven GPU launch:
//   * label matching + result index order : computeLabels (itensor/tensor/contract.h:155-202),
//                                           contractIS   (itensor/indexset_impl.h:77-129)
//   * block-pair enumeration + C offsets  : getContractedOffsets (itensor/itdata/qutil.h:93-242)
//   * per-pair GEMM shape analysis        : CProps::compute (itensor/tensor/contract.cc:240-548)
//
// Differences by design (B200-first, not a port):
//   * pairs are found through a hash on the contracted block coordinates (O(nA+nB+npairs))
//     instead of the O(nA*nB) scan, but are emitted in the SAME order (A outer, B inner);
//   * no operand is ever permuted/materialised: each pair carries separable offset tables
//     (extent/stride lists for its M, K, N index groups) and the kernel gathers tiles directly;
//   * all pairs of one C block form one K-loop owned by one CTA tile (no beta, no atomics);
//   * complex operands are folded into a REAL problem (see fold_complex) so one FP64 kernel
//     serves the four real/complex pairings (reference: 4 DGEMMs, tensor/gemm.cc:165-230).
#ifdef ITB_PLAN_PROFILE
#include <chrono>
#endif
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <unordered_map>

namespace itb {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error_cstr() { return g_err.c_str(); }

int parse_desc(const itb_tensor_desc* d, TensorStruct& t, const char* what) {
    if (!d) { set_error(std::string(what) + ": null descriptor"); return ITB_ERR_INVALID; }
    if (d->order < 0 || d->order > ITB_MAX_ORDER) {
        set_error(std::string(what) + ": order out of range");
        return ITB_ERR_INVALID;
    }
    if (d->dtype != ITB_F64 && d->dtype != ITB_C64) {
        set_error(std::string(what) + ": bad dtype");
        return ITB_ERR_INVALID;
    }
    t.order = d->order;
    t.dtype = d->dtype;
    t.nsect.assign(d->nsect, d->nsect + d->order);
    t.sect_start.assign(d->order + 1, 0);
    for (int j = 0; j < d->order; ++j) {
        if (t.nsect[j] <= 0) { set_error(std::string(what) + ": index with no sectors"); return ITB_ERR_INVALID; }
        t.sect_start[j + 1] = t.sect_start[j] + t.nsect[j];
    }
    t.sect.assign(d->sect, d->sect + t.sect_start[d->order]);
    for (auto s : t.sect)
        if (s <= 0) { set_error(std::string(what) + ": non-positive sector size"); return ITB_ERR_INVALID; }
    t.nblocks = d->nblocks;
    if (t.nblocks < 0) { set_error(std::string(what) + ": negative block count"); return ITB_ERR_INVALID; }
    t.blocks.assign(d->blocks, d->blocks + t.nblocks * d->order);
    t.offsets.assign(d->offsets, d->offsets + t.nblocks);
    t.nelems = d->nelems;
    for (int64_t b = 0; b < t.nblocks; ++b) {
        int64_t sz = 1;
        for (int j = 0; j < t.order; ++j) {
            int32_t c = t.block(b)[j];
            if (c < 0 || c >= t.nsect[j]) { set_error(std::string(what) + ": block coordinate out of range"); return ITB_ERR_INVALID; }
            sz *= t.ext(j, c);
        }
        if (t.offsets[b] < 0 || t.offsets[b] + sz > t.nelems) {
            set_error(std::string(what) + ": block exceeds storage");
            return ITB_ERR_INVALID;
        }
    }
    return ITB_OK;
}

// reference Block ordering (itensor/itdata/qdense.cc:60-65): lexicographic on the REVERSED coordinates
static bool block_less(const int32_t* a, const int32_t* b, int r) {
    for (int j = r - 1; j >= 0; --j) {
        if (a[j] != b[j]) return a[j] < b[j];
    }
    return false;
}

struct VecHash {
    size_t operator()(const std::vector<int32_t>& v) const {
        uint64_t h = 1469598103934665603ull;
        for (auto x : v) { h ^= (uint32_t)x + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); }
        return (size_t)h;
    }
};

#ifdef ITB_PLAN_PROFILE
double g_plan_phase[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // harness builds only (build/planprof)
#define PLAN_T0() auto plan_t_ = std::chrono::steady_clock::now()
#define PLAN_PHASE(i) do { auto n_ = std::chrono::steady_clock::now(); g_plan_phase[i] += std::chrono::duration<double>(n_ - plan_t_).count(); plan_t_ = n_; } while (0)
#else
#define PLAN_T0() do {} while (0)
#define PLAN_PHASE(i) do {} while (0)
#endif
int build_contract_plan(itb_contract_plan& P) {
    PLAN_T0();
    const TensorStruct &A = P.A, &B = P.B;
    TensorStruct& C = P.C;
    const int rA = A.order, rB = B.order;
    // ---- label matching (computeLabels: first match wins) -------------------------------------
    std::vector<int> AtoB(rA, -1), BtoA(rB, -1);
    for (int i = 0; i < rA; ++i)
        for (int j = 0; j < rB; ++j)
            if (P.labA[i] == P.labB[j] && BtoA[j] < 0) { AtoB[i] = j; BtoA[j] = i; break; }
    for (int i = 0; i < rA; ++i) {
        if (AtoB[i] < 0) continue;
        int j = AtoB[i];
        if (A.nsect[i] != B.nsect[j]) { set_error("contract: contracted indices have different sector counts"); return ITB_ERR_INVALID; }
        for (int s = 0; s < A.nsect[i]; ++s)
            if (A.ext(i, s) != B.ext(j, s)) { set_error("contract: contracted indices have different sector sizes"); return ITB_ERR_INVALID; }
    }
    // ---- result index order (contractIS, sortResult=false) -----------------------------------
    std::vector<int> AtoC(rA, -1), BtoC(rB, -1);
    C = TensorStruct();
    C.dtype = (A.dtype == ITB_C64 || B.dtype == ITB_C64) ? ITB_C64 : ITB_F64;
    P.labC.clear();
    C.sect_start.assign(1, 0);
    auto push_index = [&](const TensorStruct& T, int j, int32_t lab) {
        C.nsect.push_back(T.nsect[j]);
        for (int s = 0; s < T.nsect[j]; ++s) C.sect.push_back(T.ext(j, s));
        C.sect_start.push_back((int64_t)C.sect.size());
        P.labC.push_back(lab);
    };
    for (int i = 0; i < rA; ++i)
        if (AtoB[i] < 0) { AtoC[i] = (int)C.nsect.size(); push_index(A, i, P.labA[i]); }
    for (int j = 0; j < rB; ++j)
        if (BtoA[j] < 0) { BtoC[j] = (int)C.nsect.size(); push_index(B, j, P.labB[j]); }
    const int rC = (int)C.nsect.size();
    if (rC > ITB_MAX_ORDER) { set_error("contract: result order too large"); return ITB_ERR_UNSUPPORTED; }
    C.order = rC;

    PLAN_PHASE(6);
    // ---- pair enumeration + C block list ----------------------------------------------------------
    std::vector<int> contA, contB;
    for (int i = 0; i < rA; ++i)
        if (AtoB[i] >= 0) { contA.push_back(i); contB.push_back(AtoB[i]); }
    // Fast path: block coordinates packed into one 64-bit key per tuple (sector counts are small, so the C tuple
    // almost always fits). The LAST index goes to the most significant bits, which makes integer order the
    // reference's block order (qdense.cc:60-65). Same pairs in the same order as the generic path below.
    auto bits_for = [](int32_t n) { int b = 1; while ((1ll << b) < n) ++b; return b; };
    int cbits = 0, kbits = 0;
    std::vector<int> c_shift(rC, 0), c_width(rC, 0);
    for (int j = 0; j < rC; ++j) { c_shift[j] = cbits; c_width[j] = bits_for(C.nsect[j]); cbits += c_width[j]; }
    std::vector<int> k_shift(contA.size(), 0);
    for (size_t c = 0; c < contA.size(); ++c) { k_shift[c] = kbits; kbits += bits_for(A.nsect[contA[c]]); }
    struct PairRec { int64_t a, b; std::vector<int32_t> cb; };
    std::vector<PairRec> recs;           // generic path
    struct PackedRec { int64_t a, b; uint64_t ck; };
    std::vector<PackedRec> precs;        // fast path
    std::unordered_map<std::vector<int32_t>, int64_t, VecHash> cpos;
    std::vector<uint64_t> ckeys;         // fast path: sorted unique C keys
    static const bool force_generic = getenv("ITB_PLAN_GENERIC") != nullptr; // tests: exercise the vector-key path
    const bool packed = cbits <= 63 && kbits <= 63 && !force_generic;
    if (packed) {
        std::vector<std::pair<uint64_t, int64_t>> bkeys(B.nblocks); // (contracted key, B block), sorted: B order kept inside a key
        for (int64_t b = 0; b < B.nblocks; ++b) {
            uint64_t k = 0;
            for (size_t c = 0; c < contB.size(); ++c) k |= (uint64_t)B.block(b)[contB[c]] << k_shift[c];
            bkeys[b] = {k, b};
        }
        std::sort(bkeys.begin(), bkeys.end());
        std::vector<uint64_t> bpart(B.nblocks); // C-key contribution of every B block
        for (int64_t b = 0; b < B.nblocks; ++b) {
            uint64_t k = 0;
            for (int j = 0; j < rB; ++j)
                if (BtoC[j] >= 0) k |= (uint64_t)B.block(b)[j] << c_shift[BtoC[j]];
            bpart[b] = k;
        }
        precs.reserve((size_t)A.nblocks * 2);
        for (int64_t a = 0; a < A.nblocks; ++a) {
            uint64_t k = 0, apart = 0;
            for (size_t c = 0; c < contA.size(); ++c) k |= (uint64_t)A.block(a)[contA[c]] << k_shift[c];
            for (int i = 0; i < rA; ++i)
                if (AtoC[i] >= 0) apart |= (uint64_t)A.block(a)[i] << c_shift[AtoC[i]];
            auto lo = std::lower_bound(bkeys.begin(), bkeys.end(), std::make_pair(k, (int64_t)-1));
            for (auto it = lo; it != bkeys.end() && it->first == k; ++it) precs.push_back({a, it->second, apart | bpart[it->second]});
        }
        ckeys.reserve(precs.size());
        for (auto& r : precs) ckeys.push_back(r.ck);
        std::sort(ckeys.begin(), ckeys.end());
        ckeys.erase(std::unique(ckeys.begin(), ckeys.end()), ckeys.end());
        C.nblocks = (int64_t)ckeys.size();
        C.blocks.resize(C.nblocks * rC);
        C.offsets.resize(C.nblocks);
        int64_t off = 0;
        for (int64_t c = 0; c < C.nblocks; ++c) {
            int64_t sz = 1;
            for (int j = 0; j < rC; ++j) {
                const int32_t v = (int32_t)((ckeys[c] >> c_shift[j]) & ((1ull << c_width[j]) - 1));
                C.blocks[c * rC + j] = v;
                sz *= C.ext(j, v);
            }
            C.offsets[c] = off;
            off += sz;
        }
        C.nelems = off;
    } else {
        std::unordered_map<std::vector<int32_t>, std::vector<int64_t>, VecHash> bucket;
        std::vector<int32_t> key(contA.size());
        for (int64_t b = 0; b < B.nblocks; ++b) {
            for (size_t c = 0; c < contB.size(); ++c) key[c] = B.block(b)[contB[c]];
            bucket[key].push_back(b);
        }
        std::vector<int32_t> cb(rC);
        for (int64_t a = 0; a < A.nblocks; ++a) {
            for (size_t c = 0; c < contA.size(); ++c) key[c] = A.block(a)[contA[c]];
            auto it = bucket.find(key);
            if (it == bucket.end()) continue;
            for (int i = 0; i < rA; ++i)
                if (AtoC[i] >= 0) cb[AtoC[i]] = A.block(a)[i];
            for (int64_t b : it->second) {
                for (int j = 0; j < rB; ++j)
                    if (BtoC[j] >= 0) cb[BtoC[j]] = B.block(b)[j];
                recs.push_back({a, b, cb});
            }
        }
        // C block list: sort + unique by the reference ordering, prefix-sum offsets
        std::vector<std::vector<int32_t>> cblocks;
        cblocks.reserve(recs.size());
        for (auto& r : recs) cblocks.push_back(r.cb);
        std::sort(cblocks.begin(), cblocks.end(),
                  [rC](const std::vector<int32_t>& x, const std::vector<int32_t>& y) { return block_less(x.data(), y.data(), rC); });
        cblocks.erase(std::unique(cblocks.begin(), cblocks.end()), cblocks.end());
        C.nblocks = (int64_t)cblocks.size();
        C.blocks.resize(C.nblocks * rC);
        C.offsets.resize(C.nblocks);
        int64_t off = 0;
        for (int64_t c = 0; c < C.nblocks; ++c) {
            int64_t sz = 1;
            for (int j = 0; j < rC; ++j) {
                C.blocks[c * rC + j] = cblocks[c][j];
                sz *= C.ext(j, cblocks[c][j]);
            }
            C.offsets[c] = off;
            off += sz;
            cpos[cblocks[c]] = c;
        }
        C.nelems = off;
    }
    PLAN_PHASE(7);
    const size_t nrec = packed ? precs.size() : recs.size();
    // ---- triples + flops -----------------------------------------------------------------------
    P.triples.resize(nrec * 3);
    P.flops = 0;
    const double cmul = (A.dtype == ITB_C64 ? 2.0 : 1.0) * (B.dtype == ITB_C64 ? 2.0 : 1.0);
    // 2*M*N*K of a pair = 2 * (elements of the A block) * (product of B's uncontracted extents): one product per block
    std::vector<double> a_elems((size_t)A.nblocks, 1.0), b_unc((size_t)B.nblocks, 1.0);
    for (int64_t a = 0; a < A.nblocks; ++a)
        for (int i = 0; i < rA; ++i) a_elems[a] *= (double)A.ext(i, A.block(a)[i]);
    for (int64_t b = 0; b < B.nblocks; ++b)
        for (int j = 0; j < rB; ++j)
            if (BtoA[j] < 0) b_unc[b] *= (double)B.ext(j, B.block(b)[j]);
    for (size_t p = 0; p < nrec; ++p) {
        const int64_t ra = packed ? precs[p].a : recs[p].a, rb = packed ? precs[p].b : recs[p].b;
        P.triples[3 * p + 0] = ra;
        P.triples[3 * p + 1] = rb;
        P.triples[3 * p + 2] = packed ? (int64_t)(std::lower_bound(ckeys.begin(), ckeys.end(), precs[p].ck) - ckeys.begin()) : cpos[recs[p].cb];
        P.flops += 2.0 * a_elems[ra] * b_unc[rb] * cmul;
    }
    PLAN_PHASE(6);
    P.tables_built = false;
    return ITB_OK;
}

// ---- per-pair index-group analysis ---------------------------------------------------------------
struct Dim {
    int64_t ext;
    int64_t sa, sb; // strides (REAL units) in the tensors that carry this dim (sb unused for M/N)
};

// index-group scratch without heap traffic (the planner runs once per new block structure, i.e. at every DMRG bond)
struct DimList {
    Dim d[ITB_MAX_ORDER + 2];
    int n = 0;
    void push_back(const Dim& x) { d[n++] = x; }
    void push_front(const Dim& x) { for (int i = n; i > 0; --i) d[i] = d[i - 1]; d[0] = x; ++n; }
    size_t size() const { return (size_t)n; }
    Dim& operator[](size_t i) { return d[i]; }
    const Dim* begin() const { return d; }
    const Dim* end() const { return d + n; }
};

// drop unit dims, then fuse neighbours that are contiguous in every tensor carrying the group
static void canon(DimList& g, bool two) {
    int o = 0;
    for (int i = 0; i < g.n; ++i) {
        const Dim d = g.d[i];
        if (d.ext == 1) continue;
        if (o > 0) {
            Dim& l = g.d[o - 1];
            const bool ok = (d.sa == l.sa * l.ext) && (!two || d.sb == l.sb * l.ext);
            if (ok) { l.ext *= d.ext; continue; }
        }
        g.d[o++] = d;
    }
    g.n = o;
}

// ---- cycle model of the DMMA tile kernel (per BK=16 K-chunk of one tile, per SM) -----------------------------
// Consumer warp w sits at (wm,wn) = ((w ^ (w>>2)) & 3, w>>2) of a 4x4 grid and runs on SMSP w&3, so every SMSP
// holds one warp of each warp-row and each warp-column. Warps skip 8x8 fragments that lie entirely outside the
// valid part of an edge tile, so the tensor-pipe time of a chunk is 4 k-steps x 16 cycles x the fragment
// products of the busiest SMSP; below that the producers (gather + barriers) set a floor per configuration.
// Constants measured on B200 (tools/tile_calib.py, tools/sched_fit.py; profiles/README.md): a full 128x128 chunk
// takes 4550 cycles (1.11 x its 4096 DMMA cycles); the four producer warps need ~2950 cycles to gather a 128x128
// chunk from cold operands however few of its rows are valid (8-byte LDGSTS path), 1700 for 64x64, 1050 for 32x32,
// which is what an edge tile costs. On top: ~4000 cycles per item plus ~3500 per block pair it walks.
// Least-squares fit of measured per-CTA cycles on the maxdim-2000 H_eff workload (tools/sched_fit.py): a 128x128
// chunk costs 2466 + 0.54 x (DMMA cycles of the busiest SMSP), i.e. 4680 full and ~3150 for a 34-wide edge.
static double g_tile_floor[ITB_NCFG] = {2950.0, 1700.0, 1050.0};
static double kTileOverhead[ITB_NCFG] = {3800.0, 8000.0, 5300.0}; // 128x128: epilogue ~3000 + ~800 to the next item (tools/tile_probe.py)
static double kPairOverhead = 1000.0; // per block pair an item walks (K loop of the 3-pair *R tiles: +85 cycles per chunk)
static const double kDmmaSlack = 1.11;
static int kForceCfg = -1;
static int64_t kRowGroupMinBytes = 32ll << 20; // streaming class smaller than this (bytes of C): row groups only once the plan proves hot
static int64_t kRowGroupAfterRuns = 3;          // ... i.e. from this execution of the plan on (ITB_ROWGROUP_TIER="bytes,runs")
static bool kUseRowGroups = true; // ITB_ROWGROUPS=0 routes every streaming C block to the C-stationary kernels
static int64_t kMinPiece = 8; // K-chunks: never cut a tile into pieces shorter than this (ITB_MIN_PIECE)
static double kGuidedFactor = 2.0;  // shared queue: piece cost = remaining work / (kGuidedFactor x grid width); ITB_GUIDED_FACTOR
static bool kSchedStreamK = true;   // ITB_SCHED=streamk (default) | guided: static stream-K partition or the guided dynamic queue.
                                    // Same-box A/B on the bench workload (profiles/r03_tile_schedule_ab.txt): static partition +
                                    // static kernel 22.9 TFLOP/s; guided queue + ring kernel 22.1-22.2 (better balance, 1.03 vs
                                    // 1.10 max/mean, and 40 % less DRAM traffic, but +4.5 % CTA cycles for the same items and a
                                    // three times larger split-K reduction)
static double kStaticFrac = 0.0;    // fraction of the modelled work handed out as static per-CTA ranges; ITB_STATIC_FRAC (0: all dynamic).
                                    // Measured (profiles/r03_tile_schedule_variants.txt): 0 -> 1378 us per H_eff*phi, 0.85 -> 1391-1419, 0.92 -> 1412
static void read_tile_env() {
    static bool done = false;
    if (done) return;
    done = true;
    if (const char* e = getenv("ITB_TILE_FLOOR")) sscanf(e, "%lf,%lf,%lf", &g_tile_floor[0], &g_tile_floor[1], &g_tile_floor[2]);
    if (const char* e = getenv("ITB_TILE_OVERHEAD")) sscanf(e, "%lf,%lf,%lf,%lf", &kTileOverhead[0], &kTileOverhead[1], &kTileOverhead[2], &kPairOverhead);
    if (const char* e = getenv("ITB_FORCE_CFG")) kForceCfg = atoi(e);
    if (const char* e = getenv("ITB_ROWGROUPS")) kUseRowGroups = atoi(e) != 0;
    if (const char* e = getenv("ITB_ROWGROUP_TIER")) { long long b = 0, r = 0; if (sscanf(e, "%lld,%lld", &b, &r) == 2) { kRowGroupMinBytes = b; kRowGroupAfterRuns = r; } }
    if (const char* e = getenv("ITB_GUIDED_FACTOR")) kGuidedFactor = std::max(0.25, atof(e));
    if (const char* e = getenv("ITB_MIN_PIECE")) kMinPiece = std::max(1, atoi(e));
    if (const char* e = getenv("ITB_SCHED")) kSchedStreamK = std::string(e) != "guided";
    if (kSchedStreamK && !getenv("ITB_TILE_OVERHEAD")) {
        // the static partition is only as good as its cycle model: these are the constants fitted WITH this partition to
        // per-CTA clock64 spans of the static kernel (tools/sched_fit.py); the guided queue uses per-item measurements
        kTileOverhead[0] = 1200.0; kPairOverhead = 4300.0;
    }
    if (const char* e = getenv("ITB_STATIC_FRAC")) kStaticFrac = std::min(0.98, std::max(0.0, atof(e)));
}
static double chunk_cycles_compute(int f, int64_t vm, int64_t vn);
// the model only depends on the number of 8-row / 8-column fragments that hold valid data: one table per configuration
static double chunk_cycles(int f, int64_t vm, int64_t vn) {
    static double tab[ITB_NCFG][17][17];
    static bool built = false;
    if (!built) {
        for (int g = 0; g < ITB_NCFG; ++g)
            for (int a = 0; a <= kTileM[g] / 8; ++a)
                for (int b = 0; b <= kTileN[g] / 8; ++b) tab[g][a][b] = chunk_cycles_compute(g, 8 * a, 8 * b);
        built = true;
    }
    const int64_t a = std::min<int64_t>((std::max<int64_t>(vm, 0) + 7) / 8, kTileM[f] / 8), b = std::min<int64_t>((std::max<int64_t>(vn, 0) + 7) / 8, kTileN[f] / 8);
    return tab[f][a][b];
}
static double chunk_cycles_compute(int f, int64_t vm, int64_t vn) {
    const int WM = kTileM[f] / 4, WN = kTileN[f] / 4, FM = WM / 8, FN = WN / 8;
    int fm[4], fn[4];
    for (int i = 0; i < 4; ++i) {
        fm[i] = (int)std::max<int64_t>(0, std::min<int64_t>(FM, (vm - i * WM + 7) / 8));
        fn[i] = (int)std::max<int64_t>(0, std::min<int64_t>(FN, (vn - i * WN + 7) / 8));
    }
    int worst = 0;
    for (int s = 0; s < 4; ++s) {
        int sum = 0;
        for (int j = 0; j < 4; ++j) sum += fm[s ^ j] * fn[j];
        worst = std::max(worst, sum);
    }
    if (f == ITB_CFG_BIG) return 2466.0 + 0.54 * 64.0 * worst;
    return std::max(kDmmaSlack * 64.0 * worst, g_tile_floor[f]);
}
static double cblk_cost(int f, int64_t M, int64_t N, double nch, int npairs) {
    const int TM = kTileM[f], TN = kTileN[f];
    double cost = 0;
    // full tiles + the (up to) three distinct edge shapes
    const int64_t fm = M / TM, fn = N / TN, rm = M % TM, rn = N % TN;
    const double ovh = kTileOverhead[f] + kPairOverhead * npairs;
    if (fm && fn) cost += (double)fm * fn * (chunk_cycles(f, TM, TN) * nch + ovh);
    if (rm) cost += (double)fn * (chunk_cycles(f, rm, TN) * nch + ovh);
    if (rn) cost += (double)fm * (chunk_cycles(f, TM, rn) * nch + ovh);
    if (rm && rn) cost += chunk_cycles(f, rm, rn) * nch + ovh;
    return cost;
}

bool plan_note_run(itb_contract_plan& P) {
    read_tile_env();
    ++P.runs;
    if (P.tables_built && P.rg_deferred && P.runs >= kRowGroupAfterRuns) { P.tables_built = false; return true; }
    return false;
}

int build_contract_tables(itb_contract_plan& P) {
    PLAN_T0();
    const TensorStruct &A = P.A, &B = P.B, &C = P.C;
    const int rA = A.order, rB = B.order;
    std::vector<int> AtoB(rA, -1), BtoA(rB, -1);
    for (int i = 0; i < rA; ++i)
        for (int j = 0; j < rB; ++j)
            if (P.labA[i] == P.labB[j] && BtoA[j] < 0) { AtoB[i] = j; BtoA[j] = i; break; }
    const bool cA = A.dtype == ITB_C64, cB = B.dtype == ITB_C64;
    const int64_t csA = cA ? 2 : 1, csB = cB ? 2 : 1, csC = (cA || cB) ? 2 : 1;

    P.pairs.clear(); P.cblks.clear(); P.skinny.clear(); P.skinny_q4.clear(); P.skinny_q8.clear(); P.dots.clear(); P.dot_outs.clear();
    P.tiles.clear(); P.qitems.clear(); P.splits.clear(); P.ws_slots = 0; P.cta_begin.clear();
    P.rgroups.clear(); P.rg_in.clear(); P.rg_out.clear(); P.rg_w.clear(); P.rg_items.clear();
    P.ndot_slots = 0;

    const int64_t npairs = (int64_t)P.triples.size() / 3;
    // group pairs by C block, keeping the reference enumeration order inside a block
    // (counting sort on the C block number: stable and linear)
    std::vector<int64_t> order(npairs);
    {
        std::vector<int64_t> start((size_t)C.nblocks + 1, 0);
        for (int64_t p = 0; p < npairs; ++p) ++start[(size_t)P.triples[3 * p + 2] + 1];
        for (int64_t c = 0; c < C.nblocks; ++c) start[c + 1] += start[c];
        for (int64_t p = 0; p < npairs; ++p) order[(size_t)start[(size_t)P.triples[3 * p + 2]]++] = p;
    }
    const int64_t cb_first = P.cb_first, cb_last = (P.cb_last < 0 ? C.nblocks : P.cb_last);

    PLAN_PHASE(0);
    std::vector<int> uncA_all, uncB_all; // uncontracted indices of A / B in order == the index list of C
    for (int i = 0; i < rA; ++i) if (AtoB[i] < 0) uncA_all.push_back(i);
    for (int j = 0; j < rB; ++j) if (BtoA[j] < 0) uncB_all.push_back(j);
    if (P.slice_index >= (int)(uncA_all.size() + uncB_all.size())) { set_error("contract: slice index out of range"); return ITB_ERR_INVALID; }
    std::vector<int64_t> pair_ia; // A block of every entry of P.pairs (host only)
    std::vector<int64_t> cblk_sl;  // per executed C block: extent of the sliced A index (0: not sliced on the A side)
    P.pairs.reserve((size_t)npairs); pair_ia.reserve((size_t)npairs);
    P.cblks.reserve((size_t)C.nblocks); cblk_sl.reserve((size_t)C.nblocks);
    // per-block extents / strides / sizes / fastest non-unit index, computed once per block instead of once per pair
    // (a block takes part in several pairs: 2.3 per A block and 30 per B block in the MPO steps at Hubbard scale)
    std::vector<int64_t> extA_all((size_t)A.nblocks * rA), strA_all((size_t)A.nblocks * rA), sizeA((size_t)A.nblocks);
    std::vector<int64_t> extB_all((size_t)B.nblocks * rB), strB_all((size_t)B.nblocks * rB), sizeB((size_t)B.nblocks);
    std::vector<int8_t> fastA((size_t)A.nblocks, -1), fastB((size_t)B.nblocks, -1);
    auto fill_block_tables = [](const TensorStruct& T, int r, int64_t cs, std::vector<int64_t>& ext, std::vector<int64_t>& str,
                                std::vector<int64_t>& size, std::vector<int8_t>& fast) {
        for (int64_t b = 0; b < T.nblocks; ++b) {
            const int32_t* blk = T.block(b);
            int64_t st = cs;
            for (int i = 0; i < r; ++i) {
                const int64_t e = T.ext(i, blk[i]);
                ext[(size_t)b * r + i] = e; str[(size_t)b * r + i] = st; st *= e;
                if (e > 1 && fast[b] < 0) fast[b] = (int8_t)i;
            }
            size[b] = st;
        }
    };
    fill_block_tables(A, rA, csA, extA_all, strA_all, sizeA, fastA);
    fill_block_tables(B, rB, csB, extB_all, strB_all, sizeB, fastB);
    int64_t pos = 0;
    while (pos < npairs) {
        const int64_t ic = P.triples[3 * order[pos] + 2];
        int64_t end = pos;
        while (end < npairs && P.triples[3 * order[end] + 2] == ic) ++end;
        if (ic < cb_first || ic >= cb_last || (!P.cb_mask.empty() && !P.cb_mask[ic])) { pos = end; continue; }
        // row-slice of one C index (sharding inside a sector): which operand index it is and the range in this block
        int sl_a = -1, sl_b = -1; // position of the sliced index in A / in B
        int64_t sl_lo = 0, sl_hi = 0;
        if (P.slice_index >= 0) {
            const int j = P.slice_index;
            const int32_t sec = C.block(ic)[j];
            sl_lo = P.slice_lo[sec]; sl_hi = P.slice_hi[sec];
            if (sl_hi <= sl_lo) { pos = end; continue; }
            if (sl_lo != 0 || sl_hi != C.ext(j, sec)) {
                if (j < (int)uncA_all.size()) sl_a = uncA_all[j]; else sl_b = uncB_all[j - (int)uncA_all.size()];
            }
        }

        ItbCBlk cbk;
        std::memset(&cbk, 0, sizeof(cbk));
        cbk.pair_begin = (int32_t)P.pairs.size();
        int64_t M = 1, N = 1; // complex-element dims of this C block
        for (int64_t q = pos; q < end; ++q) {
            const int64_t p = order[q];
            const int64_t ia = P.triples[3 * p], ib = P.triples[3 * p + 1];
            const int64_t* extA = extA_all.data() + (size_t)ia * rA;
            const int64_t* extB = extB_all.data() + (size_t)ib * rB;
            const int64_t* strA = strA_all.data() + (size_t)ia * rA;
            const int64_t* strB = strB_all.data() + (size_t)ib * rB;
            DimList gm, gk, gn;
            // which operand-fastest (non-unit) index is contracted?
            const int fa = fastA[ia], fb = fastB[ib];
            const bool a_kfast = fa >= 0 && AtoB[fa] >= 0;
            const bool b_kfast = fb >= 0 && BtoA[fb] >= 0;
            int64_t a_shift = 0, b_shift = 0; // element shift of the operand block base when its uncontracted index is sliced
            for (int i = 0; i < rA; ++i)
                if (AtoB[i] < 0) {
                    if (i == sl_a) { gm.push_back({sl_hi - sl_lo, strA[i], 0}); a_shift = sl_lo * strA[i]; }
                    else {
                        if (sl_a >= 0 && i > sl_a && extA[i] != 1) { set_error("contract: sliced index is not the slowest non-unit uncontracted index of A"); return ITB_ERR_UNSUPPORTED; }
                        gm.push_back({extA[i], strA[i], 0});
                    }
                }
            for (int j = 0; j < rB; ++j)
                if (BtoA[j] < 0) {
                    if (j == sl_b) { gn.push_back({sl_hi - sl_lo, strB[j], 0}); b_shift = sl_lo * strB[j]; }
                    else {
                        if (sl_b >= 0 && j > sl_b && extB[j] != 1) { set_error("contract: sliced index is not the slowest non-unit uncontracted index of B"); return ITB_ERR_UNSUPPORTED; }
                        gn.push_back({extB[j], strB[j], 0});
                    }
                }
            // K order: follow A unless only B is k-fast (keeps the k-fast operand contiguous in k)
            if (a_kfast || !b_kfast) {
                for (int i = 0; i < rA; ++i)
                    if (AtoB[i] >= 0) gk.push_back({extA[i], strA[i], strB[AtoB[i]]});
            } else {
                for (int j = 0; j < rB; ++j)
                    if (BtoA[j] >= 0) gk.push_back({extB[j], strA[BtoA[j]], strB[j]});
            }
            int64_t m = 1, n = 1, k = 1;
            for (auto& d : gm) m *= d.ext;
            for (auto& d : gn) n *= d.ext;
            for (auto& d : gk) k *= d.ext;
            M = m; N = n;
            // ---- complex folding: prepend a pseudo-dim of extent 2 (re,im) -------------------------
            int flags = 0;
            if (cA && cB) { // [Cr;Ci] = [[Ar,-Ai],[Ai,Ar]] [Br;Bi]
                gm.push_front({2, 1, 0});
                gk.push_front({2, 1, 1});
                flags |= ITB_PF_CCA;
            } else if (cA) { // rows of A doubled: C'(2m+p,n) = sum_k A(m,k).comp(p) B(k,n)
                gm.push_front({2, 1, 0});
            } else if (cB) { // columns of B doubled: C'(m,2n+q) = sum_k A(m,k) B(k,n).comp(q)
                gn.push_front({2, 1, 0});
            }
            canon(gm, false); canon(gk, true); canon(gn, false);
            if (gm.size() > ITB_MAXG || gk.size() > ITB_MAXG || gn.size() > ITB_MAXG) {
                set_error("contract: more than ITB_MAXG non-fusable indices in one group");
                return ITB_ERR_UNSUPPORTED;
            }
            ItbPair pr;
            std::memset(&pr, 0, sizeof(pr));
            pr.a_off = A.offsets[ia] * csA + a_shift; // (strA/strB are in REAL units already)
            pr.b_off = B.offsets[ib] * csB + b_shift;
            for (int d = 0; d < ITB_MAXG; ++d) { pr.m_ext[d] = pr.n_ext[d] = pr.k_ext[d] = 1; }
            for (size_t d = 0; d < gm.size(); ++d) { pr.m_ext[d] = (int32_t)gm[d].ext; pr.am_str[d] = gm[d].sa; }
            for (size_t d = 0; d < gn.size(); ++d) { pr.n_ext[d] = (int32_t)gn[d].ext; pr.bn_str[d] = gn[d].sa; }
            for (size_t d = 0; d < gk.size(); ++d) { pr.k_ext[d] = (int32_t)gk[d].ext; pr.ak_str[d] = gk[d].sa; pr.bk_str[d] = gk[d].sb; }
            pr.m_n = (int32_t)gm.size(); pr.n_n = (int32_t)gn.size(); pr.k_n = (int32_t)gk.size();
            const int64_t Kr = k * ((cA && cB) ? 2 : 1);
            const int64_t abl = sizeA[ia], bbl = sizeB[ib]; // real elements of the two blocks: device row offsets are 32-bit
            if (Kr >= (1ll << 31) || m * 2 >= (1ll << 31) || n * 2 >= (1ll << 31) || abl >= (1ll << 31) || bbl >= (1ll << 31)) {
                set_error("contract: block dimension exceeds 2^31");
                return ITB_ERR_UNSUPPORTED;
            }
            pr.K = (int32_t)Kr;
            if (a_kfast) flags |= ITB_PF_A_KFAST;
            if (b_kfast) flags |= ITB_PF_B_KFAST;
            pr.flags = flags;
            cbk.ksum += Kr;
            P.pairs.push_back(pr);
            pair_ia.push_back(ia);
        }
        cbk.pair_end = (int32_t)P.pairs.size();
        // the C block keeps its full leading dimension; a slice only moves the origin and shrinks M or N
        int64_t Mfull = 1, c_shift = 0;
        for (size_t u = 0; u < uncA_all.size(); ++u) Mfull *= C.ext((int)u, C.block(ic)[u]);
        if (sl_a >= 0 || sl_b >= 0) {
            int64_t st = 1;
            for (int j = 0; j < P.slice_index; ++j) st *= C.ext(j, C.block(ic)[j]);
            c_shift = sl_lo * st;
        }
        cbk.c_off = (C.offsets[ic] + c_shift) * csC;
        cbk.M = (int32_t)(M * (cA ? 2 : 1));
        cbk.N = (int32_t)(N * ((!cA && cB) ? 2 : 1));
        if (!cA && cB) { cbk.c_ms = 2; cbk.c_nmask = 1; cbk.c_nshift = 1; cbk.c_ns = 2 * Mfull; }
        else { cbk.c_ms = 1; cbk.c_nmask = 0; cbk.c_nshift = 0; cbk.c_ns = Mfull * (cA ? 2 : 1); }
        P.cblks.push_back(cbk);
        cblk_sl.push_back(sl_a >= 0 ? sl_hi - sl_lo : 0);
        pos = end;
    }

    PLAN_PHASE(1);
    // ---- classify C blocks into kernel work lists --------------------------------------------------
    read_tile_env();
    auto chunks_of = [&](const ItbCBlk& cb) {
        int64_t n = 0;
        for (int32_t p = cb.pair_begin; p < cb.pair_end; ++p) n += (P.pairs[p].K + ITB_BK - 1) / ITB_BK;
        return n;
    };
    for (double& f : P.class_flops) f = 0;
    std::vector<std::pair<int32_t, int>> tile_cblks; // (C block, tile config)
    std::vector<int32_t> rg_cands;                   // streaming C blocks eligible for the row-group kernel
    std::vector<int32_t> stream_cands;               // skinny C blocks (streaming class unless they are too few to matter)
    auto to_tiles = [&](int32_t c) {                 // tile class: pick the config with the least modelled cycles
        const ItbCBlk& cb = P.cblks[c];
        int best = 0; double bestc = 1e300;
        for (int f = 0; f < ITB_NCFG; ++f) {
            const double cost = cblk_cost(f, cb.M, cb.N, (double)chunks_of(cb), cb.pair_end - cb.pair_begin);
            if (cost < bestc) { bestc = cost; best = f; }
        }
        if (kForceCfg >= 0) best = kForceCfg;
        P.class_flops[best] += 2.0 * (double)cb.M * (double)cb.N * (double)cb.ksum;
        tile_cblks.push_back({c, best});
    };
    auto push_skinny = [&](int32_t c) {              // C-stationary streaming kernels (one C block at a time)
        const ItbCBlk& cb = P.cblks[c];
        const int64_t M = cb.M, N = cb.N;
        const int long_is_n = (N > M) ? 1 : 0;
        const int64_t L = long_is_n ? N : M, Sh = long_is_n ? M : N;
        if (cb.pair_end - cb.pair_begin <= kSkinnyQMaxPairs) {
            auto& list = Sh <= 4 ? P.skinny_q4 : P.skinny_q8;
            for (int64_t r0 = 0; r0 < L; r0 += kSkinnyQRows)
                list.push_back({c, (int32_t)r0, (int32_t)std::min<int64_t>(kSkinnyQRows, L - r0), long_is_n});
        } else {
            for (int64_t r0 = 0; r0 < L; r0 += kSkinnyRows)
                P.skinny.push_back({c, (int32_t)r0, (int32_t)std::min<int64_t>(kSkinnyRows, L - r0), long_is_n});
        }
    };
    for (int32_t c = 0; c < (int32_t)P.cblks.size(); ++c) {
        const ItbCBlk& cb = P.cblks[c];
        const int64_t M = cb.M, N = cb.N;
        const double cflops = 2.0 * (double)M * (double)N * (double)cb.ksum; // real-expanded == x2/x4 complex count
        if (M * N <= kDotMaxMN) {
            P.class_flops[4] += cflops;
            ItbDotOut o{c, (int32_t)P.ndot_slots, 0, 0};
            for (int32_t p = cb.pair_begin; p < cb.pair_end; ++p) {
                const int32_t K = P.pairs[p].K;
                for (int32_t k0 = 0; k0 < K; k0 += kDotChunk) {
                    ItbDot d{c, p, k0, std::min<int32_t>(kDotChunk, K - k0), (int32_t)P.ndot_slots, {0, 0, 0}};
                    P.dots.push_back(d);
                    ++P.ndot_slots; ++o.nslots;
                }
            }
            P.dot_outs.push_back(o);
        } else if (std::min(M, N) <= kSkinnyMax && cb.ksum <= kSkinnyQMaxK && M * N >= kSkinnyMinElems) {
            // HBM-bound streaming class: short side <= 8 and a short K loop (the MPO steps of H_eff*phi)
            stream_cands.push_back(c);
        } else {
            to_tiles(c);
        }
    }
    {
        int64_t stream_elems = 0;
        for (int32_t c : stream_cands) stream_elems += (int64_t)P.cblks[c].M * P.cblks[c].N;
        const bool ride = !tile_cblks.empty() && stream_elems < kStreamMinTotal;
        // Tiered planning of the streaming class. The row-group tables below cost about as much host time as all the rest
        // of the plan (0.7-1.0 ms for the MPO steps at Hubbard scale) and buy ~2x on an HBM-bound kernel: worth it for a
        // step that streams tens of MB or that keeps coming back (Davidson), not for the many contractions of a DMRG bond
        // that run once on a few MB. Small streaming classes therefore start on the C-stationary kernels (their item lists
        // are trivial to build); itb_contract_run re-plans with row groups when the plan reaches its kRowGroupAfterRuns-th
        // execution.
        const bool want_rg = kUseRowGroups && (stream_elems * 8 >= kRowGroupMinBytes || P.runs >= kRowGroupAfterRuns);
        P.rg_deferred = false;
        for (int32_t c : stream_cands) {
            if (ride) { to_tiles(c); continue; }
            const ItbCBlk& cb = P.cblks[c];
            P.class_flops[3] += 2.0 * (double)cb.M * (double)cb.N * (double)cb.ksum;
            // All four real / complex pairings stream through the row-group kernel when A is the long operand. Real B with
            // complex A: interleaved (re,im) makes a complex A block a real block with a doubled leading extent. Complex B:
            // the output slots are the (re,im) components of C (row stride 2), and for complex A the input slots are the
            // components of A with the weights +-B.re / B.im of the complex product.
            const bool eligible = cb.M >= cb.N && kUseRowGroups;
            if (eligible && want_rg) rg_cands.push_back(c); // row-group kernel (below); else C-stationary kernels
            else { push_skinny(c); P.rg_deferred |= eligible; }
        }
    }
    PLAN_PHASE(2);
    // ---- row groups: streaming C blocks that share their long-side (A-uncontracted) block coordinates ----------
    if (!rg_cands.empty()) {
        std::vector<int> uncA; // A's uncontracted indices, in order (they lead the C index list)
        for (int i = 0; i < rA; ++i) if (AtoB[i] < 0) uncA.push_back(i);
        auto host_off = [](int64_t idx, const int32_t* ext, const int64_t* str, int n) {
            int64_t o = 0;
            for (int d = 0; d < n; ++d) {
                if (d == n - 1) { o += idx * str[d]; break; }
                o += (idx % ext[d]) * str[d];
                idx /= ext[d];
            }
            return o;
        };
        // candidates sorted by their long-side coordinates (first appearance order of the groups is kept)
        // (keys are flat: uncA.size() coordinates per candidate, compared lexicographically)
        const size_t nu = uncA.size();
        std::vector<int32_t> cand_key(rg_cands.size() * std::max<size_t>(nu, 1));
        for (size_t i = 0; i < rg_cands.size(); ++i) {
            const int32_t* ab = A.block(pair_ia[P.cblks[rg_cands[i]].pair_begin]);
            for (size_t u = 0; u < nu; ++u) cand_key[i * nu + u] = ab[uncA[u]];
        }
        auto key_less = [&](size_t x, size_t y) { return std::lexicographical_compare(&cand_key[x * nu], &cand_key[x * nu] + nu, &cand_key[y * nu], &cand_key[y * nu] + nu); };
        auto key_eq = [&](size_t x, size_t y) { return std::equal(&cand_key[x * nu], &cand_key[x * nu] + nu, &cand_key[y * nu]); };
        std::vector<size_t> ord(rg_cands.size());
        std::iota(ord.begin(), ord.end(), (size_t)0);
        std::stable_sort(ord.begin(), ord.end(), key_less);
        struct Grp { size_t first, begin, end; }; // first candidate position; members = ord[begin..end)
        std::vector<Grp> groups;
        for (size_t q = 0; q < ord.size();) {
            size_t e = q, first = ord[q];
            while (e < ord.size() && key_eq(ord[e], ord[q])) { first = std::min(first, ord[e]); ++e; }
            groups.push_back({first, q, e});
            q = e;
        }
        std::sort(groups.begin(), groups.end(), [](const Grp& x, const Grp& y) { return x.first < y.first; });
        std::vector<int64_t> A_blocks;
        std::vector<int32_t> cs;
        for (auto& grp : groups) {
            cs.clear();
            for (size_t q = grp.begin; q < grp.end; ++q) cs.push_back(rg_cands[ord[q]]);
            const int32_t* key = &cand_key[grp.first * nu];
            // long-side dims (unit extents dropped), strides per A block of the group
            A_blocks.clear();
            for (int32_t c : cs)
                for (int32_t p = P.cblks[c].pair_begin; p < P.cblks[c].pair_end; ++p)
                    if (std::find(A_blocks.begin(), A_blocks.end(), pair_ia[p]) == A_blocks.end()) A_blocks.push_back(pair_ia[p]);
            struct LDim { int64_t ext; std::vector<int64_t> str; };
            std::vector<LDim> ld;
            const int64_t sl_ext = cblk_sl[cs[0]]; // all C blocks of a group share the long-side sectors, hence the slice
            if (cA && !cB) ld.push_back({2, std::vector<int64_t>(A_blocks.size(), 1)}); // (re,im): fuses with the first long dim below
            for (size_t u = 0; u < uncA.size(); ++u) {
                const bool sliced = sl_ext > 0 && P.slice_index == (int)u; // (C's leading indices are A's uncontracted ones)
                const int64_t e = sliced ? sl_ext : A.ext(uncA[u], key[u]);
                if (e == 1 && !sliced) continue;
                LDim d{e, {}};
                d.str.reserve(A_blocks.size());
                for (int64_t ia : A_blocks) {
                    int64_t st = csA; // strides in REAL elements
                    for (int i = 0; i < uncA[u]; ++i) st *= A.ext(i, A.block(ia)[i]);
                    d.str.push_back(st);
                }
                bool fused = false;
                if (!ld.empty()) {
                    bool ok = true;
                    for (size_t q = 0; q < A_blocks.size(); ++q) ok = ok && d.str[q] == ld.back().str[q] * ld.back().ext;
                    if (ok) { ld.back().ext *= e; fused = true; }
                }
                if (!fused) ld.push_back(std::move(d));
            }
            int64_t L = 1;
            for (auto& d : ld) L *= d.ext;
            const bool cc = cA && cB;                // complex * complex: rows are complex elements, slots are components
            const int out_per_n = cc ? 2 : 1;        // output slots per column of the C block
            bool eligible = (int)ld.size() <= ITB_RG_MAXL && L < (1ll << 30);
            for (int32_t c : cs)
                eligible = eligible && P.cblks[c].N * out_per_n <= ITB_RG_MAXOUT && P.cblks[c].ksum <= ITB_RG_MAXIN && P.cblks[c].M == L * (cc ? 2 : 1);
            if (!eligible) { for (int32_t c : cs) push_skinny(c); continue; }
            if (ld.empty()) ld.push_back({1, std::vector<int64_t>(A_blocks.size(), 0)});
            // sub-groups of whole C blocks: <= ITB_RG_MAXOUT output slots and <= ITB_RG_MAXIN input slots each
            size_t ci = 0;
            while (ci < cs.size()) {
                ItbRowGroup g;
                std::memset(&g, 0, sizeof(g));
                g.nL = (int32_t)ld.size();
                g.ostr = cB ? 2 : 1;
                for (int d = 0; d < ITB_RG_MAXL; ++d) g.ext[d] = d < g.nL ? (int32_t)ld[d].ext : 1;
                g.L = L;
                g.in_begin = (int32_t)P.rg_in.size(); g.out_begin = (int32_t)P.rg_out.size(); g.w_begin = (int32_t)P.rg_w.size();
                int64_t slot_key[ITB_RG_MAXIN]; // A element offset of (block, k) of every input slot taken so far (slot = position)
                auto find_slot = [&](int64_t ae) { for (int32_t q = 0; q < g.nin; ++q) if (slot_key[q] == ae) return q; return (int32_t)-1; };
                while (ci < cs.size()) {
                    const ItbCBlk& cb = P.cblks[cs[ci]];
                    // input slots this C block would add
                    int64_t fresh[ITB_RG_MAXIN + 1]; // (a C block has ksum <= ITB_RG_MAXIN, checked by `eligible`)
                    int nfresh = 0;
                    for (int32_t p = cb.pair_begin; p < cb.pair_end; ++p)
                        for (int32_t k = 0; k < P.pairs[p].K; ++k) {
                            const int64_t ae = P.pairs[p].a_off + host_off(k, P.pairs[p].k_ext, P.pairs[p].ak_str, P.pairs[p].k_n);
                            if (find_slot(ae) < 0 && std::find(fresh, fresh + nfresh, ae) == fresh + nfresh && nfresh <= ITB_RG_MAXIN) fresh[nfresh++] = ae;
                        }
                    if (g.nout > 0 && (g.nout + cb.N * out_per_n > ITB_RG_MAXOUT || g.nin + nfresh > ITB_RG_MAXIN)) break;
                    for (int32_t p = cb.pair_begin; p < cb.pair_end; ++p) {
                        const ItbPair& pr = P.pairs[p];
                        const size_t qa = std::find(A_blocks.begin(), A_blocks.end(), pair_ia[p]) - A_blocks.begin();
                        for (int32_t k = 0; k < pr.K; ++k) {
                            const int64_t ae = pr.a_off + host_off(k, pr.k_ext, pr.ak_str, pr.k_n);
                            int32_t slot = find_slot(ae);
                            if (slot < 0) {
                                ItbRgIn in;
                                std::memset(&in, 0, sizeof(in));
                                in.base = ae;
                                for (int d = 0; d < g.nL; ++d) in.str[d] = ld[d].str[qa];
                                slot = g.nin++;
                                slot_key[slot] = ae;
                                P.rg_in.push_back(in);
                            }
                            const int64_t bk = pr.b_off + host_off(k, pr.k_ext, pr.bk_str, pr.k_n);
                            for (int32_t n = 0; n < cb.N; ++n) {
                                const int64_t bn = host_off(n, pr.n_ext, pr.bn_str, pr.n_n);
                                if (!cc) { P.rg_w.push_back({slot, g.nout + n, bk + bn}); continue; }
                                // input slot = component pp = k & 1 of A (the folded k runs (re,im) fastest); output slots 2n, 2n+1 =
                                // re, im of C: C.re = A.re B.re - A.im B.im, C.im = A.re B.im + A.im B.re
                                const int pp = k & 1;
                                const int64_t b0 = bk - pp + bn; // B(k,n).re
                                for (int p = 0; p < 2; ++p) {
                                    const int64_t bo = b0 + (p ^ pp);
                                    P.rg_w.push_back({slot, g.nout + 2 * n + p, (pp == 1 && p == 0) ? ~bo : bo});
                                }
                            }
                        }
                    }
                    for (int32_t n = 0; n < cb.N; ++n) {
                        if (cc) { for (int p = 0; p < 2; ++p) P.rg_out.push_back(cb.c_off + p + (int64_t)n * cb.c_ns); }
                        else P.rg_out.push_back(cb.c_off + (n & cb.c_nmask) + (int64_t)(n >> cb.c_nshift) * cb.c_ns);
                    }
                    g.nout += cb.N * out_per_n;
                    ++ci;
                }
                g.w_count = (int32_t)P.rg_w.size() - g.w_begin;
                const int32_t gi = (int32_t)P.rgroups.size();
                P.rgroups.push_back(g);
                for (int64_t r0 = 0; r0 < L; r0 += kRowGroupRows)
                    P.rg_items.push_back({gi, (int32_t)r0, (int32_t)std::min<int64_t>(kRowGroupRows, L - r0), 0});
            }
        }
    }
    PLAN_PHASE(3);
    // ---- tile items: one in-order queue, tail cut into shrinking pieces (guided self-scheduling) ------------------
    // The kernel's CTAs pull items from the head of this list through an atomic counter (kernels_gemm.cu), so the
    // assignment of items to CTAs is decided at run time by whoever is free. What the planner fixes is the ORDER
    // (C-block order, n0 outer / m0 inner: the ~148 items in flight at any time belong to a few neighbouring C blocks
    // whose operand panels are shared through L2) and the GRANULARITY: an item is a whole tile as long as a tile costs
    // less than the work still to be handed out divided by the grid width; after that tiles are cut at K-chunk
    // boundaries into pieces of (remaining work / grid width), never shorter than kMinPiece chunks, so the last CTAs
    // finish within one small piece of each other. The cycle model only sets piece sizes — an error in it costs
    // balance in proportion to the smallest pieces, not to the whole share of a CTA as with a static partition.
    // A cut tile writes partial sums to workspace slots which bsc_splitk_reduce_kernel adds in piece order
    // (deterministic, independent of which CTA ran which piece).
    {
        struct Proto { int32_t c, m0, n0, f; int64_t nch; double w; int np; };
        std::vector<Proto> protos;
        double total = 0;
        for (auto& tc : tile_cblks) {
            const int32_t c = tc.first; const int f = tc.second;
            const ItbCBlk& cb = P.cblks[c];
            const int TM = kTileM[f], TN = kTileN[f];
            const int64_t nch = chunks_of(cb);
            // a C block has at most four distinct tile shapes (full, m-edge, n-edge, corner): model each once
            const int64_t rm = cb.M % TM, rn = cb.N % TN;
            const double w_full = chunk_cycles(f, TM, TN), w_me = rm ? chunk_cycles(f, rm, TN) : w_full, w_ne = rn ? chunk_cycles(f, TM, rn) : w_full,
                         w_c = (rm && rn) ? chunk_cycles(f, rm, rn) : (rm ? w_me : w_ne);
            const int np = cb.pair_end - cb.pair_begin;
            const double ovh = kTileOverhead[f] + kPairOverhead * np;
            for (int32_t n0 = 0; n0 < cb.N; n0 += TN)
                for (int32_t m0 = 0; m0 < cb.M; m0 += TM) {
                    const bool me = m0 + TM > cb.M, ne = n0 + TN > cb.N;
                    const double w = me ? (ne ? w_c : w_me) : (ne ? w_ne : w_full);
                    protos.push_back({c, m0, n0, f, nch, w, np});
                    total += w * (double)nch + ovh;
                }
        }
        // measured correction of the cycle model (itb_contract_plan_refine): one factor per tile on its per-chunk cost
        if (P.tile_scale.size() == protos.size()) {
            total = 0;
            for (size_t i = 0; i < protos.size(); ++i) {
                protos[i].w *= P.tile_scale[i];
                total += protos[i].w * (double)protos[i].nch + kTileOverhead[protos[i].f] + kPairOverhead * protos[i].np;
            }
        } else
            P.tile_scale.clear();
        P.item_tile.clear();
        const int G = kNumSMs;
        // Hybrid schedule. STATIC part: the first kStaticFrac of the modelled work is cut stream-K fashion into G
        // contiguous ranges of (tile, K-chunk) space, one per CTA — at most G-1 tiles are cut there, which keeps the
        // split-K workspace traffic (one 128 KB partial per piece, written and read back by the reduce kernel) at its
        // minimum where tiles are few and K loops long (the *R step: 230 tiles of ~150 chunks for 148 CTAs). DYNAMIC part:
        // the rest of the list is a shared queue of pieces of geometrically shrinking cost (guided self-scheduling) that
        // the CTAs pull through an atomic head once their static range is done; it absorbs whatever the cycle model got
        // wrong in the static part and lets the CTAs finish within one small piece of each other.
        // cta_begin: G+2 entries — static range of CTA b = [cta_begin[b], cta_begin[b+1]), shared queue =
        // [cta_begin[G], cta_begin[G+1]).
        std::vector<double>& item_cost = P.item_cost;
        item_cost.clear();
        struct Piece { int64_t c0, c1; };
        size_t first_item_of_tile = 0;
        std::vector<Piece> cur_pieces; // pieces of the tile being emitted (consecutive items)
        auto flush_tile = [&](const Proto& t) { // assign workspace slots once the number of pieces of a tile is known
            if (cur_pieces.size() > 1) {
                P.splits.push_back({t.c, t.m0, t.n0, t.f, (int32_t)P.ws_slots, (int32_t)cur_pieces.size(), {0, 0}});
                for (size_t q = 0; q < cur_pieces.size(); ++q) {
                    P.tiles[first_item_of_tile + q].ws_slot = (int32_t)P.ws_slots++;
                    P.tiles[first_item_of_tile + q].split = (int32_t)P.splits.size() - 1;
                }
            }
            cur_pieces.clear();
        };
        auto emit = [&](const Proto& t, int64_t c0, int64_t c1) {
            if (cur_pieces.empty()) first_item_of_tile = P.tiles.size();
            cur_pieces.push_back({c0, c1});
            P.tiles.push_back({t.c, t.m0, t.n0, t.f, (int32_t)c0, (int32_t)c1, -1, -1});
            P.item_tile.push_back((int32_t)(&t - protos.data()));
            item_cost.push_back((double)(c1 - c0) * t.w + kTileOverhead[t.f] + kPairOverhead * std::ceil((double)t.np * (double)(c1 - c0) / (double)t.nch));
        };
        P.cta_begin.assign(G + 2, 0);
        size_t ti = 0;      // current tile
        int64_t coff = 0;   // chunks of it already emitted
        if (kSchedStreamK) {
            // Pure static stream-K partition (the round-1 schedule, kept selectable: ITB_SCHED=streamk): every CTA gets the
            // same modelled cycles, the tile list is cut at K-chunk boundaries where a share is full (<= G-1 cut tiles), no
            // shared queue. Its constants were fitted to per-CTA clock64 spans of this kernel on the bench workload.
            const double pair0 = kPairOverhead;
            double tot = 0;
            for (auto& t : protos) tot += t.w * (double)t.nch + kTileOverhead[t.f] + pair0 * t.np;
            double target = tot / G, assigned = 0;
            int b = 0; double load = 0;
            auto close_cta = [&]() {
                if (b < G - 1) { ++b; P.cta_begin[b] = (int32_t)P.tiles.size(); load = 0; target = std::max(0.0, tot - assigned) / (G - b); }
            };
            for (auto& t : protos) {
                auto piece_ovh = [&](int64_t take) { return kTileOverhead[t.f] + pair0 * std::ceil((double)t.np * (double)take / (double)t.nch); };
                int64_t c0 = 0;
                while (c0 < t.nch) {
                    const int64_t rem = t.nch - c0;
                    const double space = target - load - piece_ovh(t.nch - c0);
                    const int64_t fit = (int64_t)std::floor(space / t.w);
                    int64_t take;
                    if (b == G - 1 || fit >= rem) take = rem;
                    else if (fit >= kMinPiece && rem - fit >= kMinPiece) take = fit;
                    else {
                        const int64_t x = (rem - std::max<int64_t>(fit, 0) < kMinPiece || rem < 2 * kMinPiece) ? rem : kMinPiece;
                        const double over = (double)x * t.w - space, under = target - load;
                        if (load > 0 && under <= over) { close_cta(); continue; }
                        take = x;
                    }
                    if (take <= 0) { close_cta(); continue; }
                    if (!cur_pieces.empty()) tot += kTileOverhead[t.f] + pair0;
                    emit(t, c0, c0 + take);
                    load += (double)take * t.w + piece_ovh(take);
                    assigned += (double)take * t.w + piece_ovh(take);
                    c0 += take;
                    if (load >= target - 0.5 * t.w) close_cta();
                }
                flush_tile(t);
            }
            for (int g = b + 1; g <= G; ++g) P.cta_begin[g] = (int32_t)P.tiles.size();
            ti = protos.size();
        }
        const bool hybrid = !kSchedStreamK && kStaticFrac > 0 && total / G >= 40.0 * 4700.0 && protos.size() >= (size_t)G / 2;
        double remaining = total;
        if (hybrid) {
            const double share = kStaticFrac * total / G;
            for (int b = 0; b < G; ++b) {
                P.cta_begin[b] = (int32_t)P.tiles.size();
                double load = 0;
                while (ti < protos.size() && load < share) {
                    const Proto& t = protos[ti];
                    const double ovh = kTileOverhead[t.f] + kPairOverhead * t.np;
                    const int64_t rem = t.nch - coff;
                    const double cost_rem = (double)rem * t.w + ovh, space = share - load;
                    if (cost_rem <= space * 1.05) { // the rest of the tile fits
                        emit(t, coff, t.nch);
                        flush_tile(t);
                        load += cost_rem; remaining -= cost_rem;
                        ++ti; coff = 0;
                        continue;
                    }
                    const int64_t fit = (int64_t)std::floor((space - ovh) / t.w);
                    if (fit >= kMinPiece && rem - fit >= kMinPiece) { // cut here: a viable piece stays behind
                        emit(t, coff, coff + fit);
                        remaining -= (double)fit * t.w + ovh;
                        coff += fit;
                    } else if (load == 0) { // nothing fits an empty range: take the rest of the tile anyway
                        emit(t, coff, t.nch);
                        flush_tile(t);
                        remaining -= cost_rem;
                        ++ti; coff = 0;
                    }
                    break; // range closed
                }
            }
        }
        if (!kSchedStreamK) P.cta_begin[G] = (int32_t)P.tiles.size();
        // shared queue: what is left, cut by the guided rule
        for (; ti < protos.size(); ++ti, coff = 0) {
            const Proto& t = protos[ti];
            const double ovh = kTileOverhead[t.f] + kPairOverhead * t.np;
            const int64_t rem = t.nch - coff;
            const double cost = t.w * (double)rem + ovh;
            const double want = std::max(remaining / (kGuidedFactor * G), (double)kMinPiece * t.w);
            int64_t npieces = 1;
            if (cost > 1.25 * want) npieces = std::min<int64_t>((int64_t)std::ceil(cost / want), std::max<int64_t>(1, rem / kMinPiece));
            for (int64_t q = 0; q < npieces; ++q) emit(t, coff + rem * q / npieces, coff + rem * (q + 1) / npieces); // as equal as integers allow
            flush_tile(t);
            remaining -= cost;
        }
        P.cta_begin[G + 1] = (int32_t)P.tiles.size();
        PLAN_PHASE(4);
        // flattened device records of the dynamic-queue kernel (160 bytes per item: only built when that kernel will run —
        // the static kernel reads the 32-byte tile records themselves)
        static const bool ring_forced = [] { const char* e = getenv("ITB_TILE_KERNEL"); return e && std::string(e) == "ring"; }();
        P.qitems.resize((!kSchedStreamK || ring_forced) ? P.tiles.size() : 0);
        for (size_t i = 0; i < P.qitems.size(); ++i) {
            ItbQItem& q = P.qitems[i];
            std::memset(&q, 0, sizeof(q));
            q.tile = P.tiles[i];
            q.cb = P.cblks[q.tile.cblk];
            for (int32_t pp = q.cb.pair_begin; pp < q.cb.pair_end && pp - q.cb.pair_begin < ITB_QPAIRS; ++pp) {
                q.pK[pp - q.cb.pair_begin] = P.pairs[pp].K;
                q.pflags[pp - q.cb.pair_begin] = P.pairs[pp].flags;
            }
        }
    }
    std::stable_sort(P.skinny.begin(), P.skinny.end(), [&](const ItbSkinny& x, const ItbSkinny& y) {
        return P.cblks[x.cblk].ksum * (P.cblks[x.cblk].M + P.cblks[x.cblk].N) > P.cblks[y.cblk].ksum * (P.cblks[y.cblk].M + P.cblks[y.cblk].N);
    });
    P.table_bytes = (int64_t)(P.rgroups.size() * sizeof(ItbRowGroup) + P.rg_in.size() * sizeof(ItbRgIn) + P.rg_out.size() * 8 +
                              P.rg_w.size() * sizeof(ItbRgW) + P.rg_items.size() * sizeof(ItbRgItem)) +
                    (int64_t)(P.pairs.size() * sizeof(ItbPair) + P.cblks.size() * sizeof(ItbCBlk) +
                              (P.skinny.size() + P.skinny_q4.size() + P.skinny_q8.size()) * sizeof(ItbSkinny) + P.dots.size() * sizeof(ItbDot) +
                              P.dot_outs.size() * sizeof(ItbDotOut) + P.qitems.size() * sizeof(ItbQItem) +
                              P.splits.size() * sizeof(ItbSplitOut) + P.cta_begin.size() * sizeof(int32_t));
    PLAN_PHASE(5);
    P.tables_built = true;
    return ITB_OK;
}

// ---- strided block copies (shared by permute plans and the generic block-copy plans) -------------------
// One source block -> one destination (sub-)block, arbitrary strides on both sides. Dims are put in
// destination order (smallest dst stride first), unit extents dropped, neighbours fused when contiguous in
// BOTH tensors. If the src-fastest dim differs from the dst-fastest one the block goes to the shared-memory
// transposing path (one record per tile), else to the direct path (one record per 4096-element chunk).
struct CopyDim { int64_t ext, sstr, dstr; };

static int add_copy_block(itb_permute_plan& P, int64_t s_off, int64_t d_off, std::vector<CopyDim> in, int cs, int promote) {
    std::vector<CopyDim> dims;
    std::stable_sort(in.begin(), in.end(), [](const CopyDim& x, const CopyDim& y) { return x.dstr < y.dstr; });
    int64_t nelem = 1;
    for (auto& d : in) {
        nelem *= d.ext;
        if (d.ext == 1) continue;
        if (!dims.empty() && d.sstr == dims.back().sstr * dims.back().ext && d.dstr == dims.back().dstr * dims.back().ext) {
            dims.back().ext *= d.ext;
            continue;
        }
        dims.push_back(d);
    }
    if (nelem == 0) return ITB_OK;
    if (dims.empty()) dims.push_back({1, 1, 1});
    if ((int)dims.size() > ITB_MAXG) { set_error("permute: more than ITB_MAXG non-fusable indices"); return ITB_ERR_UNSUPPORTED; }
    ItbPermBlk pb;
    std::memset(&pb, 0, sizeof(pb));
    pb.s_off = s_off;
    pb.d_off = d_off;
    pb.n = (int32_t)dims.size();
    pb.nelem = nelem;
    pb.cs = cs;
    pb.promote = promote;
    for (int d = 0; d < ITB_MAXG; ++d) pb.ext[d] = 1;
    int tdim = 0;
    for (size_t d = 0; d < dims.size(); ++d) {
        if (dims[d].ext >= (1ll << 31)) { set_error("permute: fused extent exceeds 2^31"); return ITB_ERR_UNSUPPORTED; }
        pb.ext[d] = (int32_t)dims[d].ext;
        pb.sstr[d] = dims[d].sstr;
        pb.dstr[d] = dims[d].dstr;
        if (dims[d].sstr == 1) tdim = (int)d;
    }
    if (pb.dstr[0] != 1) tdim = 0; // destination not unit-stride along its fastest dim: direct path only
    pb.tdim = tdim;
    if (tdim == 0) {
        pb.item_begin = P.items_copy;
        const int64_t nch = (nelem + kPermCopyChunk - 1) / kPermCopyChunk;
        for (int64_t c = 0; c < nch; ++c) P.chunk_items.push_back({(int32_t)P.blks_copy.size(), 0, c * kPermCopyChunk});
        P.items_copy += nch;
        P.blks_copy.push_back(pb);
    } else {
        const int tile = (cs == 2) ? 32 : 64; // must match perm_tile_kernel's PT (kernels_permute.cu)
        pb.tiles0 = (pb.ext[0] + tile - 1) / tile;
        pb.tilesT = (pb.ext[tdim] + tile - 1) / tile;
        int64_t rest = 1;
        for (int d = 1; d < pb.n; ++d) if (d != tdim) rest *= pb.ext[d];
        pb.item_begin = P.items_tiled;
        // enumerate tiles: dim-0 tiles fastest, then dim-T tiles, then the remaining dims (odometer)
        std::vector<int64_t> idx(pb.n, 0);
        for (int64_t rr = 0; rr < rest; ++rr) {
            int64_t bs = 0, bd = 0;
            for (int d = 1; d < pb.n; ++d) if (d != tdim) { bs += idx[d] * pb.sstr[d]; bd += idx[d] * pb.dstr[d]; }
            for (int32_t tT = 0; tT < pb.tilesT; ++tT)
                for (int32_t t0 = 0; t0 < pb.tiles0; ++t0) {
                    ItbPermTile it;
                    it.s_base = pb.s_off + bs + (int64_t)tT * tile * pb.sstr[tdim] + (int64_t)t0 * tile * pb.sstr[0];
                    it.d_base = pb.d_off + bd + (int64_t)tT * tile * pb.dstr[tdim] + (int64_t)t0 * tile * pb.dstr[0];
                    it.ss0 = pb.sstr[0];
                    it.dsT = pb.dstr[tdim];
                    it.n0 = std::min<int32_t>(tile, pb.ext[0] - t0 * tile);
                    it.nT = std::min<int32_t>(tile, pb.ext[tdim] - tT * tile);
                    P.tile_items.push_back(it);
                }
            for (int d = 1; d < pb.n; ++d) {
                if (d == tdim) continue;
                if (++idx[d] < pb.ext[d]) break;
                idx[d] = 0;
            }
        }
        P.items_tiled += (int64_t)pb.tiles0 * pb.tilesT * rest;
        P.blks_tiled.push_back(pb);
    }
    return ITB_OK;
}

// ---- permute -------------------------------------------------------------------------------------
int build_permute_plan(itb_permute_plan& P) {
    const TensorStruct &S = P.S, &D = P.D;
    const int r = S.order;
    if (D.order != r) { set_error("permute: order mismatch"); return ITB_ERR_INVALID; }
    std::vector<int> inv(r, -1);
    for (int i = 0; i < r; ++i) {
        int d = P.perm[i];
        if (d < 0 || d >= r || inv[d] >= 0) { set_error("permute: perm is not a permutation"); return ITB_ERR_INVALID; }
        inv[d] = i;
    }
    for (int i = 0; i < r; ++i) {
        int d = P.perm[i];
        if (S.nsect[i] != D.nsect[d]) { set_error("permute: sector count mismatch"); return ITB_ERR_INVALID; }
        for (int s = 0; s < S.nsect[i]; ++s)
            if (S.ext(i, s) != D.ext(d, s)) { set_error("permute: sector size mismatch"); return ITB_ERR_INVALID; }
    }
    if (S.dtype == ITB_C64 && D.dtype == ITB_F64) { set_error("permute: cannot demote complex to real"); return ITB_ERR_INVALID; }
    const int promote = (S.dtype == ITB_F64 && D.dtype == ITB_C64) ? 1 : 0;
    const int cs = (S.dtype == ITB_C64) ? 2 : 1; // complex pairs move as 16-byte units

    std::unordered_map<std::vector<int32_t>, int64_t, VecHash> dpos;
    std::vector<int32_t> key(r);
    for (int64_t b = 0; b < D.nblocks; ++b) {
        for (int j = 0; j < r; ++j) key[j] = D.block(b)[j];
        dpos[key] = b;
    }
    P.blks_copy.clear(); P.blks_tiled.clear(); P.chunk_items.clear(); P.tile_items.clear();
    P.items_copy = P.items_tiled = 0;
    P.bytes = 0;
    std::vector<char> hit(D.nblocks, 0);
    std::vector<int64_t> ss(r), ds(r);
    for (int64_t b = 0; b < S.nblocks; ++b) {
        const int32_t* sb = S.block(b);
        for (int i = 0; i < r; ++i) key[P.perm[i]] = sb[i];
        auto it = dpos.find(key);
        if (it == dpos.end()) { set_error("permute: destination lacks the image of a source block"); return ITB_ERR_INVALID; }
        const int64_t db = it->second;
        if (hit[db]) { set_error("permute: two source blocks map to one destination block"); return ITB_ERR_INVALID; }
        hit[db] = 1;
        int64_t s = 1;
        for (int i = 0; i < r; ++i) { ss[i] = s; s *= S.ext(i, sb[i]); }
        const int64_t nelem = s;
        s = 1;
        for (int j = 0; j < r; ++j) { ds[j] = s; s *= D.ext(j, key[j]); }
        std::vector<CopyDim> dims;
        for (int j = 0; j < r; ++j) dims.push_back({D.ext(j, key[j]), ss[inv[j]], ds[j]});
        int rc2 = add_copy_block(P, S.offsets[b], D.offsets[db], dims, cs, promote);
        if (rc2 != ITB_OK) return rc2;
        P.bytes += nelem * 8 * ((S.dtype == ITB_C64 ? 2 : 1) + (D.dtype == ITB_C64 ? 2 : 1));
    }
    P.need_zero = false;
    P.zero_ranges.clear();
    for (int64_t b = 0; b < D.nblocks; ++b) {
        if (hit[b]) continue;
        P.need_zero = true;
        int64_t sz = 1;
        for (int j = 0; j < r; ++j) sz *= D.ext(j, D.block(b)[j]);
        if (!P.zero_ranges.empty() && P.zero_ranges[P.zero_ranges.size() - 2] + P.zero_ranges.back() == D.offsets[b]) P.zero_ranges.back() += sz;
        else { P.zero_ranges.push_back(D.offsets[b]); P.zero_ranges.push_back(sz); }
    }
    // Small fill-ins ride the transposing kernel's item list as zero-fill items (no extra launches); a destination
    // that is mostly fill-in keeps the memset path (need_zero stays set).
    {
        int64_t zero_elems = 0;
        for (size_t i = 1; i < P.zero_ranges.size(); i += 2) zero_elems += P.zero_ranges[i];
        if (P.need_zero && P.items_tiled > 0 && zero_elems <= (int64_t)4096 * 256) {
            for (size_t i = 0; i + 1 < P.zero_ranges.size(); i += 2)
                for (int64_t e = 0; e < P.zero_ranges[i + 1]; e += 4096) {
                    ItbPermTile it;
                    std::memset(&it, 0, sizeof(it));
                    it.d_base = P.zero_ranges[i] + e;
                    it.n0 = (int32_t)std::min<int64_t>(4096, P.zero_ranges[i + 1] - e);
                    it.nT = -1;
                    P.tile_items.push_back(it);
                    ++P.items_tiled;
                }
            P.need_zero = false;
            P.zero_in_items = true;
        }
    }
    return ITB_OK;
}

} // namespace itb

// ---- C ABI: host-only entry points -----------------------------------------------------------------
using namespace itb;

extern "C" {

const char* itb_last_error(void) { return itb::last_error_cstr(); }

int itb_contract_plan_create(const itb_tensor_desc* A, const int32_t* labA, const itb_tensor_desc* B,
                             const int32_t* labB, itb_contract_plan** out) {
    if (!out) { set_error("plan_create: null out"); return ITB_ERR_INVALID; }
    *out = nullptr;
    auto* P = new itb_contract_plan();
    int rc = parse_desc(A, P->A, "contract A");
    if (rc == ITB_OK) rc = parse_desc(B, P->B, "contract B");
    if (rc == ITB_OK) {
        P->labA.assign(labA, labA + A->order);
        P->labB.assign(labB, labB + B->order);
        rc = build_contract_plan(*P);
    }
    // the device tables are built on first need (run, info, introspection, slicing): a caller that first looks at the
    // result structure and then restricts the plan to a row slice pays for ONE table build, not two
    if (rc != ITB_OK) { delete P; return rc; }
    *out = P;
    return ITB_OK;
}

static int ensure_tables(const itb_contract_plan* P) {
    if (P->tables_built) return ITB_OK;
    return build_contract_tables(*const_cast<itb_contract_plan*>(P));
}

int itb_contract_plan_shape(const itb_contract_plan* P, int32_t* c_order, int32_t* c_dtype, int64_t* c_nblocks, int64_t* c_nelems,
                            int64_t* npairs, double* flops) {
    if (!P) { set_error("plan_shape: null"); return ITB_ERR_INVALID; }
    if (c_order) *c_order = P->C.order;
    if (c_dtype) *c_dtype = P->C.dtype;
    if (c_nblocks) *c_nblocks = P->C.nblocks;
    if (c_nelems) *c_nelems = P->C.nelems;
    if (npairs) *npairs = (int64_t)P->triples.size() / 3;
    if (flops) *flops = P->flops;
    return ITB_OK;
}

void itb_contract_plan_release_device(itb_contract_plan* plan); // api.cu
void itb_permute_plan_release_device(itb_permute_plan* plan);   // api.cu

int itb_contract_plan_destroy(itb_contract_plan* plan) {
    if (!plan) return ITB_OK;
    itb_contract_plan_release_device(plan);
    delete plan;
    return ITB_OK;
}

int itb_contract_plan_info(const itb_contract_plan* P, itb_contract_info* o) {
    if (!P || !o) { set_error("plan_info: null"); return ITB_ERR_INVALID; }
    { int rc = ensure_tables(P); if (rc != ITB_OK) return rc; }
    o->c_order = P->C.order;
    o->c_dtype = P->C.dtype;
    o->c_nblocks = P->C.nblocks;
    o->c_nelems = P->C.nelems;
    o->npairs = (int64_t)P->triples.size() / 3;
    o->flops = P->flops;
    o->n_gemm_tiles = (int64_t)P->tiles.size();
    o->n_skinny = (int64_t)(P->skinny.size() + P->skinny_q4.size() + P->skinny_q8.size() + P->rg_items.size());
    o->n_dot = (int64_t)P->dots.size();
    o->table_bytes = P->table_bytes;
    for (int i = 0; i < 5; ++i) o->class_flops[i] = P->class_flops[i];
    return ITB_OK;
}
int itb_contract_plan_c_labels(const itb_contract_plan* P, int32_t* v) { std::copy(P->labC.begin(), P->labC.end(), v); return ITB_OK; }
int itb_contract_plan_c_nsect(const itb_contract_plan* P, int32_t* v) { std::copy(P->C.nsect.begin(), P->C.nsect.end(), v); return ITB_OK; }
int itb_contract_plan_c_sect(const itb_contract_plan* P, int64_t* v) { std::copy(P->C.sect.begin(), P->C.sect.end(), v); return ITB_OK; }
int itb_contract_plan_c_blocks(const itb_contract_plan* P, int32_t* v) { std::copy(P->C.blocks.begin(), P->C.blocks.end(), v); return ITB_OK; }
int itb_contract_plan_c_offsets(const itb_contract_plan* P, int64_t* v) { std::copy(P->C.offsets.begin(), P->C.offsets.end(), v); return ITB_OK; }
int itb_contract_plan_pairs(const itb_contract_plan* P, int64_t* v) { std::copy(P->triples.begin(), P->triples.end(), v); return ITB_OK; }

int64_t itb_contract_plan_tiles(const itb_contract_plan* P, int32_t* out, int64_t cap) {
    if (!P) return ITB_ERR_INVALID;
    if (ensure_tables(P) != ITB_OK) return ITB_ERR_INVALID;
    for (int64_t i = 0; out && i < (int64_t)P->tiles.size() && i < cap; ++i) {
        const ItbTile& t = P->tiles[i];
        const int32_t v[8] = {t.cblk, t.m0, t.n0, kTileM[t.cfg], kTileN[t.cfg], t.chunk_begin, t.chunk_end, t.ws_slot};
        std::copy(v, v + 8, out + 8 * i);
    }
    return (int64_t)P->tiles.size();
}
int64_t itb_contract_plan_cta_begin(const itb_contract_plan* P, int32_t* out, int64_t cap) {
    if (!P) return ITB_ERR_INVALID;
    if (ensure_tables(P) != ITB_OK) return ITB_ERR_INVALID;
    for (int64_t i = 0; out && i < (int64_t)P->cta_begin.size() && i < cap; ++i) out[i] = P->cta_begin[i];
    return (int64_t)P->cta_begin.size();
}
int64_t itb_contract_plan_rowgroups(const itb_contract_plan* P, int64_t* out, int64_t cap) {
    if (!P) return ITB_ERR_INVALID;
    if (ensure_tables(P) != ITB_OK) return ITB_ERR_INVALID;
    for (int64_t i = 0; out && i < (int64_t)P->rgroups.size() && i < cap; ++i) {
        const ItbRowGroup& g = P->rgroups[i];
        const int64_t v[4] = {g.nin, g.nout, g.nL, g.L};
        std::copy(v, v + 4, out + 4 * i);
    }
    return (int64_t)P->rgroups.size();
}
int64_t itb_contract_plan_cblks(const itb_contract_plan* P, int64_t* out, int64_t cap) {
    if (!P) return ITB_ERR_INVALID;
    if (ensure_tables(P) != ITB_OK) return ITB_ERR_INVALID;
    for (int64_t i = 0; out && i < (int64_t)P->cblks.size() && i < cap; ++i) {
        const ItbCBlk& c = P->cblks[i];
        const int64_t v[4] = {c.M, c.N, c.ksum, c.pair_end - c.pair_begin};
        std::copy(v, v + 4, out + 4 * i);
    }
    return (int64_t)P->cblks.size();
}

int itb_contract_plan_set_cblock_range(itb_contract_plan* P, int64_t first, int64_t last) {
    if (!P) { set_error("set_cblock_range: null"); return ITB_ERR_INVALID; }
    if (first < 0 || (last >= 0 && last < first) || last > P->C.nblocks) { set_error("set_cblock_range: bad range"); return ITB_ERR_INVALID; }
    P->cb_first = first;
    P->cb_last = last;
    itb_contract_plan_release_device(P);
    return build_contract_tables(*P);
}

int itb_contract_plan_set_cblock_mask(itb_contract_plan* P, const uint8_t* mask) {
    if (!P) { set_error("set_cblock_mask: null"); return ITB_ERR_INVALID; }
    if (mask) P->cb_mask.assign(mask, mask + P->C.nblocks);
    else P->cb_mask.clear();
    itb_contract_plan_release_device(P);
    return build_contract_tables(*P);
}

int itb_contract_plan_cblock_flops(const itb_contract_plan* P, double* out) {
    if (!P || !out) { set_error("cblock_flops: null"); return ITB_ERR_INVALID; }
    const TensorStruct &A = P->A, &B = P->B;
    for (int64_t c = 0; c < P->C.nblocks; ++c) out[c] = 0;
    std::vector<char> contA(A.order, 0), contB(B.order, 0);
    for (int i = 0; i < A.order; ++i)
        for (int j = 0; j < B.order; ++j)
            if (P->labA[i] == P->labB[j] && !contB[j]) { contA[i] = 1; contB[j] = 1; break; }
    const double cmul = (A.dtype == ITB_C64 ? 2.0 : 1.0) * (B.dtype == ITB_C64 ? 2.0 : 1.0);
    for (size_t p = 0; p + 2 < P->triples.size() + 0; p += 3) {
        const int64_t ia = P->triples[p], ib = P->triples[p + 1], ic = P->triples[p + 2];
        double m = 1, n = 1, k = 1;
        for (int i = 0; i < A.order; ++i) { const double e = (double)A.ext(i, A.block(ia)[i]); if (contA[i]) k *= e; else m *= e; }
        for (int j = 0; j < B.order; ++j) if (!contB[j]) n *= (double)B.ext(j, B.block(ib)[j]);
        out[ic] += 2.0 * m * n * k * cmul;
    }
    return ITB_OK;
}

// Modelled device work of the plan as currently restricted (range / mask / row slices): cycles of the DMMA tile class summed over
// all CTAs (the quantity the stream-K partition divides by the grid width) and the algorithmic bytes of the streaming class.
// What a multi-GPU row partition balances instead of flops: a cut through a sector that leaves a short remainder tile costs
// almost a full tile per K-chunk, and the cycle model knows it.
int itb_contract_plan_model_work(const itb_contract_plan* P, double* tile_cycles, double* stream_bytes) {
    if (!P) { set_error("plan_model_work: null"); return ITB_ERR_INVALID; }
    if (ensure_tables(P) != ITB_OK) return ITB_ERR_INVALID;
    double cyc = 0;
    for (double c : P->item_cost) cyc += c;
    if (tile_cycles) *tile_cycles = cyc;
    if (stream_bytes) {
        double bytes = 0;
        auto add = [&](int32_t c) { const ItbCBlk& cb = P->cblks[c]; bytes += 8.0 * ((double)cb.M * cb.N + (double)std::max(cb.M, cb.N) * cb.ksum); };
        for (auto& it : P->skinny) if (it.row0 == 0) add(it.cblk);
        for (auto& it : P->skinny_q4) if (it.row0 == 0) add(it.cblk);
        for (auto& it : P->skinny_q8) if (it.row0 == 0) add(it.cblk);
        for (auto& g : P->rgroups) bytes += 8.0 * (double)g.L * ((double)g.nin + (double)g.nout);
        *stream_bytes = bytes;
    }
    return ITB_OK;
}

int itb_contract_plan_set_index_slices(itb_contract_plan* P, int32_t c_index, const int64_t* lo, const int64_t* hi) {
    if (!P) { set_error("set_index_slices: null"); return ITB_ERR_INVALID; }
    const int32_t old_index = P->slice_index;
    const std::vector<int64_t> old_lo = P->slice_lo, old_hi = P->slice_hi;
    if (!lo || !hi) { P->slice_index = -1; P->slice_lo.clear(); P->slice_hi.clear(); }
    else {
        if (c_index < 0 || c_index >= P->C.order) { set_error("set_index_slices: index out of range"); return ITB_ERR_INVALID; }
        const int32_t ns = P->C.nsect[c_index];
        for (int32_t q = 0; q < ns; ++q)
            if (lo[q] < 0 || hi[q] > P->C.ext(c_index, q)) { set_error("set_index_slices: range outside the sector"); return ITB_ERR_INVALID; }
        P->slice_index = c_index;
        P->slice_lo.assign(lo, lo + ns);
        P->slice_hi.assign(hi, hi + ns);
    }
    itb_contract_plan_release_device(P);
    P->tile_scale.clear(); // measured corrections belong to the tiles of the previous extent
    int rc = build_contract_tables(*P);
    if (rc != ITB_OK) { // leave the plan as it was
        const std::string msg = last_error_cstr();
        P->slice_index = old_index; P->slice_lo = old_lo; P->slice_hi = old_hi;
        build_contract_tables(*P);
        set_error(msg);
    }
    return rc;
}

int itb_permute_plan_create(const itb_tensor_desc* src, const itb_tensor_desc* dst, const int32_t* perm,
                            itb_permute_plan** out) {
    if (!out) { set_error("permute_plan_create: null out"); return ITB_ERR_INVALID; }
    *out = nullptr;
    auto* P = new itb_permute_plan();
    int rc = parse_desc(src, P->S, "permute src");
    if (rc == ITB_OK) rc = parse_desc(dst, P->D, "permute dst");
    if (rc == ITB_OK) {
        P->perm.assign(perm, perm + src->order);
        rc = build_permute_plan(*P);
    }
    if (rc != ITB_OK) { delete P; return rc; }
    *out = P;
    return ITB_OK;
}
int itb_blockcopy_plan_create(int64_t nitems, const itb_copy_item* items, int32_t src_dtype, int32_t dst_dtype,
                              itb_permute_plan** out) {
    if (!out || (nitems > 0 && !items)) { set_error("blockcopy_plan_create: null"); return ITB_ERR_INVALID; }
    *out = nullptr;
    if (src_dtype == ITB_C64 && dst_dtype == ITB_F64) { set_error("blockcopy: cannot demote complex to real"); return ITB_ERR_INVALID; }
    auto* P = new itb_permute_plan();
    P->S.dtype = src_dtype;
    P->D.dtype = dst_dtype;
    P->S.nblocks = nitems; // (only used to skip empty plans)
    const int cs = src_dtype == ITB_C64 ? 2 : 1;
    const int promote = (src_dtype == ITB_F64 && dst_dtype == ITB_C64) ? 1 : 0;
    for (int64_t i = 0; i < nitems; ++i) {
        const itb_copy_item& it = items[i];
        if (it.n < 0 || it.n > ITB_MAX_ORDER) { delete P; set_error("blockcopy: bad item order"); return ITB_ERR_INVALID; }
        std::vector<CopyDim> dims;
        int64_t ne = 1;
        for (int d = 0; d < it.n; ++d) { dims.push_back({it.ext[d], it.sstr[d], it.dstr[d]}); ne *= it.ext[d]; }
        int rc = add_copy_block(*P, it.s_off, it.d_off, dims, cs, promote);
        if (rc != ITB_OK) { delete P; return rc; }
        P->bytes += ne * 8 * ((src_dtype == ITB_C64 ? 2 : 1) + (dst_dtype == ITB_C64 ? 2 : 1));
    }
    P->need_zero = false;
    *out = P;
    return ITB_OK;
}
int itb_permute_plan_destroy(itb_permute_plan* plan) {
    if (!plan) return ITB_OK;
    itb_permute_plan_release_device(plan);
    delete plan;
    return ITB_OK;
}
int64_t itb_permute_plan_bytes(const itb_permute_plan* P) { return P ? P->bytes : 0; }

// getBlockOffsets(IndexSet,QN) (itensor/itdata/qdense.cc:133-173): iterate every block with the
// FIRST index fastest (== reference-sorted order) and keep those whose flux matches. QN values
// follow QNum::set (itensor/qn.cc:10-29): modulo |mod| with non-negative representatives.
static inline int32_t qn_norm(int64_t v, int32_t mod) {
    int64_t m = mod < 0 ? -(int64_t)mod : mod;
    if (m > 1) { int64_t a = v < 0 ? -v : v; return (int32_t)((m * a + v) % m); }
    return (int32_t)v;
}
int64_t itb_flux_blocks(int32_t order, const int32_t* nsect, const int32_t* qn, int32_t nqn, const int32_t* mod,
                        const int32_t* dir, const int32_t* flux, int32_t* blocks, int64_t cap) {
    if (order < 0 || order > ITB_MAX_ORDER || nqn < 0 || nqn > 4) { set_error("flux_blocks: bad arguments"); return ITB_ERR_INVALID; }
    if (order == 0) return 1; // rank-0: single block with empty coordinates
    std::vector<int64_t> start(order + 1, 0);
    for (int j = 0; j < order; ++j) start[j + 1] = start[j] + nsect[j];
    std::vector<int32_t> I(order, 0);
    int64_t count = 0;
    while (true) {
        bool ok = true;
        for (int c = 0; c < nqn && ok; ++c) {
            // blockqn += J.qn(1+I[j])*J.dir()  — each term and each partial sum is normalised
            int32_t acc = qn_norm(0, mod[c]);
            for (int j = 0; j < order; ++j) {
                int32_t term = qn_norm((int64_t)qn_norm(qn[(start[j] + I[j]) * nqn + c], mod[c]) * dir[j], mod[c]);
                acc = qn_norm((int64_t)acc + term, mod[c]);
            }
            if (acc != qn_norm(flux[c], mod[c])) ok = false;
        }
        if (ok) {
            if (blocks && count < cap) std::copy(I.begin(), I.end(), blocks + count * order);
            ++count;
        }
        int j = 0;
        while (j < order) {
            if (++I[j] < nsect[j]) break;
            I[j] = 0;
            ++j;
        }
        if (j == order) break;
    }
    return count;
}

} // extern "C"
