// api.cu — device-facing half of the C ABI (include/itb200.h): context, stream-ordered memory pool,
// table upload and kernel launches. Host-only planning lives in plan.cc.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "plan.h"

namespace itb {
cudaError_t launch_gemm(const ItbQItem* items, int n_items, int* queue, const int32_t* cta_begin, int n_static_ctas, int grid,
                        const ItbSplitOut* souts, int nsouts,
                        const ItbCBlk* cblks, const ItbPair* pairs, const double* A, const double* B, double* C, double* ws,
                        long long* cta_cycles,
                        cudaStream_t st);
cudaError_t launch_gemm_static(const ItbTile* items, const int32_t* cta_begin, int grid, const ItbSplitOut* souts, int nsouts,
                               const ItbCBlk* cblks, const ItbPair* pairs, const double* A, const double* B, double* C, double* ws,
                               long long* cta_cycles, const ItbMirrors* mir, cudaStream_t st);
cudaError_t launch_skinny(const ItbSkinny* items, int n, const ItbSkinny* q4, int nq4, const ItbSkinny* q8, int nq8,
                          const ItbCBlk* cblks, const ItbPair* pairs, const double* A, const double* B, double* C, cudaStream_t st);
cudaError_t launch_dot(const ItbDot* items, int n, const ItbDotOut* outs, int nouts, const ItbCBlk* cblks, const ItbPair* pairs,
                       const double* A, const double* B, double* partial, double* C, cudaStream_t st);
cudaError_t launch_rowgroups(const ItbRgItem* items, int nitems, const ItbRowGroup* groups, const ItbRgIn* ins, const int64_t* outs,
                             const ItbRgW* wents, int max_nout, const double* A, const double* B, double* C, cudaStream_t st);
cudaError_t launch_peak(int which, int iters, double* out, int num_sms, cudaStream_t st);
cudaError_t launch_permute(int src_cplx, int dst_cplx, const ItbPermBlk* bc, const ItbPermChunk* chunks, int64_t items_c,
                           const ItbPermTile* tiles, int64_t items_t, const void* src, void* dst, double ar, double ai, int accum,
                           cudaStream_t st, int* launches);
cudaError_t launch_scal(int cplx, int64_t n, void* x, double ar, double ai, int sms, cudaStream_t st);
cudaError_t launch_axpy(int cplx, int64_t n, double ar, double ai, const void* x, void* y, int sms, cudaStream_t st);
cudaError_t launch_fill(int cplx, int64_t n, void* x, double re, double im, int sms, cudaStream_t st);
cudaError_t launch_conj(int64_t n, void* x, int sms, cudaStream_t st);
cudaError_t launch_r2c(int64_t n, const void* x, void* y, int sms, cudaStream_t st);
cudaError_t launch_part(int64_t n, const void* x, void* y, int imag, int sms, cudaStream_t st);
cudaError_t launch_ssq(int64_t nreal, const void* x, double scale, double* scratch, int* grid_out, int sms, cudaStream_t st);
cudaError_t launch_dot1(int cplx, int64_t n, const void* x, const void* y, int conj_x, double* scratch, int* grid_out, int sms,
                        cudaStream_t st);

struct DeviceTables {
    void* base = nullptr; // one pooled allocation holding every table of the plan
    size_t bytes = 0;
    const ItbPair* pairs = nullptr;
    const ItbCBlk* cblks = nullptr;
    const ItbQItem* qitems = nullptr;
    const ItbTile* tiles = nullptr;
    const int32_t* cta_begin = nullptr;
    const ItbSplitOut* splits = nullptr;
    const ItbRowGroup* rgroups = nullptr;
    const ItbRgIn* rg_in = nullptr;
    const int64_t* rg_out = nullptr;
    const ItbRgW* rg_w = nullptr;
    const ItbRgItem* rg_items = nullptr;
    int rg_max_nout = 0;
    const ItbSkinny* skinny = nullptr;
    const ItbSkinny* skinny_q4 = nullptr;
    const ItbSkinny* skinny_q8 = nullptr;
    const ItbDot* dots = nullptr;
    const ItbDotOut* dot_outs = nullptr;
    double* dot_partial = nullptr;
    int* counters = nullptr; // work-queue head
    const ItbPermBlk* pcopy = nullptr;
    const ItbPermChunk* pchunks = nullptr;
    const ItbPermTile* ptiles = nullptr;
};
} // namespace itb

using namespace itb;

struct itb_solver;
struct itb_svd_batch;
struct itb_eigh_batch;
extern "C" {
int itb_solver_eigh_batch_run(itb_solver* s, int32_t dtype, int64_t nblocks, const int64_t* a_off, const int32_t* n, const void* dA,
                              int negate, itb_eigh_batch** out);
}
extern "C" {
int itb_solver_create(void* stream, itb_solver** out);
int itb_solver_set_stream(itb_solver* s, void* stream);
int itb_solver_destroy(itb_solver* s);
}
void itb_warm_library_pages(); // solver.cu: background read of the cuSOLVER/cuBLAS shared objects
extern "C" {
int itb_solver_syevd(itb_solver* s, int32_t dtype, int32_t n, void* hA, double* hW, int32_t* info);
int itb_solver_gesvd(itb_solver* s, int32_t dtype, int32_t m, int32_t n, void* hA, double* hS, void* hU, void* hVT, int32_t* info);
int itb_solver_svd_batch_run(itb_solver* s, int32_t dtype, int64_t nblocks, const int64_t* a_off, const int32_t* m, const int32_t* n,
                             const void* dA, itb_svd_batch** out);
}

struct itb_ctx {
    itb_solver* solver = nullptr;
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int64_t launches = 0;
    // caching pool: freed blocks are kept by rounded size; everything runs on one stream so reuse
    // is stream-ordered and needs no events
    std::multimap<size_t, void*> free_blocks;
    std::unordered_map<void*, size_t> live;
    size_t pooled_bytes = 0;
    // pinned host staging for table uploads: a ring of buffers, so that the host can plan and queue up to kStagingSlots
    // new structures ahead of the GPU (with ONE buffer every upload waited for the previous one, which sits in stream order
    // behind the previous contraction: host planning and device execution took turns instead of overlapping)
    static constexpr int kStagingSlots = 8;
    void* staging[kStagingSlots] = {};
    size_t staging_bytes[kStagingSlots] = {};
    int staging_next = 0;
    double* scratch = nullptr; // device scratch for reductions
    double* ws = nullptr;      // split-K workspace (grow-only, stream-ordered reuse)
    size_t ws_doubles = 0;
    size_t scratch_doubles = 0;
    double* h_result = nullptr; // pinned 4 doubles
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool profile = false;
    cudaStream_t aux = nullptr;          // side stream: small streaming / split-K-dot launches overlap the tile kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_staging[kStagingSlots] = {}; // completion of the last table upload out of each staging buffer
    bool staging_busy[kStagingSlots] = {};
    float last_ms[5] = {0, 0, 0, 0, 0};
    long long* d_cta_cycles = nullptr;   // profile mode: per-CTA clock64 span of the last tile-kernel launch
    std::vector<long long> h_cta_cycles;
    std::vector<long long> h_item_cycles;
    size_t cta_cycles_words = 0;
    cudaEvent_t pev[10] = {};
};

#define CUDA_TRY(expr)                                                                                    \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));                               \
            return ITB_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

static size_t round_size(size_t b) {
    if (b < 512) return 512;
    if (b < (1u << 20)) return (b + 511) & ~(size_t)511;
    return (b + ((1u << 20) - 1)) & ~(size_t)((1u << 20) - 1);
}

static int pool_alloc(itb_ctx* c, size_t bytes, void** out) {
    const size_t r = round_size(bytes);
    auto it = c->free_blocks.lower_bound(r);
    if (it != c->free_blocks.end() && it->first <= r + r / 4 + 4096) {
        *out = it->second;
        c->live[*out] = it->first;
        c->pooled_bytes -= it->first;
        c->free_blocks.erase(it);
        return ITB_OK;
    }
    // A miss goes to the driver's stream-ordered allocator (release threshold raised at context creation, so its memory
    // stays mapped): unlike cudaMalloc it neither synchronises the device nor costs 0.1-0.3 ms — in a DMRG ramp the block
    // sizes change at every bond and a third of all result allocations miss (75 us per Contract on average before).
    void* p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, r, c->stream);
    if (e != cudaSuccess) {
        // trim the cache and retry once
        for (auto& kv : c->free_blocks) cudaFree(kv.second);
        c->free_blocks.clear();
        c->pooled_bytes = 0;
        (void)cudaGetLastError();
        cudaStreamSynchronize(c->stream);
        e = cudaMallocAsync(&p, r, c->stream);
        if (e != cudaSuccess) {
            set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? ITB_ERR_NOMEM : ITB_ERR_CUDA;
        }
    }
    c->live[p] = r;
    *out = p;
    return ITB_OK;
}
static int pool_free(itb_ctx* c, void* p) {
    if (!p) return ITB_OK;
    auto it = c->live.find(p);
    if (it == c->live.end()) { set_error("itb_free: pointer not owned by this context"); return ITB_ERR_INVALID; }
    c->free_blocks.emplace(it->second, p);
    c->pooled_bytes += it->second;
    c->live.erase(it);
    return ITB_OK;
}

static int ensure_scratch(itb_ctx* c, size_t doubles) {
    if (c->scratch_doubles >= doubles) return ITB_OK;
    if (c->scratch) { CUDA_TRY(cudaStreamSynchronize(c->stream)); cudaFree(c->scratch); c->scratch = nullptr; }
    CUDA_TRY(cudaMalloc(&c->scratch, doubles * sizeof(double)));
    c->scratch_doubles = doubles;
    return ITB_OK;
}
static int ensure_ws(itb_ctx* c, size_t doubles) {
    if (c->ws_doubles >= doubles) return ITB_OK;
    if (c->ws) { CUDA_TRY(cudaStreamSynchronize(c->stream)); cudaFree(c->ws); c->ws = nullptr; }
    CUDA_TRY(cudaMalloc(&c->ws, doubles * sizeof(double)));
    c->ws_doubles = doubles;
    return ITB_OK;
}
static int ensure_staging(itb_ctx* c, int slot, size_t bytes) { // (the slot is idle: its last copy has completed)
    if (c->staging_bytes[slot] >= bytes) return ITB_OK;
    if (c->staging[slot]) { cudaFreeHost(c->staging[slot]); c->staging[slot] = nullptr; c->staging_bytes[slot] = 0; }
    size_t nb = std::max<size_t>(bytes, 256u << 10);
    CUDA_TRY(cudaMallocHost(&c->staging[slot], nb));
    c->staging_bytes[slot] = nb;
    return ITB_OK;
}

// pack host vectors into one staging buffer, one H2D copy, carve device pointers
struct Packer {
    std::vector<std::pair<const void*, size_t>> parts;
    std::vector<size_t> offs;
    size_t total = 0;
    size_t add(const void* p, size_t bytes) {
        const size_t o = total;
        parts.push_back({p, bytes});
        offs.push_back(o);
        total += (bytes + 255) & ~(size_t)255;
        return o;
    }
};

extern "C" {

const char* itb_version(void) { return "itb200 0.1 (sm_100a)"; }

int itb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

int itb_ctx_create(int device, itb_ctx** out) {
    if (!out) { set_error("ctx_create: null out"); return ITB_ERR_INVALID; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        set_error("itb_ctx_create: no usable CUDA device (this library has no CPU fallback)");
        return ITB_ERR_CUDA;
    }
    if (device < 0 || device >= n) { set_error("ctx_create: bad device ordinal"); return ITB_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(device));
    auto* c = new itb_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    if (prop.major < 10) {
        set_error("itb_ctx_create: kernels are built for sm_100a only; found compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor));
        delete c;
        return ITB_ERR_CUDA;
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {   // keep the stream-ordered allocator's memory mapped between uses (the context's own pool sits on top of it)
        cudaMemPool_t mp = nullptr;
        if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        (void)cudaGetLastError();
    }
    CUDA_TRY(cudaMallocHost(&c->h_result, 8 * sizeof(double)));
    CUDA_TRY(cudaEventCreate(&c->ev0));
    CUDA_TRY(cudaEventCreate(&c->ev1));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    for (auto& e : c->ev_staging) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    itb_warm_library_pages();
    *out = c;
    return ITB_OK;
}

int itb_ctx_destroy(itb_ctx* c) {
    if (!c) return ITB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& kv : c->free_blocks) cudaFree(kv.second);
    for (auto& kv : c->live) cudaFree(kv.first);
    if (c->scratch) cudaFree(c->scratch);
    if (c->ws) cudaFree(c->ws);
    for (void* p : c->staging) if (p) cudaFreeHost(p);
    if (c->h_result) cudaFreeHost(c->h_result);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->aux) cudaStreamDestroy(c->aux);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (auto e : c->ev_staging) if (e) cudaEventDestroy(e);
    for (auto& e : c->pev) if (e) cudaEventDestroy(e);
    if (c->d_cta_cycles) cudaFree(c->d_cta_cycles);
    if (c->solver) itb_solver_destroy(c->solver);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return ITB_OK;
}

void* itb_ctx_stream(itb_ctx* c) { return c ? (void*)c->stream : nullptr; }
int itb_ctx_device(itb_ctx* c) { return c ? c->device : -1; }
int itb_ctx_set_stream(itb_ctx* c, void* s) {
    if (!c) { set_error("set_stream: null ctx"); return ITB_ERR_INVALID; }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    // the lazily created solver (its cuSOLVER handle and the batched-SVD lanes' fence target) is bound to the stream
    if (c->solver) { int rc = itb_solver_set_stream(c->solver, s); if (rc != ITB_OK) return rc; }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)s;
    c->own_stream = false;
    for (bool& b : c->staging_busy) b = false; // the old stream was drained above
    return ITB_OK;
}
int itb_synchronize(itb_ctx* c) { CUDA_TRY(cudaStreamSynchronize(c->stream)); return ITB_OK; }
int64_t itb_launch_count(itb_ctx* c) { return c ? c->launches : 0; }

int itb_malloc(itb_ctx* c, size_t bytes, void** dptr) {
    if (!c || !dptr) { set_error("itb_malloc: null"); return ITB_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(c->device));
    return pool_alloc(c, bytes ? bytes : 1, dptr);
}
int itb_free(itb_ctx* c, void* p) { return pool_free(c, p); }
int itb_memcpy_h2d(itb_ctx* c, void* d, const void* s, size_t b) {
    if (b) CUDA_TRY(cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, c->stream));
    return ITB_OK;
}
int itb_memcpy_d2h(itb_ctx* c, void* d, const void* s, size_t b) {
    if (b) CUDA_TRY(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ITB_OK;
}
int itb_memcpy_d2d(itb_ctx* c, void* d, const void* s, size_t b) {
    if (b) CUDA_TRY(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToDevice, c->stream));
    return ITB_OK;
}
int itb_memset0(itb_ctx* c, void* d, size_t b) {
    if (b) CUDA_TRY(cudaMemsetAsync(d, 0, b, c->stream));
    return ITB_OK;
}
int itb_pool_trim(itb_ctx* c) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (auto& kv : c->free_blocks) cudaFree(kv.second);
    c->free_blocks.clear();
    c->pooled_bytes = 0;
    return ITB_OK;
}

// ---- plan device tables ---------------------------------------------------------------------------
static void release_tables(DeviceTables*& dev, void*& dev_ctx) {
    if (!dev) return;
    auto* c = (itb_ctx*)dev_ctx;
    if (c && dev->base) pool_free(c, dev->base);
    delete dev;
    dev = nullptr;
    dev_ctx = nullptr;
}
void itb_contract_plan_release_device(itb_contract_plan* P) { release_tables(P->dev, P->dev_ctx); }
void itb_permute_plan_release_device(itb_permute_plan* P) { release_tables(P->dev, P->dev_ctx); }

static int upload(itb_ctx* c, Packer& pk, size_t extra_dev_bytes, DeviceTables* dev) {
    const size_t total = pk.total + extra_dev_bytes;
    int rc = pool_alloc(c, total ? total : 256, &dev->base);
    if (rc != ITB_OK) return rc;
    dev->bytes = total;
    if (pk.total) {
        const int slot = c->staging_next;
        c->staging_next = (slot + 1) % itb_ctx::kStagingSlots;
        // this buffer's previous upload (kStagingSlots uploads ago) may still be in flight: wait for THAT copy only
        if (c->staging_busy[slot]) { CUDA_TRY(cudaEventSynchronize(c->ev_staging[slot])); c->staging_busy[slot] = false; }
        rc = ensure_staging(c, slot, pk.total);
        if (rc != ITB_OK) return rc;
        char* stg = (char*)c->staging[slot];
        for (size_t i = 0; i < pk.parts.size(); ++i)
            if (pk.parts[i].second) std::memcpy(stg + pk.offs[i], pk.parts[i].first, pk.parts[i].second);
        CUDA_TRY(cudaMemcpyAsync(dev->base, stg, pk.total, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaEventRecord(c->ev_staging[slot], c->stream));
        c->staging_busy[slot] = true;
    }
    return ITB_OK;
}

static int ensure_contract_tables(itb_ctx* c, itb_contract_plan* P) {
    if (P->dev && P->dev_ctx == c) return ITB_OK;
    if (P->dev) release_tables(P->dev, P->dev_ctx);
    if (!P->tables_built) { int rc = build_contract_tables(*P); if (rc != ITB_OK) return rc; }
    auto* dev = new DeviceTables();
    Packer pk;
    const size_t o_pairs = pk.add(P->pairs.data(), P->pairs.size() * sizeof(ItbPair));
    const size_t o_cblk = pk.add(P->cblks.data(), P->cblks.size() * sizeof(ItbCBlk));
    const size_t o_tiles = pk.add(P->qitems.data(), P->qitems.size() * sizeof(ItbQItem));
    const size_t o_tiles32 = pk.add(P->tiles.data(), P->tiles.size() * sizeof(ItbTile));
    const size_t o_splits = pk.add(P->splits.data(), P->splits.size() * sizeof(ItbSplitOut));
    const size_t o_cta = pk.add(P->cta_begin.data(), P->cta_begin.size() * sizeof(int32_t));
    const size_t o_rg = pk.add(P->rgroups.data(), P->rgroups.size() * sizeof(ItbRowGroup));
    const size_t o_rgi = pk.add(P->rg_in.data(), P->rg_in.size() * sizeof(ItbRgIn));
    const size_t o_rgo = pk.add(P->rg_out.data(), P->rg_out.size() * sizeof(int64_t));
    const size_t o_rgw = pk.add(P->rg_w.data(), P->rg_w.size() * sizeof(ItbRgW));
    const size_t o_rgit = pk.add(P->rg_items.data(), P->rg_items.size() * sizeof(ItbRgItem));
    const size_t o_sk = pk.add(P->skinny.data(), P->skinny.size() * sizeof(ItbSkinny));
    const size_t o_sq4 = pk.add(P->skinny_q4.data(), P->skinny_q4.size() * sizeof(ItbSkinny));
    const size_t o_sq8 = pk.add(P->skinny_q8.data(), P->skinny_q8.size() * sizeof(ItbSkinny));
    const size_t o_dot = pk.add(P->dots.data(), P->dots.size() * sizeof(ItbDot));
    const size_t o_dout = pk.add(P->dot_outs.data(), P->dot_outs.size() * sizeof(ItbDotOut));
    const size_t partial_bytes = ((size_t)P->ndot_slots * 4 * sizeof(double) + 255) & ~(size_t)255;
    const size_t counter_bytes = ((16 + P->splits.size()) * sizeof(int) + 255) & ~(size_t)255; // queue head, finished CTAs, per-cut-tile arrivals
    const size_t extra = partial_bytes + counter_bytes;
    int rc = upload(c, pk, extra, dev);
    if (rc != ITB_OK) { delete dev; return rc; }
    char* b = (char*)dev->base;
    dev->pairs = (const ItbPair*)(b + o_pairs);
    dev->cblks = (const ItbCBlk*)(b + o_cblk);
    dev->qitems = (const ItbQItem*)(b + o_tiles);
    dev->tiles = (const ItbTile*)(b + o_tiles32);
    dev->cta_begin = (const int32_t*)(b + o_cta);
    dev->splits = (const ItbSplitOut*)(b + o_splits);
    dev->rgroups = (const ItbRowGroup*)(b + o_rg);
    dev->rg_in = (const ItbRgIn*)(b + o_rgi);
    dev->rg_out = (const int64_t*)(b + o_rgo);
    dev->rg_w = (const ItbRgW*)(b + o_rgw);
    dev->rg_items = (const ItbRgItem*)(b + o_rgit);
    dev->rg_max_nout = 0;
    for (auto& g : P->rgroups) dev->rg_max_nout = std::max(dev->rg_max_nout, (int)g.nout);
    dev->skinny = (const ItbSkinny*)(b + o_sk);
    dev->skinny_q4 = (const ItbSkinny*)(b + o_sq4);
    dev->skinny_q8 = (const ItbSkinny*)(b + o_sq8);
    dev->dots = (const ItbDot*)(b + o_dot);
    dev->dot_outs = (const ItbDotOut*)(b + o_dout);
    dev->dot_partial = (double*)(b + pk.total);
    dev->counters = (int*)(b + pk.total + partial_bytes);
    CUDA_TRY(cudaMemsetAsync(dev->counters, 0, counter_bytes, c->stream)); // all rearmed by the kernel itself after every launch
    P->dev = dev;
    P->dev_ctx = c;
    return ITB_OK;
}

static int contract_run_impl(itb_ctx* c, itb_contract_plan* P, const void* dA, const void* dB, void* dC, const ItbMirrors* mir);
int itb_contract_run(itb_ctx* c, itb_contract_plan* P, const void* dA, const void* dB, void* dC) { return contract_run_impl(c, P, dA, dB, dC, nullptr); }

// Multi-GPU: the same contraction, every element of C additionally stored into n peer copies of C (buffers of other ranks mapped
// with itb_p2p_open): the exchange of the rows a rank owns rides the epilogue of the kernel that produces them, tile by
// tile over NVLink, instead of following it as a collective. Supported when all executed C blocks are in the DMMA tile class
// on the static kernel (the *R step of LocalOp::product); otherwise ITB_ERR_UNSUPPORTED and the caller pushes the rows
// with a block-copy plan after itb_contract_run.
int itb_contract_run_mirrored(itb_ctx* c, itb_contract_plan* P, const void* dA, const void* dB, void* dC, int32_t n, void* const* peer_dC) {
    if (!c || !P || n < 0 || n > ITB_MAX_MIRRORS || (n > 0 && !peer_dC)) { set_error("contract_run_mirrored: bad arguments"); return ITB_ERR_INVALID; }
    if (n == 0) return contract_run_impl(c, P, dA, dB, dC, nullptr);
    ItbMirrors mir;
    mir.n = n; mir.pad_ = 0;
    for (int q = 0; q < ITB_MAX_MIRRORS; ++q) mir.delta[q] = 0;
    for (int q = 0; q < n; ++q) {
        const long long d = (const char*)peer_dC[q] - (const char*)dC;
        if (d % 8 != 0) { set_error("contract_run_mirrored: peer buffer not 8-byte aligned relative to C"); return ITB_ERR_INVALID; }
        mir.delta[q] = d / 8;
    }
    return contract_run_impl(c, P, dA, dB, dC, &mir);
}

static int contract_run_impl(itb_ctx* c, itb_contract_plan* P, const void* dA, const void* dB, void* dC, const ItbMirrors* mir) {
    if (!c || !P) { set_error("contract_run: null"); return ITB_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(c->device));
    if (P->C.nelems == 0 || P->triples.empty()) return ITB_OK; // no output blocks: nothing to do
    if (plan_note_run(*P) && P->dev) release_tables(P->dev, P->dev_ctx); // (stream-ordered pool: launches already queued keep their tables)
    int rc = ensure_contract_tables(c, P);
    if (rc != ITB_OK) return rc;
    DeviceTables* d = P->dev;
    const double* A = (const double*)dA;
    const double* B = (const double*)dB;
    double* C = (double*)dC;
    bool ran[5] = {false, false, false, false, false};
    if (c->profile && !c->pev[0])
        for (auto& e : c->pev) CUDA_TRY(cudaEventCreate(&e));
#define PROF_BEGIN(i) do { if (c->profile) CUDA_TRY(cudaEventRecord(c->pev[2 * (i)], c->stream)); } while (0)
#define PROF_END(i) do { if (c->profile) { CUDA_TRY(cudaEventRecord(c->pev[2 * (i) + 1], c->stream)); ran[i] = true; } } while (0)
    // The streaming and split-K-dot classes write disjoint C blocks; when the tile kernel also runs they go to
    // the side stream FIRST (fork/join with events) so that their latency-bound CTAs overlap the persistent
    // tile kernel instead of trailing it. In profile mode everything stays on one stream for per-class timing.
    const bool has_tiles = !P->tiles.empty();
    const bool has_stream = !P->skinny.empty() || !P->skinny_q4.empty() || !P->skinny_q8.empty() || !P->rg_items.empty();
    const bool has_dots = !P->dots.empty();
    if (mir) {
        // mirrored stores exist in the static tile kernel's epilogue and in the split-K reduction only
        static const bool ring_forced = [] { const char* e = getenv("ITB_TILE_KERNEL"); return e && std::string(e) == "ring"; }();
        const int ns = (int)P->cta_begin.size() - 2;
        const bool static_sched = ns > 0 && has_tiles && P->cta_begin[ns] == (int32_t)P->tiles.size();
        if (!has_tiles || has_stream || has_dots || !static_sched || ring_forced) {
            set_error("contract_run_mirrored: this plan has work outside the static DMMA tile class");
            return ITB_ERR_UNSUPPORTED;
        }
    }
    const bool fork = has_tiles && (has_stream || has_dots) && !c->profile;
    cudaStream_t side = fork ? c->aux : c->stream;
    if (fork) {
        CUDA_TRY(cudaEventRecord(c->ev_fork, c->stream));
        CUDA_TRY(cudaStreamWaitEvent(c->aux, c->ev_fork, 0));
    }
    auto launch_side = [&]() -> int {
        if (has_stream) {
            PROF_BEGIN(3);
            CUDA_TRY(launch_rowgroups(d->rg_items, (int)P->rg_items.size(), d->rgroups, d->rg_in, d->rg_out, d->rg_w, d->rg_max_nout, A, B, C, side));
            c->launches += P->rg_items.empty() ? 0 : 1;
            CUDA_TRY(launch_skinny(d->skinny, (int)P->skinny.size(), d->skinny_q4, (int)P->skinny_q4.size(), d->skinny_q8,
                                   (int)P->skinny_q8.size(), d->cblks, d->pairs, A, B, C, side));
            PROF_END(3);
            c->launches += (P->skinny.empty() ? 0 : 1) + (P->skinny_q4.empty() ? 0 : 1) + (P->skinny_q8.empty() ? 0 : 1);
        }
        if (has_dots) {
            PROF_BEGIN(4);
            CUDA_TRY(launch_dot(d->dots, (int)P->dots.size(), d->dot_outs, (int)P->dot_outs.size(), d->cblks, d->pairs, A, B, d->dot_partial, C, side));
            PROF_END(4);
            c->launches += 2;
        }
        return ITB_OK;
    };
    // from here on every exit path joins the side stream back into the main stream (join_side), so a failed launch or
    // allocation cannot leave work on aux that the caller's next operation on the main stream would race with
    auto join_side = [&]() {
        if (!fork) return;
        if (cudaEventRecord(c->ev_join, c->aux) != cudaSuccess || cudaStreamWaitEvent(c->stream, c->ev_join, 0) != cudaSuccess) {
            (void)cudaGetLastError();
            cudaStreamSynchronize(c->aux);
        }
    };
#define SIDE_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { join_side(); set_error(std::string(#expr) + ": " + cudaGetErrorString(e__)); return ITB_ERR_CUDA; } } while (0)
    if (fork) { rc = launch_side(); if (rc != ITB_OK) { join_side(); return rc; } SIDE_TRY(cudaEventRecord(c->ev_join, c->aux)); }
    if (has_tiles) {
        if (P->ws_slots > 0) { rc = ensure_ws(c, (size_t)P->ws_slots * ITB_WS_TILE); if (rc != ITB_OK) { join_side(); return rc; } }
        PROF_BEGIN(0);
        // persistent grid, one CTA per SM (fewer when the queue is shorter than that); items are pulled from the
        // in-order queue through the head counter (rearmed by the kernel itself)
        // hybrid plans carry one static range per CTA of the planner's grid width: launch exactly that many CTAs
        const int n_static = (int)P->cta_begin.size() - 2;
        const bool has_static = n_static > 0 && P->cta_begin[n_static] > 0;
        const int grid = has_static ? n_static : (int)std::min<size_t>((size_t)c->num_sms, P->tiles.size());
        const size_t prof_words = 1024 + 4 * P->tiles.size(); // per-CTA spans, then {cta, start, K-loop end, end} per item
        if (c->profile && c->cta_cycles_words < prof_words) {
            if (c->d_cta_cycles) { SIDE_TRY(cudaStreamSynchronize(c->stream)); cudaFree(c->d_cta_cycles); c->d_cta_cycles = nullptr; }
            SIDE_TRY(cudaMalloc(&c->d_cta_cycles, prof_words * sizeof(long long)));
            c->cta_cycles_words = prof_words;
        }
        // two kernels with the same producer / consumer loops: a purely static schedule (every item in some CTA's range,
        // nothing in the shared queue) runs on the kernel without the item ring (kernels_gemm_static.cu; ITB_TILE_KERNEL=ring
        // forces the other one), anything with a dynamic part on the ring kernel (kernels_gemm.cu)
        static const bool force_ring = [] { const char* e = getenv("ITB_TILE_KERNEL"); return e && std::string(e) == "ring"; }();
        if (!force_ring && has_static && P->cta_begin[n_static] == (int32_t)P->tiles.size()) {
            SIDE_TRY(launch_gemm_static(d->tiles, d->cta_begin, grid, d->splits, (int)P->splits.size(), d->cblks, d->pairs, A, B, C, c->ws,
                                        c->profile ? c->d_cta_cycles : nullptr, mir, c->stream));
            c->h_item_cycles.clear();
        } else if (P->qitems.size() != P->tiles.size()) {
            join_side();
            set_error("contract_run: the dynamic-queue kernel was selected but the plan carries no queue records");
            return ITB_ERR_INVALID;
        } else
        SIDE_TRY(launch_gemm(d->qitems, (int)P->tiles.size(), d->counters, d->cta_begin, n_static, grid, d->splits, (int)P->splits.size(), d->cblks, d->pairs,
                             A, B, C, c->ws, c->profile ? c->d_cta_cycles : nullptr, c->stream));
        if (c->profile) {
            c->h_cta_cycles.assign(grid, 0);
            CUDA_TRY(cudaMemcpyAsync(c->h_cta_cycles.data(), c->d_cta_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
            const bool ring_ran = force_ring || !(has_static && P->cta_begin[n_static] == (int32_t)P->tiles.size());
            c->h_item_cycles.assign(ring_ran ? 4 * P->tiles.size() : 0, 0);
            if (ring_ran) CUDA_TRY(cudaMemcpyAsync(c->h_item_cycles.data(), c->d_cta_cycles + 1024, 4 * P->tiles.size() * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        }
        PROF_END(0);
        c->launches += P->splits.empty() ? 1 : 2;
    }
    if (fork) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    else { rc = launch_side(); if (rc != ITB_OK) return rc; }
#undef PROF_BEGIN
#undef PROF_END
#undef SIDE_TRY
    if (c->profile) {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < 5; ++i) {
            c->last_ms[i] = 0.f;
            if (ran[i]) CUDA_TRY(cudaEventElapsedTime(&c->last_ms[i], c->pev[2 * i], c->pev[2 * i + 1]));
        }
    }
    return ITB_OK;
}

// Measured refinement of the static stream-K partition. The planner's cycle model decides how much of the (tile, K-chunk)
// list every CTA of the persistent grid gets; what it gets wrong shows up as spread between the per-CTA clock64 spans
// (max/mean 1.09-1.12 on the maxdim-2000 H_eff step, i.e. ~10 % of the tile kernel's time). A plan that is executed many
// times (Davidson repeats the same four structures) can close that loop: run, read the spans, scale the modelled cost of
// every tile by measured/modelled of the CTAs that ran it, re-partition, and keep the partition with the smallest
// longest span. The result of every run is the correct C, so this can replace the first executions of the plan.
int itb_contract_plan_refine(itb_ctx* c, itb_contract_plan* P, const void* dA, const void* dB, void* dC, int rounds, double* gain) {
    if (!c || !P) { set_error("plan_refine: null"); return ITB_ERR_INVALID; }
    if (gain) *gain = 1.0;
    const bool was_profile = c->profile;
    c->profile = true;
    struct Restore { itb_ctx* c; bool v; ~Restore() { c->profile = v; } } restore{c, was_profile};
    std::vector<double> best_scale;
    double best_max = 0, first_max = 0;
    bool have_best = false;
    for (int r = 0; r <= rounds; ++r) {
        int rc = itb_contract_run(c, P, dA, dB, dC); // profile mode: synchronises, h_cta_cycles valid afterwards
        if (rc != ITB_OK) return rc;
        const int n_static = (int)P->cta_begin.size() - 2;
        const bool purely_static = n_static > 0 && !P->tiles.empty() && P->cta_begin[n_static] == (int32_t)P->tiles.size();
        if (!purely_static || (int)c->h_cta_cycles.size() < n_static || P->item_tile.size() != P->tiles.size()) return ITB_OK; // nothing to refine
        int32_t ntile = 0;
        for (int32_t t : P->item_tile) ntile = std::max(ntile, t + 1);
        if (P->tile_scale.size() != (size_t)ntile) P->tile_scale.assign(ntile, 1.0);
        double mx = 0;
        for (int b = 0; b < n_static; ++b) mx = std::max(mx, (double)c->h_cta_cycles[b]);
        if (r == 0) first_max = mx;
        if (!have_best || mx < best_max) { best_max = mx; best_scale = P->tile_scale; have_best = true; }
        if (r == rounds) break;
        // per-CTA measured / modelled, normalised to mean 1 so that the scales do not drift
        std::vector<double> ratio(n_static, 1.0);
        double sum_meas = 0, sum_model = 0;
        for (int b = 0; b < n_static; ++b) {
            double model = 0;
            for (int32_t i = P->cta_begin[b]; i < P->cta_begin[b + 1]; ++i) model += P->item_cost[i];
            ratio[b] = model > 0 ? (double)c->h_cta_cycles[b] / model : 1.0;
            if (model > 0) { sum_meas += (double)c->h_cta_cycles[b]; sum_model += model; }
        }
        const double norm = sum_model > 0 ? sum_meas / sum_model : 1.0;
        // tile factor = cost-weighted mean of the ratios of the CTAs that ran its pieces (damped)
        std::vector<double> num(ntile, 0.0), den(ntile, 0.0);
        for (int b = 0; b < n_static; ++b)
            for (int32_t i = P->cta_begin[b]; i < P->cta_begin[b + 1]; ++i) {
                num[P->item_tile[i]] += P->item_cost[i] * ratio[b] / norm;
                den[P->item_tile[i]] += P->item_cost[i];
            }
        for (int32_t t = 0; t < ntile; ++t)
            if (den[t] > 0) {
                const double f = std::min(1.5, std::max(0.67, num[t] / den[t]));
                P->tile_scale[t] *= std::pow(f, 0.8);
            }
        P->tables_built = false;
        if (P->dev) { CUDA_TRY(cudaStreamSynchronize(c->stream)); release_tables(P->dev, P->dev_ctx); P->dev = nullptr; }
    }
    if (have_best && best_scale != P->tile_scale) {
        P->tile_scale = best_scale;
        P->tables_built = false;
        if (P->dev) { CUDA_TRY(cudaStreamSynchronize(c->stream)); release_tables(P->dev, P->dev_ctx); P->dev = nullptr; }
    }
    if (gain && best_max > 0) *gain = first_max / best_max;
    return ITB_OK;
}

int itb_contract_host(itb_ctx* c, itb_contract_plan* P, const void* hA, const void* hB, void* hC) {
    if (!c || !P) { set_error("contract_host: null"); return ITB_ERR_INVALID; }
    const size_t ba = (size_t)P->A.nelems * (P->A.dtype == ITB_C64 ? 16 : 8);
    const size_t bb = (size_t)P->B.nelems * (P->B.dtype == ITB_C64 ? 16 : 8);
    const size_t bc = (size_t)P->C.nelems * (P->C.dtype == ITB_C64 ? 16 : 8);
    void *dA = nullptr, *dB = nullptr, *dC = nullptr;
    int rc = itb_malloc(c, ba, &dA);
    if (rc == ITB_OK) rc = itb_malloc(c, bb, &dB);
    if (rc == ITB_OK) rc = itb_malloc(c, bc, &dC);
    if (rc == ITB_OK) rc = itb_memcpy_h2d(c, dA, hA, ba);
    if (rc == ITB_OK) rc = itb_memcpy_h2d(c, dB, hB, bb);
    if (rc == ITB_OK) rc = itb_contract_run(c, P, dA, dB, dC);
    if (rc == ITB_OK) rc = itb_memcpy_d2h(c, hC, dC, bc);
    if (dA) pool_free(c, dA);
    if (dB) pool_free(c, dB);
    if (dC) pool_free(c, dC);
    return rc;
}

static int ensure_permute_tables(itb_ctx* c, itb_permute_plan* P) {
    if (P->dev && P->dev_ctx == c) return ITB_OK;
    if (P->dev) release_tables(P->dev, P->dev_ctx);
    auto* dev = new DeviceTables();
    Packer pk;
    const size_t o_c = pk.add(P->blks_copy.data(), P->blks_copy.size() * sizeof(ItbPermBlk));
    const size_t o_ch = pk.add(P->chunk_items.data(), P->chunk_items.size() * sizeof(ItbPermChunk));
    const size_t o_t = pk.add(P->tile_items.data(), P->tile_items.size() * sizeof(ItbPermTile));
    int rc = upload(c, pk, 0, dev);
    if (rc != ITB_OK) { delete dev; return rc; }
    dev->pcopy = (const ItbPermBlk*)((char*)dev->base + o_c);
    dev->pchunks = (const ItbPermChunk*)((char*)dev->base + o_ch);
    dev->ptiles = (const ItbPermTile*)((char*)dev->base + o_t);
    P->dev = dev;
    P->dev_ctx = c;
    return ITB_OK;
}

int itb_permute_run(itb_ctx* c, itb_permute_plan* P, const void* dSrc, void* dDst, double ar, double ai, int accumulate) {
    if (!c || !P) { set_error("permute_run: null"); return ITB_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(c->device));
    if (P->S.dtype == ITB_F64 && P->D.dtype == ITB_F64 && ai != 0.0) { set_error("permute_run: complex alpha on a real destination"); return ITB_ERR_INVALID; }
    if (!accumulate && P->need_zero && P->D.nelems > 0) {
        // only the destination blocks without a source need zeros (permuteQDense fill-in)
        const size_t es = P->D.dtype == ITB_C64 ? 16 : 8;
        if (P->zero_ranges.size() <= 2 * 64) {
            for (size_t i = 0; i + 1 < P->zero_ranges.size(); i += 2)
                CUDA_TRY(cudaMemsetAsync((char*)dDst + (size_t)P->zero_ranges[i] * es, 0, (size_t)P->zero_ranges[i + 1] * es, c->stream));
        } else {
            CUDA_TRY(cudaMemsetAsync(dDst, 0, (size_t)P->D.nelems * es, c->stream));
        }
    }
    if (P->S.nblocks == 0) return ITB_OK;
    int rc = ensure_permute_tables(c, P);
    if (rc != ITB_OK) return rc;
    int launches = 0;
    CUDA_TRY(launch_permute(P->S.dtype == ITB_C64, P->D.dtype == ITB_C64, P->dev->pcopy, P->dev->pchunks, P->items_copy,
                            P->dev->ptiles, P->items_tiled, dSrc, dDst, ar, ai, accumulate, c->stream, &launches));
    c->launches += launches;
    return ITB_OK;
}

int itb_permute_host(itb_ctx* c, itb_permute_plan* P, const void* hSrc, void* hDst, double ar, double ai, int accumulate) {
    if (!c || !P) { set_error("permute_host: null"); return ITB_ERR_INVALID; }
    const size_t bs = (size_t)P->S.nelems * (P->S.dtype == ITB_C64 ? 16 : 8);
    const size_t bd = (size_t)P->D.nelems * (P->D.dtype == ITB_C64 ? 16 : 8);
    void *dS = nullptr, *dD = nullptr;
    int rc = itb_malloc(c, bs, &dS);
    if (rc == ITB_OK) rc = itb_malloc(c, bd, &dD);
    if (rc == ITB_OK) rc = itb_memcpy_h2d(c, dS, hSrc, bs);
    if (rc == ITB_OK && accumulate) rc = itb_memcpy_h2d(c, dD, hDst, bd);
    if (rc == ITB_OK) rc = itb_permute_run(c, P, dS, dD, ar, ai, accumulate);
    if (rc == ITB_OK) rc = itb_memcpy_d2h(c, hDst, dD, bd);
    if (dS) pool_free(c, dS);
    if (dD) pool_free(c, dD);
    return rc;
}

// ---- BLAS-1 ---------------------------------------------------------------------------------------------
int itb_scal(itb_ctx* c, int32_t dtype, int64_t n, void* x, double ar, double ai) {
    if (dtype == ITB_F64 && ai != 0.0) { set_error("itb_scal: complex factor on real data"); return ITB_ERR_INVALID; }
    CUDA_TRY(launch_scal(dtype == ITB_C64, n, x, ar, ai, c->num_sms, c->stream));
    if (n > 0) ++c->launches;
    return ITB_OK;
}
int itb_axpy(itb_ctx* c, int32_t dtype, int64_t n, double ar, double ai, const void* x, void* y) {
    if (dtype == ITB_F64 && ai != 0.0) { set_error("itb_axpy: complex factor on real data"); return ITB_ERR_INVALID; }
    CUDA_TRY(launch_axpy(dtype == ITB_C64, n, ar, ai, x, y, c->num_sms, c->stream));
    if (n > 0) ++c->launches;
    return ITB_OK;
}
int itb_fill(itb_ctx* c, int32_t dtype, int64_t n, void* x, double re, double im) {
    CUDA_TRY(launch_fill(dtype == ITB_C64, n, x, re, im, c->num_sms, c->stream));
    if (n > 0) ++c->launches;
    return ITB_OK;
}
int itb_conj(itb_ctx* c, int64_t n, void* x) {
    CUDA_TRY(launch_conj(n, x, c->num_sms, c->stream));
    if (n > 0) ++c->launches;
    return ITB_OK;
}
int itb_real_to_cplx(itb_ctx* c, int64_t n, const void* x, void* y) {
    CUDA_TRY(launch_r2c(n, x, y, c->num_sms, c->stream));
    if (n > 0) ++c->launches;
    return ITB_OK;
}
int itb_take_part(itb_ctx* c, int64_t n, const void* x, void* y, int imag) {
    CUDA_TRY(launch_part(n, x, y, imag, c->num_sms, c->stream));
    if (n > 0) ++c->launches;
    return ITB_OK;
}

int itb_nrm2(itb_ctx* c, int32_t dtype, int64_t n, const void* x, double* out) {
    if (!out) { set_error("itb_nrm2: null out"); return ITB_ERR_INVALID; }
    *out = 0.0;
    if (n <= 0) return ITB_OK;
    const int64_t nr = dtype == ITB_C64 ? 2 * n : n;
    int rc = ensure_scratch(c, (size_t)c->num_sms * 8 * 2 + 8);
    if (rc != ITB_OK) return rc;
    int g = 0;
    CUDA_TRY(launch_ssq(nr, x, 1.0, c->scratch, &g, c->num_sms, c->stream));
    c->launches += 2;
    CUDA_TRY(cudaMemcpyAsync(c->h_result, c->scratch + 2 * g, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    double ssq = c->h_result[0];
    const double mx = c->h_result[1];
    // dnrm2 is overflow/underflow safe (scaled sum of squares); redo scaled only when the plain sum left the
    // comfortable range
    if (mx > 0.0 && (!(ssq < 1e300) || ssq < 1e-280)) {
        CUDA_TRY(launch_ssq(nr, x, 1.0 / mx, c->scratch, &g, c->num_sms, c->stream));
        c->launches += 2;
        CUDA_TRY(cudaMemcpyAsync(c->h_result, c->scratch + 2 * g, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        *out = mx * std::sqrt(c->h_result[0]);
        return ITB_OK;
    }
    *out = std::sqrt(ssq);
    return ITB_OK;
}

int itb_dot(itb_ctx* c, int32_t dtype, int64_t n, const void* x, const void* y, int conj_x, double out[2]) {
    out[0] = out[1] = 0.0;
    if (n <= 0) return ITB_OK;
    int rc = ensure_scratch(c, (size_t)c->num_sms * 8 * 2 + 8);
    if (rc != ITB_OK) return rc;
    int g = 0;
    CUDA_TRY(launch_dot1(dtype == ITB_C64, n, x, y, conj_x, c->scratch, &g, c->num_sms, c->stream));
    c->launches += 2;
    CUDA_TRY(cudaMemcpyAsync(c->h_result, c->scratch + 2 * g, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    out[0] = c->h_result[0];
    out[1] = c->h_result[1];
    return ITB_OK;
}

int itb_get_elt(itb_ctx* c, int32_t dtype, const void* x, int64_t offset, double out[2]) {
    const size_t es = dtype == ITB_C64 ? 16 : 8;
    out[1] = 0.0;
    CUDA_TRY(cudaMemcpyAsync(c->h_result, (const char*)x + offset * es, es, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    out[0] = c->h_result[0];
    if (dtype == ITB_C64) out[1] = c->h_result[1];
    return ITB_OK;
}

int itb_peak_fp64(itb_ctx* c, int which, int iters, double* tflops) {
    int rc = ensure_scratch(c, 64);
    if (rc != ITB_OK) return rc;
    CUDA_TRY(launch_peak(which, 16, c->scratch, c->num_sms, c->stream)); // warm-up
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    CUDA_TRY(launch_peak(which, iters, c->scratch, c->num_sms, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->ev1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    const double warps = (double)c->num_sms * 8 * 8;
    const double flops = which == 1 ? warps * 32.0 * 16.0 * 2.0 * iters : warps * 8.0 * 512.0 * iters;
    *tflops = flops / (ms * 1e-3) / 1e12;
    c->launches += 2;
    return ITB_OK;
}

static int ensure_solver(itb_ctx* c) {
    if (c->solver) return ITB_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    return itb_solver_create((void*)c->stream, &c->solver);
}
int itb_syevd_host(itb_ctx* c, int32_t dtype, int32_t n, void* hA, double* hW, int32_t* info) {
    int rc = ensure_solver(c);
    if (rc != ITB_OK) return rc;
    c->launches += 1;
    return itb_solver_syevd(c->solver, dtype, n, hA, hW, info);
}
int itb_gesvd_host(itb_ctx* c, int32_t dtype, int32_t m, int32_t n, void* hA, double* hS, void* hU, void* hVT, int32_t* info) {
    int rc = ensure_solver(c);
    if (rc != ITB_OK) return rc;
    c->launches += 1;
    return itb_solver_gesvd(c->solver, dtype, m, n, hA, hS, hU, hVT, info);
}

int itb_svd_batch_run(itb_ctx* c, int32_t dtype, int64_t nblocks, const int64_t* a_off, const int32_t* m, const int32_t* n,
                      const void* dA, itb_svd_batch** out) {
    if (!c || !out || nblocks < 0) { set_error("svd_batch_run: bad arguments"); return ITB_ERR_INVALID; }
    int rc = ensure_solver(c);
    if (rc != ITB_OK) return rc;
    c->launches += nblocks;
    return itb_solver_svd_batch_run(c->solver, dtype, nblocks, a_off, m, n, dA, out);
}

int itb_eigh_batch_run(itb_ctx* c, int32_t dtype, int64_t nblocks, const int64_t* a_off, const int32_t* n, const void* dA, int negate,
                       itb_eigh_batch** out) {
    if (!c || !out || nblocks < 0) { set_error("eigh_batch_run: bad arguments"); return ITB_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(c->device)); // (the plugin drives this call from a helper thread)
    int rc = ensure_solver(c);
    if (rc != ITB_OK) return rc;
    c->launches += 2 * nblocks;
    return itb_solver_eigh_batch_run(c->solver, dtype, nblocks, a_off, n, dA, negate, out);
}

int itb_ctx_set_profile(itb_ctx* c, int profile) { c->profile = profile != 0; return ITB_OK; }
int64_t itb_contract_last_cta_cycles(itb_ctx* c, int64_t* out, int64_t cap) {
    for (int64_t i = 0; out && i < (int64_t)c->h_cta_cycles.size() && i < cap; ++i) out[i] = c->h_cta_cycles[i];
    return (int64_t)c->h_cta_cycles.size();
}
int64_t itb_contract_last_item_cycles(itb_ctx* c, int64_t* out, int64_t cap) {
    for (int64_t i = 0; out && i < (int64_t)c->h_item_cycles.size() && i < cap; ++i) out[i] = c->h_item_cycles[i];
    return (int64_t)c->h_item_cycles.size() / 4;
}
int itb_contract_last_ms(itb_ctx* c, float ms[5]) {
    for (int i = 0; i < 5; ++i) ms[i] = c->last_ms[i];
    return ITB_OK;
}

int itb_timer_start(itb_ctx* c) { CUDA_TRY(cudaEventRecord(c->ev0, c->stream)); return ITB_OK; }
int itb_timer_stop_ms(itb_ctx* c, float* ms) {
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->ev1));
    CUDA_TRY(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return ITB_OK;
}

} // extern "C"
