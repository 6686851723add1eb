// comm.cu — one-process-per-GPU communicator behind the C ABI (include/itb200.h, itb_comm_*): the all-gather that
// re-replicates row-sharded tensors (H*phi after LocalOp::product) over NVLink / NVSwitch.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2) and only when a communicator is created: a single-GPU process,
// the Python mirror (which brings torch's own NCCL) and the CPU-only build check never load it. The transport is plain
// ncclAllGather on the context's stream; what is B200-specific is how little is sent: the caller packs exactly the rows
// a rank owns, so every element of the result crosses the switch once.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <string>

#include "../../include/itb200.h"

namespace itb { void set_error(const std::string& msg); }

struct itb_comm {
    void* lib = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static void* nccl_lib() {
    static void* h = nullptr;
    if (!h) {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
    }
    return h;
}
template <typename F> static bool sym(void* lib, const char* name, F& f) {
    f = reinterpret_cast<F>(dlsym(lib, name));
    return f != nullptr;
}

extern "C" {

int itb_comm_unique_id(uint8_t out[ITB_COMM_ID_BYTES]) {
    static_assert(ITB_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "unique id size");
    void* lib = nccl_lib();
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    if (!lib || !sym(lib, "ncclGetUniqueId", GetUniqueId)) { itb::set_error("itb_comm: libnccl.so.2 not found"); return ITB_ERR_UNSUPPORTED; }
    ncclUniqueId id;
    if (GetUniqueId(&id) != ncclSuccess) { itb::set_error("itb_comm: ncclGetUniqueId failed"); return ITB_ERR_CUDA; }
    std::memcpy(out, id.internal, NCCL_UNIQUE_ID_BYTES);
    return ITB_OK;
}

int itb_comm_create(itb_ctx* ctx, int32_t world, int32_t rank, const uint8_t id_bytes[ITB_COMM_ID_BYTES], itb_comm** out) {
    if (!ctx || !out || world < 1 || rank < 0 || rank >= world) { itb::set_error("itb_comm_create: bad arguments"); return ITB_ERR_INVALID; }
    void* lib = nccl_lib();
    ncclResult_t (*InitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    auto* c = new itb_comm();
    c->lib = lib; c->world = world; c->rank = rank;
    if (!lib || !sym(lib, "ncclCommInitRank", InitRank) || !sym(lib, "ncclAllGather", c->AllGather) ||
        !sym(lib, "ncclCommDestroy", c->CommDestroy) || !sym(lib, "ncclGetErrorString", c->GetErrorString)) {
        delete c;
        itb::set_error("itb_comm: libnccl.so.2 not found or incomplete");
        return ITB_ERR_UNSUPPORTED;
    }
    ncclUniqueId id;
    std::memcpy(id.internal, id_bytes, NCCL_UNIQUE_ID_BYTES);
    const ncclResult_t r = InitRank(&c->comm, world, id, rank); // (the caller has made the context's device current)
    if (r != ncclSuccess) { itb::set_error(std::string("ncclCommInitRank: ") + c->GetErrorString(r)); delete c; return ITB_ERR_CUDA; }
    *out = c;
    return ITB_OK;
}

int32_t itb_comm_world(const itb_comm* c) { return c ? c->world : 1; }
int32_t itb_comm_rank(const itb_comm* c) { return c ? c->rank : 0; }

// every rank contributes count doubles at dSend; dRecv receives world*count doubles in rank order (in place when
// dSend == dRecv + rank*count). Ordered on the context's stream like every other launch.
int itb_comm_allgather(itb_comm* c, itb_ctx* ctx, const void* dSend, void* dRecv, int64_t count) {
    if (!c || !ctx || count < 0) { itb::set_error("itb_comm_allgather: bad arguments"); return ITB_ERR_INVALID; }
    if (count == 0) return ITB_OK;
    const ncclResult_t r = c->AllGather(dSend, dRecv, (size_t)count, ncclDouble, c->comm, (cudaStream_t)itb_ctx_stream(ctx));
    if (r != ncclSuccess) { itb::set_error(std::string("ncclAllGather: ") + c->GetErrorString(r)); return ITB_ERR_CUDA; }
    return ITB_OK;
}

// ---- peer memory (one process per GPU on one NVSwitch domain) ---------------------------------------------------
// A buffer that the other ranks' kernels store into directly over NVLink: the owner allocates it (plain cudaMalloc, the
// pointer is the base of its allocation, which is what CUDA IPC exports) and publishes the 64-byte handle through
// whatever channel the host has; every peer maps it with itb_p2p_open, which also turns on peer access from the
// calling context's device. No NCCL involved: the rows of H*phi a rank owns are written into every peer's copy by a
// launch of the block-copy kernel whose destination offsets point into the mapped buffers.
int itb_p2p_alloc(itb_ctx* ctx, int64_t bytes, void** dptr, uint8_t handle[ITB_P2P_HANDLE_BYTES]) {
    if (!ctx || !dptr || !handle || bytes <= 0) { itb::set_error("itb_p2p_alloc: bad arguments"); return ITB_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) <= ITB_P2P_HANDLE_BYTES, "handle size");
    if (cudaSetDevice(itb_ctx_device(ctx)) != cudaSuccess) { itb::set_error("itb_p2p_alloc: cudaSetDevice failed"); return ITB_ERR_CUDA; }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) { itb::set_error(std::string("itb_p2p_alloc: ") + cudaGetErrorString(e)); return e == cudaErrorMemoryAllocation ? ITB_ERR_NOMEM : ITB_ERR_CUDA; }
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); itb::set_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); return ITB_ERR_UNSUPPORTED; }
    std::memset(handle, 0, ITB_P2P_HANDLE_BYTES);
    std::memcpy(handle, &h, sizeof(h));
    *dptr = p;
    return ITB_OK;
}
int itb_p2p_open(itb_ctx* ctx, const uint8_t handle[ITB_P2P_HANDLE_BYTES], void** peer_ptr) {
    if (!ctx || !handle || !peer_ptr) { itb::set_error("itb_p2p_open: bad arguments"); return ITB_ERR_INVALID; }
    if (cudaSetDevice(itb_ctx_device(ctx)) != cudaSuccess) { itb::set_error("itb_p2p_open: cudaSetDevice failed"); return ITB_ERR_CUDA; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { (void)cudaGetLastError(); itb::set_error(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); return ITB_ERR_UNSUPPORTED; }
    *peer_ptr = p;
    return ITB_OK;
}
int itb_p2p_close(itb_ctx* ctx, void* peer_ptr) {
    if (!peer_ptr) return ITB_OK;
    if (ctx) cudaSetDevice(itb_ctx_device(ctx));
    if (cudaIpcCloseMemHandle(peer_ptr) != cudaSuccess) { (void)cudaGetLastError(); itb::set_error("cudaIpcCloseMemHandle failed"); return ITB_ERR_CUDA; }
    return ITB_OK;
}
int itb_p2p_free(itb_ctx* ctx, void* dptr) {
    if (!dptr) return ITB_OK;
    if (ctx) cudaSetDevice(itb_ctx_device(ctx));
    if (cudaFree(dptr) != cudaSuccess) { (void)cudaGetLastError(); itb::set_error("itb_p2p_free: cudaFree failed"); return ITB_ERR_CUDA; }
    return ITB_OK;
}

// Arrival barrier over peer memory. Every rank owns a small flag block (itb_p2p_alloc'ed, zeroed by itb_p2p_barrier_init)
// that its peers have mapped: word r is written by rank r only. One launch per barrier: the kernel takes the next epoch from
// the block's own counter (so a CUDA-graph replay needs no changing argument), publishes it into word `rank` of every peer's
// block with a system-scope release (the stores of the kernels launched before it on this stream — the pushed rows — are
// ordered before it), then waits until every peer's word in the local block has reached the epoch. The wait is bounded
// (~2 s of clock): on expiry the error word is set instead of hanging the device.
struct P2PFlags { unsigned long long arrive[ITB_P2P_MAX_WORLD]; unsigned long long epoch; unsigned long long error; };
struct P2PPeers { P2PFlags* p[ITB_P2P_MAX_WORLD]; };

__global__ void p2p_barrier_kernel(P2PFlags* local, P2PPeers peers, int world, int rank) {
    const int t = threadIdx.x;
    unsigned long long e = 0;
    if (t == 0) { e = local->epoch + 1; local->epoch = e; }
    e = __shfl_sync(0xffffffffu, e, 0);
    __threadfence_system();
    if (t < world && t != rank) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&peers.p[t]->arrive[rank]), "l"(e) : "memory");
        const long long t0 = clock64();
        unsigned long long seen = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(&local->arrive[t]) : "memory");
            if (seen >= e) break;
            if (clock64() - t0 > 4000000000ll) { local->error = e; break; }
            __nanosleep(100);
        }
    }
    __threadfence_system();
}

int itb_p2p_barrier_init(itb_ctx* ctx, void* local_flags) {
    if (!ctx || !local_flags) { itb::set_error("itb_p2p_barrier_init: bad arguments"); return ITB_ERR_INVALID; }
    if (cudaMemsetAsync(local_flags, 0, sizeof(P2PFlags), (cudaStream_t)itb_ctx_stream(ctx)) != cudaSuccess ||
        cudaStreamSynchronize((cudaStream_t)itb_ctx_stream(ctx)) != cudaSuccess) { itb::set_error("itb_p2p_barrier_init: memset failed"); return ITB_ERR_CUDA; }
    return ITB_OK;
}
int itb_p2p_barrier(itb_ctx* ctx, void* local_flags, void* const* peer_flags, int32_t world, int32_t rank) {
    if (!ctx || !local_flags || !peer_flags || world < 1 || world > ITB_P2P_MAX_WORLD || rank < 0 || rank >= world) { itb::set_error("itb_p2p_barrier: bad arguments"); return ITB_ERR_INVALID; }
    P2PPeers pp;
    for (int r = 0; r < ITB_P2P_MAX_WORLD; ++r) pp.p[r] = (r < world && r != rank) ? (P2PFlags*)peer_flags[r] : nullptr;
    p2p_barrier_kernel<<<1, 32, 0, (cudaStream_t)itb_ctx_stream(ctx)>>>((P2PFlags*)local_flags, pp, world, rank);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { itb::set_error(std::string("itb_p2p_barrier: ") + cudaGetErrorString(e)); return ITB_ERR_CUDA; }
    return ITB_OK;
}
// epochs completed and the epoch at which a wait expired (0: never); synchronises the stream
int itb_p2p_barrier_status(itb_ctx* ctx, const void* local_flags, int64_t* epoch, int64_t* error) {
    if (!ctx || !local_flags) { itb::set_error("itb_p2p_barrier_status: bad arguments"); return ITB_ERR_INVALID; }
    P2PFlags h;
    if (cudaMemcpyAsync(&h, local_flags, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)itb_ctx_stream(ctx)) != cudaSuccess ||
        cudaStreamSynchronize((cudaStream_t)itb_ctx_stream(ctx)) != cudaSuccess) { itb::set_error("itb_p2p_barrier_status: copy failed"); return ITB_ERR_CUDA; }
    if (epoch) *epoch = (int64_t)h.epoch;
    if (error) *error = (int64_t)h.error;
    return ITB_OK;
}

int itb_comm_destroy(itb_comm* c) {
    if (!c) return ITB_OK;
    if (c->comm) c->CommDestroy(c->comm);
    delete c;
    return ITB_OK;
}

} // extern "C"
