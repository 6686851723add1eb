// gather_bench.cu — microbenchmark of the ways a producer can gather a 256-row x 16-column FP64 chunk (32 KB)
// from strided global rows into shared memory on sm_100a: 8-byte cp.async (LDGSTS.64, what arbitrary element
// alignment allows), 16-byte cp.async, LDG+STS through registers, and 1-D bulk copies (cp.async.bulk, TMA engine).
// Reports cycles per chunk per SM, hot (L2-resident source) and cold (source >> L2). Build: see tools/gpu_micro.sh
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

constexpr int ROWS = 256, BK = 16, STAGES = 4, PITCH = BK + 4;

__device__ __forceinline__ void cp8(double* s, const double* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp16(double* s, const double* g, bool cg) {
    if (cg) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// mode 0: cp.async 8B, k-fast rows (lane -> k = lane%16, row = lane/16 + 2i)
// mode 1: cp.async 16B .ca ; mode 2: cp.async 16B .cg (lane -> k2 = lane%8, row = lane/8 + 4i)
// mode 3: LDG.64 + STS.64 through registers, 16 elements per thread per batch
// mode 5: cp.async 8B, m-fast (lane -> row, 256 B contiguous per warp instruction, column-major source)
__global__ void __launch_bounds__(1024) gather_kernel(const double* __restrict__ src, int64_t ld, int64_t nrows_total, int nchunks,
                                                      int mode, long long* cycles, double* sink) {
    extern __shared__ __align__(16) double sm[];
    const int nt = blockDim.x, t = threadIdx.x;
    // each CTA walks its own band of rows; consecutive chunks move along k, then to the next 256-row band
    const int64_t kchunks_per_band = ld / BK;
    __syncthreads();
    const long long t0 = clock64();
    double acc = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int64_t lin = (int64_t)blockIdx.x * nchunks + c;
        const int64_t band = (lin / kchunks_per_band) % (nrows_total / ROWS), kc = lin % kchunks_per_band;
        const double* g0 = src + band * ROWS * ld + kc * BK;
        double* st = sm + (c % STAGES) * ROWS * PITCH;
        if (mode == 0) {
            for (int i = t; i < ROWS * BK; i += nt) { const int k = i % BK, r = i / BK; cp8(st + r * PITCH + k, g0 + r * ld + k); }
        } else if (mode == 1 || mode == 2) {
            for (int i = t; i < ROWS * BK / 2; i += nt) { const int k2 = i % (BK / 2), r = i / (BK / 2); cp16(st + r * PITCH + 2 * k2, g0 + r * ld + 2 * k2, mode == 2); }
        } else if (mode == 3) {
            for (int i0 = t; i0 < ROWS * BK; i0 += nt * 16) {
                double v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) { const int i = i0 + u * nt; const int k = i % BK, r = i / BK; v[u] = (i < ROWS * BK) ? g0[r * ld + k] : 0.0; }
#pragma unroll
                for (int u = 0; u < 16; ++u) { const int i = i0 + u * nt; const int k = i % BK, r = i / BK; if (i < ROWS * BK) st[r * PITCH + k] = v[u]; }
            }
        } else if (mode == 5) {
            // column-major source: element (r,k) at g1[r + k*ldm]; here ldm = ROWS*? emulate with ld as column pitch
            const double* g1 = src + (lin % ((nrows_total * ld) / (ROWS * BK) - 1)) * ROWS * BK;
            for (int i = t; i < ROWS * BK; i += nt) { const int r = i % ROWS, k = i / ROWS; cp8(st + r + k * (ROWS + 4), g1 + r + (int64_t)k * ROWS); }
        }
        commit();
        wait_group<STAGES - 1>();
        if (mode != 3) acc += st[(t * 7) % (ROWS * PITCH)] * 1e-300; // touch (keeps the copies observable)
    }
    wait_group<0>();
    __syncthreads();
    if (t == 0) cycles[blockIdx.x] = clock64() - t0;
    if (acc == 12345.0) sink[0] = acc;
}

// mode 4: 1-D bulk copies, one row (row_bytes) per issuing thread, mbarrier complete_tx per stage
__global__ void __launch_bounds__(256) bulk_kernel(const double* __restrict__ src, int64_t ld, int64_t nrows_total, int nchunks,
                                                   int row_elems, long long* cycles) {
    extern __shared__ __align__(128) double sm[];
    __shared__ __align__(8) uint64_t bar[STAGES];
    const int t = threadIdx.x;
    const int rows = ROWS * BK / row_elems; // rows of row_elems per 32 KB chunk
    if (t == 0) for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int64_t kchunks_per_band = ld / row_elems;
    const long long t0 = clock64();
    for (int c = 0; c < nchunks + STAGES - 1; ++c) {
        if (c < nchunks) {
            const int s = c % STAGES;
            const int64_t lin = (int64_t)blockIdx.x * nchunks + c;
            const int64_t band = (lin / kchunks_per_band) % (nrows_total / rows), kc = lin % kchunks_per_band;
            const double* g0 = src + band * rows * ld + kc * row_elems;
            double* st = sm + s * ROWS * BK;
            if (t == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar[s])), "r"(ROWS * BK * 8) : "memory");
            __syncwarp();
            for (int r = t; r < rows; r += blockDim.x)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(st + r * row_elems)),
                             "l"(g0 + r * ld), "r"(row_elems * 8), "r"((unsigned)__cvta_generic_to_shared(&bar[s])) : "memory");
        }
        const int w = c - (STAGES - 1);
        if (w >= 0) { // wait for chunk w
            const int s = w % STAGES, ph = (w / STAGES) & 1;
            asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DN;\n\tbra WL;\n\tDN:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(&bar[s])), "r"(ph) : "memory");
            __syncthreads(); // stage reusable
        }
    }
    if (t == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    const int G = prop.multiProcessorCount;
    const int64_t ld = 640; // row pitch in elements (a 634-wide sector)
    long long* dcyc; cudaMalloc(&dcyc, G * sizeof(long long));
    double* sink; cudaMalloc(&sink, 8);
    std::vector<long long> h(G);
    const size_t smem = (size_t)STAGES * ROWS * (PITCH > ROWS / 16 + 4 ? PITCH : PITCH) * 8 + 4096;
    cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int cold = 0; cold < 2; ++cold) {
        const int64_t nrows = cold ? (int64_t)1 << 20 : (int64_t)4096; // cold: 1M rows x 640 x 8 B = 5.2 GB ; hot: 21 MB
        double* src; if (cudaMalloc(&src, nrows * ld * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
        cudaMemset(src, 0, nrows * ld * 8);
        const int nchunks = 400;
        const char* names[] = {"cp.async 8B  k-fast", "cp.async 16B.ca k-fast", "cp.async 16B.cg k-fast", "LDG64+STS64 k-fast", "", "cp.async 8B  m-fast"};
        for (int mode : {0, 1, 2, 3, 5}) {
            for (int nt : {128, 256, 512}) {
                if (mode == 3 && nt == 128) { /* 16 regs x 128 thr = one batch of 2048 elems */ }
                for (int rep = 0; rep < 2; ++rep) gather_kernel<<<G, nt, 200 * 1024>>>(src, ld, nrows, nchunks, mode, dcyc, sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s nt=%d: %s\n", names[mode], nt, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h.data(), dcyc, G * sizeof(long long), cudaMemcpyDeviceToHost);
                double mean = 0; for (auto v : h) mean += v; mean /= G;
                printf("%-5s %-24s threads %4d: %7.0f cycles/chunk(32KB)  %5.1f B/cycle/SM\n", cold ? "cold" : "hot", names[mode], nt, mean / nchunks, 32768.0 * nchunks / mean);
            }
        }
        for (int row_elems : {16, 32, 128, 512}) {
            for (int nt : {32, 128}) {
                for (int rep = 0; rep < 2; ++rep) bulk_kernel<<<G, nt, 200 * 1024>>>(src, ld, nrows, nchunks, row_elems, dcyc);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("bulk row=%d nt=%d: %s\n", row_elems, nt, cudaGetErrorString(e)); break; }
                cudaMemcpy(h.data(), dcyc, G * sizeof(long long), cudaMemcpyDeviceToHost);
                double mean = 0; for (auto v : h) mean += v; mean /= G;
                printf("%-5s cp.async.bulk rows of %4d B, %3d issuing threads: %7.0f cycles/chunk(32KB)  %5.1f B/cycle/SM\n", cold ? "cold" : "hot", row_elems * 8, nt, mean / nchunks, 32768.0 * nchunks / mean);
            }
        }
        cudaFree(src);
    }
    (void)smem;
    return 0;
}
