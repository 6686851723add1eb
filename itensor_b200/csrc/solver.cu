// solver.cu — per-block eigh / SVD on the device for the svdBond path.
//
// Stands behind the reference's LAPACK wrapper boundary (itensor/tensor/lapack_wrap.{h,cc}:
// dsyev_wrapper :322-347, zheev_wrapper :678-706, dgesdd/zgesdd_wrapper :365-486) — the plugin's
// lapack_gpu.cc routes blocks above a size threshold here. Library calls (cuSOLVER syevd / gesvd) are the right
// tool for a plain dense factorisation; the hand-written part of this repo is the contraction path.
// Host buffers in, host buffers out: these entry points are called from the reference's host-side per-block loops
// (hermitian.cc:231-257, svd.cc:199-222), which own the truncation logic.
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <dlfcn.h>
#include <fcntl.h>
#include <link.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/itb200.h"

namespace itb {
void set_error(const std::string& msg);
}

struct itb_solver {
    cusolverDnHandle_t h = nullptr;
    void* d_a = nullptr; size_t a_bytes = 0;
    void* d_b = nullptr; size_t b_bytes = 0;
    void* d_c = nullptr; size_t c_bytes = 0;
    void* d_w = nullptr; size_t w_bytes = 0;
    void* d_work = nullptr; size_t work_bytes = 0;
    int* d_info = nullptr;
    cudaStream_t stream = nullptr;
    gesvdjInfo_t jinfo = nullptr;
    int svd_method = 2; // 0: gesvd (QR iteration), 1: gesvdj (one-sided Jacobi), 2: gesvdp (polar decomposition) — ITB_SVD_METHOD
    cusolverDnParams_t params = nullptr;
    void* h_work = nullptr; size_t h_work_bytes = 0;
    double last_err_sigma = 0;
};

#define S_TRY(expr)                                                                         \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess) { itb::set_error(std::string(#expr) + ": " + cudaGetErrorString(e__)); return ITB_ERR_CUDA; } \
    } while (0)
#define CS_TRY(expr)                                                                        \
    do {                                                                                    \
        cusolverStatus_t s__ = (expr);                                                      \
        if (s__ != CUSOLVER_STATUS_SUCCESS) { itb::set_error(std::string(#expr) + ": cusolver status " + std::to_string((int)s__)); return ITB_ERR_CUDA; } \
    } while (0)

static int grow(void** p, size_t* have, size_t need) {
    if (*have >= need) return ITB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    S_TRY(cudaMalloc(p, need));
    *have = need;
    return ITB_OK;
}

// out[j*ldo + i] = in[i*ldi + j] (dense transpose, optionally conjugating), 32x32 tiles
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int rows, int cols) {
    __shared__ T tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = bx + threadIdx.x, j = by + r; // in is rows x cols column-major: element (i,j) at i + rows*j
        if (i < rows && j < cols) tile[r][threadIdx.x] = in[(size_t)i + (size_t)rows * j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = by + threadIdx.x, i = bx + r; // out is cols x rows column-major: element (j,i) at j + cols*i
        if (i < rows && j < cols) out[(size_t)j + (size_t)cols * i] = tile[threadIdx.x][r];
    }
}
template <typename T>
static void launch_transpose(const T* in, T* out, int rows, int cols, cudaStream_t st) {
    dim3 grid((rows + 31) / 32, (cols + 31) / 32), block(32, 8);
    transpose_kernel<T><<<grid, block, 0, st>>>(in, out, rows, cols);
}
__global__ void conj_inplace_kernel(double2* x, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i].y = -x[i].y;
}

// cuSOLVER / cuBLAS load their kernels lazily, module by module, the first time a shape needs them; on a box whose
// page cache is cold that turns every first use into random page faults against ~1.2 GB of shared objects (measured
// on a fresh B200 box: 140 s of svdBond time in the first sweep that met new block sizes, 277 s with
// CUDA_MODULE_LOADING=EAGER). One sequential read of those files in a background thread, started when the solver
// is created (DMRG spends its first sweeps on blocks too small for the device solvers), makes the later lazy loads
// hit RAM. ITB_WARM_LIBS=0 disables it.
static int collect_cuda_libs(struct dl_phdr_info* info, size_t, void* data) {
    auto* v = (std::vector<std::string>*)data;
    const char* n = info->dlpi_name;
    if (n && (strstr(n, "libcusolver") || strstr(n, "libcublas"))) v->push_back(n);
    return 0;
}
void itb_warm_library_pages() {
    static bool started = false;
    if (started) return;
    started = true;
    const char* e = getenv("ITB_WARM_LIBS");
    if (e && atoi(e) == 0) return;
    std::vector<std::string> libs;
    dl_iterate_phdr(collect_cuda_libs, &libs);
    std::thread([libs]() {
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<char> buf(8u << 20);
        size_t total = 0;
        for (auto& p : libs) {
            const int fd = open(p.c_str(), O_RDONLY);
            if (fd < 0) continue;
            posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
            ssize_t r;
            while ((r = read(fd, buf.data(), buf.size())) > 0) total += (size_t)r;
            close(fd);
        }
        if (getenv("ITB_PROFILE"))
            fprintf(stderr, "[itensor_b200] read %zu MB of cuSOLVER/cuBLAS pages in %.1f s\n", total >> 20,
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }).detach();
}

extern "C" {

int itb_solver_create(void* stream, itb_solver** out) {
    itb_warm_library_pages();
    auto* s = new itb_solver();
    s->stream = (cudaStream_t)stream;
    CS_TRY(cusolverDnCreate(&s->h));
    CS_TRY(cusolverDnSetStream(s->h, s->stream));
    S_TRY(cudaMalloc((void**)&s->d_info, sizeof(int)));
    CS_TRY(cusolverDnCreateGesvdjInfo(&s->jinfo));
    CS_TRY(cusolverDnXgesvdjSetTolerance(s->jinfo, 1e-15));
    CS_TRY(cusolverDnXgesvdjSetMaxSweeps(s->jinfo, 100));
    if (const char* e = getenv("ITB_SVD_METHOD")) s->svd_method = atoi(e);
    CS_TRY(cusolverDnCreateParams(&s->params));
    *out = s;
    return ITB_OK;
}

// symmetric / Hermitian eigendecomposition, LAPACK dsyev('V','U') / zheev semantics:
// A (n x n, column-major, host) is overwritten by the eigenvectors, w gets the eigenvalues ascending
int itb_solver_syevd(itb_solver* s, int32_t dtype, int32_t n, void* hA, double* hW, int32_t* info) {
    const size_t es = dtype == ITB_C64 ? 16 : 8;
    const size_t abytes = (size_t)n * n * es;
    int rc = grow(&s->d_a, &s->a_bytes, abytes); if (rc) return rc;
    rc = grow(&s->d_w, &s->w_bytes, (size_t)n * 8); if (rc) return rc;
    S_TRY(cudaMemcpyAsync(s->d_a, hA, abytes, cudaMemcpyHostToDevice, s->stream));
    int lwork = 0;
    if (dtype == ITB_F64) CS_TRY(cusolverDnDsyevd_bufferSize(s->h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, (double*)s->d_a, n, (double*)s->d_w, &lwork));
    else CS_TRY(cusolverDnZheevd_bufferSize(s->h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, (cuDoubleComplex*)s->d_a, n, (double*)s->d_w, &lwork));
    rc = grow(&s->d_work, &s->work_bytes, (size_t)lwork * es); if (rc) return rc;
    if (dtype == ITB_F64) CS_TRY(cusolverDnDsyevd(s->h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, (double*)s->d_a, n, (double*)s->d_w, (double*)s->d_work, lwork, s->d_info));
    else CS_TRY(cusolverDnZheevd(s->h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, (cuDoubleComplex*)s->d_a, n, (double*)s->d_w, (cuDoubleComplex*)s->d_work, lwork, s->d_info));
    int hinfo = 0;
    S_TRY(cudaMemcpyAsync(hA, s->d_a, abytes, cudaMemcpyDeviceToHost, s->stream));
    S_TRY(cudaMemcpyAsync(hW, s->d_w, (size_t)n * 8, cudaMemcpyDeviceToHost, s->stream));
    S_TRY(cudaMemcpyAsync(&hinfo, s->d_info, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    S_TRY(cudaStreamSynchronize(s->stream));
    *info = hinfo;
    return ITB_OK;
}

// thin SVD, LAPACK gesdd(jobz='S') semantics: A (m x n, column-major, host; destroyed) = U diag(S) VT with
// U m x l (ldu = m), VT l x n (ldvt = l), l = min(m,n). cuSOLVER gesvd needs m >= n, so wide blocks are
// factorised through their (conjugate) transpose on the device.
// one-sided Jacobi SVD (gesvdj): any m,n; returns V (n x l), so VT = V^H needs one transpose (+ conjugation)
static int solver_gesvdj(itb_solver* s, int32_t dtype, int32_t m, int32_t n, void* hA, double* hS, void* hU, void* hVT, int32_t* info) {
    const size_t es = dtype == ITB_C64 ? 16 : 8;
    const int l = std::min(m, n);
    int rc = grow(&s->d_a, &s->a_bytes, (size_t)m * n * es); if (rc) return rc;
    rc = grow(&s->d_b, &s->b_bytes, (size_t)m * l * es); if (rc) return rc; // U
    rc = grow(&s->d_c, &s->c_bytes, (size_t)n * l * es * 2); if (rc) return rc; // V, then VT behind it
    rc = grow(&s->d_w, &s->w_bytes, (size_t)l * 8 + 64); if (rc) return rc;
    S_TRY(cudaMemcpyAsync(s->d_a, hA, (size_t)m * n * es, cudaMemcpyHostToDevice, s->stream));
    int lwork = 0;
    if (dtype == ITB_F64) CS_TRY(cusolverDnDgesvdj_bufferSize(s->h, CUSOLVER_EIG_MODE_VECTOR, 1, m, n, (double*)s->d_a, m, (double*)s->d_w, (double*)s->d_b, m, (double*)s->d_c, n, &lwork, s->jinfo));
    else CS_TRY(cusolverDnZgesvdj_bufferSize(s->h, CUSOLVER_EIG_MODE_VECTOR, 1, m, n, (cuDoubleComplex*)s->d_a, m, (double*)s->d_w, (cuDoubleComplex*)s->d_b, m, (cuDoubleComplex*)s->d_c, n, &lwork, s->jinfo));
    rc = grow(&s->d_work, &s->work_bytes, (size_t)lwork * es); if (rc) return rc;
    if (dtype == ITB_F64) CS_TRY(cusolverDnDgesvdj(s->h, CUSOLVER_EIG_MODE_VECTOR, 1, m, n, (double*)s->d_a, m, (double*)s->d_w, (double*)s->d_b, m, (double*)s->d_c, n, (double*)s->d_work, lwork, s->d_info, s->jinfo));
    else CS_TRY(cusolverDnZgesvdj(s->h, CUSOLVER_EIG_MODE_VECTOR, 1, m, n, (cuDoubleComplex*)s->d_a, m, (double*)s->d_w, (cuDoubleComplex*)s->d_b, m, (cuDoubleComplex*)s->d_c, n, (cuDoubleComplex*)s->d_work, lwork, s->d_info, s->jinfo));
    char* vt = (char*)s->d_c + (size_t)n * l * es;
    if (dtype == ITB_F64) launch_transpose<double>((const double*)s->d_c, (double*)vt, n, l, s->stream);
    else {
        launch_transpose<double2>((const double2*)s->d_c, (double2*)vt, n, l, s->stream);
        conj_inplace_kernel<<<148, 256, 0, s->stream>>>((double2*)vt, (size_t)n * l);
    }
    int hinfo = 0;
    S_TRY(cudaMemcpyAsync(hU, s->d_b, (size_t)m * l * es, cudaMemcpyDeviceToHost, s->stream));
    S_TRY(cudaMemcpyAsync(hVT, vt, (size_t)l * n * es, cudaMemcpyDeviceToHost, s->stream));
    S_TRY(cudaMemcpyAsync(hS, s->d_w, (size_t)l * 8, cudaMemcpyDeviceToHost, s->stream));
    S_TRY(cudaMemcpyAsync(&hinfo, s->d_info, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    S_TRY(cudaStreamSynchronize(s->stream));
    *info = hinfo;
    return ITB_OK;
}

// polar-decomposition SVD (gesvdp): GEMM-rich, the fast route for large blocks; returns V (n x l)
static int solver_gesvdp(itb_solver* s, int32_t dtype, int32_t m, int32_t n, void* hA, double* hS, void* hU, void* hVT, int32_t* info) {
    const size_t es = dtype == ITB_C64 ? 16 : 8;
    const int l = std::min(m, n);
    const cudaDataType dt = dtype == ITB_C64 ? CUDA_C_64F : CUDA_R_64F;
    int rc = grow(&s->d_a, &s->a_bytes, (size_t)m * n * es); if (rc) return rc;
    rc = grow(&s->d_b, &s->b_bytes, (size_t)m * l * es); if (rc) return rc; // U
    rc = grow(&s->d_c, &s->c_bytes, (size_t)n * l * es * 2); if (rc) return rc; // V, then VT behind it
    rc = grow(&s->d_w, &s->w_bytes, (size_t)l * 8 + 64); if (rc) return rc;
    S_TRY(cudaMemcpyAsync(s->d_a, hA, (size_t)m * n * es, cudaMemcpyHostToDevice, s->stream));
    size_t wd = 0, wh = 0;
    CS_TRY(cusolverDnXgesvdp_bufferSize(s->h, s->params, CUSOLVER_EIG_MODE_VECTOR, 1, m, n, dt, s->d_a, m, CUDA_R_64F, s->d_w, dt, s->d_b, m, dt, s->d_c, n, dt, &wd, &wh));
    rc = grow(&s->d_work, &s->work_bytes, wd + 256); if (rc) return rc;
    if (s->h_work_bytes < wh) { free(s->h_work); s->h_work = malloc(wh + 64); s->h_work_bytes = wh; }
    CS_TRY(cusolverDnXgesvdp(s->h, s->params, CUSOLVER_EIG_MODE_VECTOR, 1, m, n, dt, s->d_a, m, CUDA_R_64F, s->d_w, dt, s->d_b, m, dt, s->d_c, n, dt,
                             s->d_work, wd, s->h_work, wh, s->d_info, &s->last_err_sigma));
    char* vt = (char*)s->d_c + (size_t)n * l * es;
    if (dtype == ITB_F64) launch_transpose<double>((const double*)s->d_c, (double*)vt, n, l, s->stream);
    else {
        launch_transpose<double2>((const double2*)s->d_c, (double2*)vt, n, l, s->stream);
        conj_inplace_kernel<<<148, 256, 0, s->stream>>>((double2*)vt, (size_t)n * l);
    }
    int hinfo = 0;
    S_TRY(cudaMemcpyAsync(hU, s->d_b, (size_t)m * l * es, cudaMemcpyDeviceToHost, s->stream));
    S_TRY(cudaMemcpyAsync(hVT, vt, (size_t)l * n * es, cudaMemcpyDeviceToHost, s->stream));
    S_TRY(cudaMemcpyAsync(hS, s->d_w, (size_t)l * 8, cudaMemcpyDeviceToHost, s->stream));
    S_TRY(cudaMemcpyAsync(&hinfo, s->d_info, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    S_TRY(cudaStreamSynchronize(s->stream));
    *info = hinfo;
    return ITB_OK;
}

int itb_solver_gesvd(itb_solver* s, int32_t dtype, int32_t m, int32_t n, void* hA, double* hS, void* hU, void* hVT, int32_t* info) {
    if (s->svd_method == 1) return solver_gesvdj(s, dtype, m, n, hA, hS, hU, hVT, info);
    if (s->svd_method == 2) return solver_gesvdp(s, dtype, m, n, hA, hS, hU, hVT, info);
    const size_t es = dtype == ITB_C64 ? 16 : 8;
    const int l = std::min(m, n);
    const bool wide = m < n;
    const int M = wide ? n : m, N = wide ? m : n; // the tall problem handed to cuSOLVER
    int rc = grow(&s->d_a, &s->a_bytes, (size_t)m * n * es); if (rc) return rc;
    rc = grow(&s->d_b, &s->b_bytes, (size_t)M * N * es); if (rc) return rc;   // U of the tall problem (M x N)
    rc = grow(&s->d_c, &s->c_bytes, (size_t)M * N * es); if (rc) return rc;   // VT of the tall problem (N x N) / scratch (M x N)
    rc = grow(&s->d_w, &s->w_bytes, (size_t)l * 8 + 64); if (rc) return rc;
    S_TRY(cudaMemcpyAsync(s->d_a, hA, (size_t)m * n * es, cudaMemcpyHostToDevice, s->stream));
    void* dA = s->d_a;
    if (wide) { // A^T (n x m) into d_c, then use it as the tall input
        if (dtype == ITB_F64) launch_transpose<double>((const double*)s->d_a, (double*)s->d_c, m, n, s->stream);
        else launch_transpose<double2>((const double2*)s->d_a, (double2*)s->d_c, m, n, s->stream);
        S_TRY(cudaMemcpyAsync(s->d_a, s->d_c, (size_t)m * n * es, cudaMemcpyDeviceToDevice, s->stream));
    }
    int lwork = 0;
    if (dtype == ITB_F64) CS_TRY(cusolverDnDgesvd_bufferSize(s->h, M, N, &lwork));
    else CS_TRY(cusolverDnZgesvd_bufferSize(s->h, M, N, &lwork));
    rc = grow(&s->d_work, &s->work_bytes, (size_t)lwork * es + (size_t)N * 16); if (rc) return rc;
    // tall problem: At = Ut St VTt, Ut (M x N) -> d_b, VTt (N x N) -> d_c
    if (dtype == ITB_F64)
        CS_TRY(cusolverDnDgesvd(s->h, 'S', 'S', M, N, (double*)dA, M, (double*)s->d_w, (double*)s->d_b, M, (double*)s->d_c, N, (double*)s->d_work, lwork, nullptr, s->d_info));
    else
        CS_TRY(cusolverDnZgesvd(s->h, 'S', 'S', M, N, (cuDoubleComplex*)dA, M, (double*)s->d_w, (cuDoubleComplex*)s->d_b, M, (cuDoubleComplex*)s->d_c, N, (cuDoubleComplex*)s->d_work, lwork, nullptr, s->d_info));
    int hinfo = 0;
    if (!wide) {
        S_TRY(cudaMemcpyAsync(hU, s->d_b, (size_t)m * l * es, cudaMemcpyDeviceToHost, s->stream));
        S_TRY(cudaMemcpyAsync(hVT, s->d_c, (size_t)l * n * es, cudaMemcpyDeviceToHost, s->stream));
    } else {
        // A = (At)^T = (Ut St VTt)^T = VTt^T St Ut^T  (plain transpose; for complex data A^T = conj... see below)
        // real:    U = VTt^T (m x m), VT = Ut^T (m x n)
        // complex: we factorised the PLAIN transpose At = A^T = Ut St VTt^H-form (cuSOLVER returns VTt = Vt^H),
        //          so A = (At)^T = conj(Vt) St Ut^T -> U = conj(Vt) = (VTt)^T, VT = Ut^T: the same two transposes.
        if (dtype == ITB_F64) {
            launch_transpose<double>((const double*)s->d_c, (double*)s->d_a, N, N, s->stream);             // U (m x m)
            S_TRY(cudaMemcpyAsync(hU, s->d_a, (size_t)m * l * es, cudaMemcpyDeviceToHost, s->stream));
            S_TRY(cudaStreamSynchronize(s->stream));
            launch_transpose<double>((const double*)s->d_b, (double*)s->d_a, M, N, s->stream);             // VT (m x n)
            S_TRY(cudaMemcpyAsync(hVT, s->d_a, (size_t)l * n * es, cudaMemcpyDeviceToHost, s->stream));
        } else {
            launch_transpose<double2>((const double2*)s->d_c, (double2*)s->d_a, N, N, s->stream);
            S_TRY(cudaMemcpyAsync(hU, s->d_a, (size_t)m * l * es, cudaMemcpyDeviceToHost, s->stream));
            S_TRY(cudaStreamSynchronize(s->stream));
            launch_transpose<double2>((const double2*)s->d_b, (double2*)s->d_a, M, N, s->stream);
            S_TRY(cudaMemcpyAsync(hVT, s->d_a, (size_t)l * n * es, cudaMemcpyDeviceToHost, s->stream));
        }
    }
    S_TRY(cudaMemcpyAsync(hS, s->d_w, (size_t)l * 8, cudaMemcpyDeviceToHost, s->stream));
    S_TRY(cudaMemcpyAsync(&hinfo, s->d_info, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    S_TRY(cudaStreamSynchronize(s->stream));
    *info = hinfo;
    return ITB_OK;
}

} // extern "C"
