// kernels_blas1.cu — flat-store tasks that keep tensors resident in HBM across Davidson iterations.
//
//   nrm2        doTask(NormNoScale,...) -> dnrm2     itensor/itdata/qdense.cc:409-417, dense.cc:150-162
//   axpy        PlusEQ trivial-permutation fast path  qdense.cc:523-529, dense.cc:380-385 (daxpy)
//   scal        doTask(Mult<Real|Cplx>,...)           qdense.cc:303-325, dense.cc:112-136 (dscal)
//   fill/conj/real_to_cplx/take_part                   qdense.cc:356-380, dense.cc:90-109,167-171
//   dot         rank-0 contraction shortcut used by the host mirror for <V|q>
// All are single-pass, 128-bit vectorised, grid sized to a multiple of the SM count; reductions are
// two-stage and deterministic (fixed partial order, no atomics).
#include <cuda_runtime.h>
#include <stdint.h>

namespace itb {

constexpr int B1_NT = 256;

__global__ void __launch_bounds__(B1_NT) scal_real_kernel(double* __restrict__ x, int64_t n, double a) {
    const int64_t n2 = n >> 1;
    double2* x2 = reinterpret_cast<double2*>(x);
    for (int64_t i = blockIdx.x * (int64_t)B1_NT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * B1_NT) {
        double2 v = x2[i];
        v.x *= a; v.y *= a;
        x2[i] = v;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) x[n - 1] *= a;
}
__global__ void __launch_bounds__(B1_NT) scal_cplx_kernel(double2* __restrict__ x, int64_t n, double ar, double ai) {
    for (int64_t i = blockIdx.x * (int64_t)B1_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * B1_NT) {
        const double2 v = x[i];
        x[i] = make_double2(ar * v.x - ai * v.y, ar * v.y + ai * v.x);
    }
}
__global__ void __launch_bounds__(B1_NT) axpy_real_kernel(const double* __restrict__ x, double* __restrict__ y, int64_t n, double a) {
    const int64_t n2 = n >> 1;
    const double2* x2 = reinterpret_cast<const double2*>(x);
    double2* y2 = reinterpret_cast<double2*>(y);
    for (int64_t i = blockIdx.x * (int64_t)B1_NT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * B1_NT) {
        const double2 u = x2[i];
        double2 v = y2[i];
        v.x = fma(a, u.x, v.x); v.y = fma(a, u.y, v.y);
        y2[i] = v;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) y[n - 1] = fma(a, x[n - 1], y[n - 1]);
}
__global__ void __launch_bounds__(B1_NT) axpy_cplx_kernel(const double2* __restrict__ x, double2* __restrict__ y, int64_t n, double ar, double ai) {
    for (int64_t i = blockIdx.x * (int64_t)B1_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * B1_NT) {
        const double2 u = x[i];
        double2 v = y[i];
        v.x += ar * u.x - ai * u.y;
        v.y += ar * u.y + ai * u.x;
        y[i] = v;
    }
}
__global__ void __launch_bounds__(B1_NT) fill_kernel(double* __restrict__ x, int64_t nreal, double re, double im, int cplx) {
    for (int64_t i = blockIdx.x * (int64_t)B1_NT + threadIdx.x; i < nreal; i += (int64_t)gridDim.x * B1_NT)
        x[i] = (cplx && (i & 1)) ? im : re;
}
__global__ void __launch_bounds__(B1_NT) conj_kernel(double2* __restrict__ x, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)B1_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * B1_NT) x[i].y = -x[i].y;
}
__global__ void __launch_bounds__(B1_NT) r2c_kernel(const double* __restrict__ x, double2* __restrict__ y, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)B1_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * B1_NT) y[i] = make_double2(x[i], 0.0);
}
__global__ void __launch_bounds__(B1_NT) part_kernel(const double2* __restrict__ x, double* __restrict__ y, int64_t n, int imag) {
    for (int64_t i = blockIdx.x * (int64_t)B1_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * B1_NT) y[i] = imag ? x[i].y : x[i].x;
}

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double red[B1_NT / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < B1_NT / 32; ++w) s += red[w];
    __syncthreads();
    return s; // valid on thread 0
}

// partial[b] = sum over this CTA's slice of (x*scale)^2 ; partial[grid+b] = max|x|
__global__ void __launch_bounds__(B1_NT) ssq_kernel(const double* __restrict__ x, int64_t n, double scale, double* __restrict__ partial) {
    double s = 0.0, mx = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)B1_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * B1_NT) {
        const double v = x[i] * scale;
        s = fma(v, v, s);
        mx = fmax(mx, fabs(v));
    }
    const double t = block_sum(s);
    __shared__ double mred[B1_NT / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) mred[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < B1_NT / 32; ++w) m = fmax(m, mred[w]);
        partial[blockIdx.x] = t;
        partial[gridDim.x + blockIdx.x] = m;
    }
}
// out[0] = sum partial[0..g), out[1] = max partial[g..2g)
__global__ void __launch_bounds__(B1_NT) ssq_finish_kernel(const double* __restrict__ partial, int g, double* __restrict__ out) {
    double s = 0.0, mx = 0.0;
    for (int i = threadIdx.x; i < g; i += B1_NT) { s += partial[i]; mx = fmax(mx, partial[g + i]); }
    const double t = block_sum(s);
    __shared__ double mred[B1_NT / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) mred[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < B1_NT / 32; ++w) m = fmax(m, mred[w]);
        out[0] = t;
        out[1] = m;
    }
}

// partial[2b],[2b+1] = re,im of sum conj?(x)*y over the CTA's slice
__global__ void __launch_bounds__(B1_NT) dot_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n, int cplx,
                                                    int conj_x, double* __restrict__ partial) {
    double sr = 0.0, si = 0.0;
    if (!cplx) {
        for (int64_t i = blockIdx.x * (int64_t)B1_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * B1_NT) sr = fma(x[i], y[i], sr);
    } else {
        const double2* x2 = reinterpret_cast<const double2*>(x);
        const double2* y2 = reinterpret_cast<const double2*>(y);
        const double sg = conj_x ? -1.0 : 1.0;
        for (int64_t i = blockIdx.x * (int64_t)B1_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * B1_NT) {
            const double2 u = x2[i], v = y2[i];
            const double ui = sg * u.y;
            sr += u.x * v.x - ui * v.y;
            si += u.x * v.y + ui * v.x;
        }
    }
    const double tr = block_sum(sr);
    const double ti = block_sum(si);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = tr; partial[2 * blockIdx.x + 1] = ti; }
}
__global__ void __launch_bounds__(B1_NT) dot_finish_kernel(const double* __restrict__ partial, int g, double* __restrict__ out) {
    double sr = 0.0, si = 0.0;
    for (int i = threadIdx.x; i < g; i += B1_NT) { sr += partial[2 * i]; si += partial[2 * i + 1]; }
    const double tr = block_sum(sr);
    const double ti = block_sum(si);
    if (threadIdx.x == 0) { out[0] = tr; out[1] = ti; }
}

static inline int grid_for(int64_t n, int num_sms) {
    int64_t g = (n + B1_NT - 1) / B1_NT;
    const int64_t cap = (int64_t)num_sms * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

cudaError_t launch_scal(int cplx, int64_t n, void* x, double ar, double ai, int sms, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (!cplx) scal_real_kernel<<<grid_for(n / 2 + 1, sms), B1_NT, 0, st>>>((double*)x, n, ar);
    else if (ai == 0.0) scal_real_kernel<<<grid_for(n + 1, sms), B1_NT, 0, st>>>((double*)x, 2 * n, ar);
    else scal_cplx_kernel<<<grid_for(n, sms), B1_NT, 0, st>>>((double2*)x, n, ar, ai);
    return cudaGetLastError();
}
cudaError_t launch_axpy(int cplx, int64_t n, double ar, double ai, const void* x, void* y, int sms, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (!cplx) axpy_real_kernel<<<grid_for(n / 2 + 1, sms), B1_NT, 0, st>>>((const double*)x, (double*)y, n, ar);
    else if (ai == 0.0) axpy_real_kernel<<<grid_for(n + 1, sms), B1_NT, 0, st>>>((const double*)x, (double*)y, 2 * n, ar);
    else axpy_cplx_kernel<<<grid_for(n, sms), B1_NT, 0, st>>>((const double2*)x, (double2*)y, n, ar, ai);
    return cudaGetLastError();
}
cudaError_t launch_fill(int cplx, int64_t n, void* x, double re, double im, int sms, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int64_t nr = cplx ? 2 * n : n;
    fill_kernel<<<grid_for(nr, sms), B1_NT, 0, st>>>((double*)x, nr, re, im, cplx);
    return cudaGetLastError();
}
cudaError_t launch_conj(int64_t n, void* x, int sms, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    conj_kernel<<<grid_for(n, sms), B1_NT, 0, st>>>((double2*)x, n);
    return cudaGetLastError();
}
cudaError_t launch_r2c(int64_t n, const void* x, void* y, int sms, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    r2c_kernel<<<grid_for(n, sms), B1_NT, 0, st>>>((const double*)x, (double2*)y, n);
    return cudaGetLastError();
}
cudaError_t launch_part(int64_t n, const void* x, void* y, int imag, int sms, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    part_kernel<<<grid_for(n, sms), B1_NT, 0, st>>>((const double2*)x, (double*)y, n, imag);
    return cudaGetLastError();
}
// scratch: >= 2*grid+2 doubles. Result in scratch[2*grid], scratch[2*grid+1] (ssq, max).
cudaError_t launch_ssq(int64_t nreal, const void* x, double scale, double* scratch, int* grid_out, int sms, cudaStream_t st) {
    const int g = grid_for(nreal, sms);
    *grid_out = g;
    ssq_kernel<<<g, B1_NT, 0, st>>>((const double*)x, nreal, scale, scratch);
    ssq_finish_kernel<<<1, B1_NT, 0, st>>>(scratch, g, scratch + 2 * g);
    return cudaGetLastError();
}
cudaError_t launch_dot1(int cplx, int64_t n, const void* x, const void* y, int conj_x, double* scratch, int* grid_out, int sms,
                        cudaStream_t st) {
    const int g = grid_for(n, sms);
    *grid_out = g;
    dot_kernel<<<g, B1_NT, 0, st>>>((const double*)x, (const double*)y, n, cplx, conj_x, scratch);
    dot_finish_kernel<<<1, B1_NT, 0, st>>>(scratch, g, scratch + 2 * g);
    return cudaGetLastError();
}

} // namespace itb
