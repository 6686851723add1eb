//
// lapack_gpu.cc — device-backed eigh / SVD behind the reference's LAPACK wrapper boundary.
//
// The reference funnels every LAPACK call through itensor/tensor/lapack_wrap.{h,cc}. The plugin build
// compiles that file UNMODIFIED but with -Ddsyev_wrapper=dsyev_wrapper_host (etc., see Makefile), so the
// reference's own implementations keep existing under *_host names, and the names the rest of the library
// calls (hermitianDiag algs.cc:34-47, SVDRefLAPACK algs_impl.h:356-420) resolve to the dispatchers below:
// eigh blocks with n >= ITB_EIGH_MIN_N (default 256) and SVD blocks with min(m,n) >= ITB_SVD_MIN_N (default: never)
// go to cuSOLVER through the C ABI
// (itb_syevd_host / itb_gesvd_host), smaller ones stay on the host LAPACK where a GPU round trip cannot pay.
// The per-block loops, sorting and truncation (hermitian.cc:231-358, svd.cc:199-314, decomp.cc:306-463) remain
// the reference's host code, so kept spectra follow its rules exactly.
//
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "itensor/tensor/lapack_wrap.h"
#include "itensor/util/error.h"
#include "itensor/util/print.h"
#include "itb200.h"

struct itb_ctx;

namespace itensor {

namespace gpu { itb_ctx* context(); }

// the reference's implementations (lapack_wrap.cc compiled with renamed symbols)
void dsyev_wrapper_host(char jobz, char uplo, LAPACK_INT n, LAPACK_REAL* A, LAPACK_REAL* eigs, LAPACK_INT& info);
LAPACK_INT zheev_wrapper_host(LAPACK_INT N, Cplx* A, LAPACK_REAL* d);
void dgesdd_wrapper_host(char* jobz, LAPACK_INT* m, LAPACK_INT* n, LAPACK_REAL* A, LAPACK_REAL* s, LAPACK_REAL* u, LAPACK_REAL* vt, LAPACK_INT* info);
void zgesdd_wrapper_host(char* jobz, LAPACK_INT* m, LAPACK_INT* n, Cplx* A, LAPACK_REAL* s, Cplx* u, Cplx* vt, LAPACK_INT* info);

// thresholds (block dimension from which a call goes to the device); measured on B200 (profiles/README.md):
// cuSOLVER syevd beats host LAPACK from n~256 (1024: 17 ms vs 90 ms); gesvd/gesvdj lose to host gesdd on DMRG
// blocks up to several hundred, so the SVD path is opt-in. First use of each cuSOLVER kernel pays seconds of
// lazy module loading (a few seconds per process; do not use CUDA_MODULE_LOADING=EAGER: measured 277 s).
static long
envLong(const char* name, long dflt)
    {
    if(auto* e = std::getenv(name)) return std::atol(e);
    return dflt;
    }
static long
eighMinN() { static long v = envLong("ITB_EIGH_MIN_N",256); return v; }
static long
svdMinN() { static long v = envLong("ITB_SVD_MIN_N",1l<<40); return v; }

static void
checkSolver(int rc, const char* what)
    {
    if(rc != ITB_OK) throw ITError(tinyformat::format("itensor_b200 (%s): %s",what,itb_last_error()));
    }

void
dsyev_wrapper(char jobz, char uplo, LAPACK_INT n, LAPACK_REAL* A, LAPACK_REAL* eigs, LAPACK_INT& info)
    {
    if(n < eighMinN() || jobz != 'V' || uplo != 'U')
        {
        dsyev_wrapper_host(jobz,uplo,n,A,eigs,info);
        return;
        }
    int32_t inf = 0;
    auto t0 = std::chrono::steady_clock::now();
    checkSolver(itb_syevd_host(gpu::context(),ITB_F64,n,A,eigs,&inf),"syevd");
    info = inf;
    if(std::getenv("ITB_SOLVER_TRACE"))
        {
        auto ms = std::chrono::duration<double,std::milli>(std::chrono::steady_clock::now()-t0).count();
        std::fprintf(stderr,"[itensor_b200 solver] syevd n=%d %.2f ms info=%d\n",int(n),ms,int(inf));
        }
    }

LAPACK_INT
zheev_wrapper(LAPACK_INT N, Cplx* A, LAPACK_REAL* d)
    {
    if(N < eighMinN()) return zheev_wrapper_host(N,A,d);
    int32_t inf = 0;
    checkSolver(itb_syevd_host(gpu::context(),ITB_C64,N,A,d,&inf),"heevd");
    return inf;
    }

void
dgesdd_wrapper(char* jobz, LAPACK_INT* m, LAPACK_INT* n, LAPACK_REAL* A, LAPACK_REAL* s, LAPACK_REAL* u, LAPACK_REAL* vt, LAPACK_INT* info)
    {
    if(std::min(*m,*n) < svdMinN() || *jobz != 'S')
        {
        dgesdd_wrapper_host(jobz,m,n,A,s,u,vt,info);
        return;
        }
    int32_t inf = 0;
    checkSolver(itb_gesvd_host(gpu::context(),ITB_F64,*m,*n,A,s,u,vt,&inf),"gesvd");
    *info = inf;
    }

void
zgesdd_wrapper(char* jobz, LAPACK_INT* m, LAPACK_INT* n, Cplx* A, LAPACK_REAL* s, Cplx* u, Cplx* vt, LAPACK_INT* info)
    {
    if(std::min(*m,*n) < svdMinN() || *jobz != 'S')
        {
        zgesdd_wrapper_host(jobz,m,n,A,s,u,vt,info);
        return;
        }
    int32_t inf = 0;
    checkSolver(itb_gesvd_host(gpu::context(),ITB_C64,*m,*n,A,s,u,vt,&inf),"gesvd");
    *info = inf;
    }

} //namespace itensor
