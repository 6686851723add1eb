//
// lapack_gpu.cc — device-backed eigh / SVD behind the reference's LAPACK wrapper boundary.
//
// The reference funnels every LAPACK call through itensor/tensor/lapack_wrap.{h,cc}. The plugin build
// compiles that file UNMODIFIED but with -Ddsyev_wrapper=dsyev_wrapper_host (etc., see Makefile), so the
// reference's own implementations keep existing under *_host names, and the names the rest of the library
// calls (hermitianDiag algs.cc:34-47, SVDRefLAPACK algs_impl.h:356-420) resolve to the dispatchers below:
// eigh blocks with n >= ITB_EIGH_MIN_N (default 256) and SVD blocks with min(m,n) >= ITB_SVD_MIN_N (default 320)
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
// cuSOLVER syevd beats host LAPACK from n~256 (1024: 17 ms vs 90 ms); for the SVD the polar-decomposition solver
// (Xgesvdp, GEMM-rich) is the one that wins: 1268^2 in 62 ms vs 313 ms host gesdd on 8 threads, same absolute
// accuracy (|ds|/s0 1e-15, orthogonality 4e-15); gesvd (QR iteration) is 2.6x slower than that and gesvdj slower
// still (tools/solver_bench3.py). First use of each cuSOLVER kernel pays seconds of
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
svdMinN() { static long v = envLong("ITB_SVD_MIN_N",320); return v; }

// ITB_PROFILE: wall time spent behind the LAPACK boundary, by route, printed at exit
struct SolverProf
    {
    double secs[4] = {0,0,0,0}; long calls[4] = {0,0,0,0};
    bool on = std::getenv("ITB_PROFILE") != nullptr;
    ~SolverProf()
        {
        if(!on) return;
        const char* names[4] = {"eigh device","eigh host","svd device","svd host"};
        for(int i = 0; i < 4; ++i) std::fprintf(stderr,"[itensor_b200 profile] %-28s %10ld %12.4f\n",names[i],calls[i],secs[i]);
        }
    };
static SolverProf& sprof() { static SolverProf p; return p; }
struct SolverScope
    {
    int slot; std::chrono::steady_clock::time_point t0;
    explicit SolverScope(int s) : slot(s), t0(std::chrono::steady_clock::now()) { }
    ~SolverScope() { auto& p = sprof(); p.secs[slot] += std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count(); p.calls[slot] += 1; }
    };

static void
checkSolver(int rc, const char* what)
    {
    if(rc != ITB_OK) throw ITError(tinyformat::format("itensor_b200 (%s): %s",what,itb_last_error()));
    }

void
dsyev_wrapper(char jobz, char uplo, LAPACK_INT n, LAPACK_REAL* A, LAPACK_REAL* eigs, LAPACK_INT& info)
    {
    if(n < eighMinN() || jobz != 'V' || uplo != 'U' || !itb_solver_ready())
        {
        SolverScope sc(1);
        dsyev_wrapper_host(jobz,uplo,n,A,eigs,info);
        return;
        }
    SolverScope sc(0);
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
    if(N < eighMinN() || !itb_solver_ready()) { SolverScope sc(1); return zheev_wrapper_host(N,A,d); }
    SolverScope sc(0);
    int32_t inf = 0;
    checkSolver(itb_syevd_host(gpu::context(),ITB_C64,N,A,d,&inf),"heevd");
    return inf;
    }

void
dgesdd_wrapper(char* jobz, LAPACK_INT* m, LAPACK_INT* n, LAPACK_REAL* A, LAPACK_REAL* s, LAPACK_REAL* u, LAPACK_REAL* vt, LAPACK_INT* info)
    {
    if(std::min(*m,*n) < svdMinN() || *jobz != 'S' || !itb_solver_ready())
        {
        SolverScope sc(3);
        dgesdd_wrapper_host(jobz,m,n,A,s,u,vt,info);
        return;
        }
    SolverScope sc(2);
    int32_t inf = 0;
    checkSolver(itb_gesvd_host(gpu::context(),ITB_F64,*m,*n,A,s,u,vt,&inf),"gesvd");
    *info = inf;
    }

void
zgesdd_wrapper(char* jobz, LAPACK_INT* m, LAPACK_INT* n, Cplx* A, LAPACK_REAL* s, Cplx* u, Cplx* vt, LAPACK_INT* info)
    {
    if(std::min(*m,*n) < svdMinN() || *jobz != 'S' || !itb_solver_ready())
        {
        SolverScope sc(3);
        zgesdd_wrapper_host(jobz,m,n,A,s,u,vt,info);
        return;
        }
    SolverScope sc(2);
    int32_t inf = 0;
    checkSolver(itb_gesvd_host(gpu::context(),ITB_C64,*m,*n,A,s,u,vt,&inf),"gesvd");
    *info = inf;
    }

} //namespace itensor
