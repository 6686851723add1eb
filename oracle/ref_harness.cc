// ref_harness.cc — C-ABI test harness around the UNMODIFIED reference (ITensor C++ v3).
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header). Our own code, compiled against the reference
// headers where they lie (/root/reference) and linked with oracle/_ref/libitensor.a into
// oracle/_ref/libitref.so by oracle/Makefile. It drives the reference's PUBLIC API only
// (Index / ITensor / operator* / permute / operator+= / norm / dmrg) so that what is checked is the
// reference's own doTask(Contract|Order|PlusEQ|NormNoScale, QDense|Dense) path.
#include <chrono>
#include <cstring>
#include <map>
#include <vector>

#include "itensor/all.h"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace itensor;

// OpenBLAS' own thread-count setter (the BLAS this harness links): lets the caller pin the three CPU modes of
// SURVEY §8(d) explicitly instead of inheriting OMP_NUM_THREADS / OPENBLAS_NUM_THREADS from a launcher (torchrun
// exports OMP_NUM_THREADS=1 to every rank).
extern "C" void openblas_set_num_threads(int);
extern "C" int openblas_get_num_threads(void);

extern "C" {

// Flat description of an ITensor for crossing the C boundary.
struct ref_tensor {
    int32_t order;
    int32_t dtype;          // 0 real, 1 complex
    int32_t nqn;            // QN components per sector; 0 => no QNs (Dense storage)
    const int64_t* labels;  // [order] index identity: equal label <=> same Index
    const int32_t* dirs;    // [order] +1 Out / -1 In
    const int32_t* nsect;   // [order]
    const int64_t* sect;    // concatenated sector sizes
    const int32_t* qn;      // [(sum nsect) * nqn]
    const int32_t* mods;    // [nqn]
    int64_t nblocks;
    const int32_t* blocks;  // [nblocks*order]
    const int64_t* offsets; // unused on input (reference recomputes), checked by tests on output
    int64_t nelems;
    const double* data;     // nelems (x2 if complex) doubles, laid out block after block
};

struct ref_result {
    std::vector<int64_t> labels;
    std::vector<int32_t> blocks;
    std::vector<int64_t> offsets;
    std::vector<double> data;
    int32_t order = 0, dtype = 0, is_qn = 0;
    int64_t nblocks = 0, nelems = 0;
    std::vector<int32_t> flux;
    std::string err;
};

} // extern "C"

namespace {

struct Registry {
    std::map<int64_t, Index> byLabel;
    std::map<Index::id_type, int64_t> labelOf;
};

Index
makeIndex(ref_tensor const& t, int j, Registry& reg)
    {
    auto it = reg.byLabel.find(t.labels[j]);
    auto dir = t.dirs[j] > 0 ? Out : In;
    if(it != reg.byLabel.end())
        {
        auto I = it->second;
        if(t.nqn > 0 && I.dir() != dir) I.dag();
        return I;
        }
    long start = 0;
    for(int i = 0; i < j; ++i) start += t.nsect[i];
    Index I;
    if(t.nqn == 0)
        {
        I = Index(t.sect[start],tinyformat::format("L%d",t.labels[j]));
        }
    else
        {
        auto qns = Index::qnstorage(t.nsect[j]);
        for(int s = 0; s < t.nsect[j]; ++s)
            {
            QN q;
            for(int c = 0; c < t.nqn; ++c)
                {
                q.addNum(QNum(tinyformat::format("q%d",c),t.qn[(start+s)*t.nqn+c],t.mods[c]));
                }
            qns[s] = std::make_pair(q,long(t.sect[start+s]));
            }
        I = Index(std::move(qns),dir,tinyformat::format("L%d",t.labels[j]));
        }
    reg.byLabel[t.labels[j]] = I;
    reg.labelOf[I.id()] = t.labels[j];
    return I;
    }

ITensor
makeTensor(ref_tensor const& t, Registry& reg)
    {
    auto inds = std::vector<Index>();
    for(int j = 0; j < t.order; ++j) inds.push_back(makeIndex(t,j,reg));
    auto is = IndexSet(inds);
    if(t.nqn == 0)
        {
        if(t.dtype == 0)
            {
            auto d = DenseReal(t.nelems);
            std::copy(t.data,t.data+t.nelems,d.store.begin());
            return ITensor(is,std::move(d));
            }
        auto d = DenseCplx(t.nelems);
        std::memcpy(d.store.data(),t.data,sizeof(double)*2*t.nelems);
        return ITensor(is,std::move(d));
        }
    auto blocks = Blocks(t.nblocks,Block(t.order));
    for(long b = 0; b < t.nblocks; ++b)
    for(int j = 0; j < t.order; ++j)
        {
        blocks[b][j] = t.blocks[b*t.order+j];
        }
    if(t.dtype == 0)
        {
        auto d = QDenseReal(is,blocks);
        if(long(d.store.size()) != t.nelems) Error("ref_harness: nelems mismatch");
        std::copy(t.data,t.data+t.nelems,d.store.begin());
        return ITensor(is,std::move(d));
        }
    auto d = QDenseCplx(is,blocks);
    if(long(d.store.size()) != t.nelems) Error("ref_harness: nelems mismatch");
    std::memcpy(d.store.data(),t.data,sizeof(double)*2*t.nelems);
    return ITensor(is,std::move(d));
    }

struct Extract
    {
    ref_result* r;
    template<typename T>
    void
    fill(std::vector<T> const&) { }
    void
    operator()(QDenseReal const& d) { r->is_qn = 1; r->dtype = 0; blocksOf(d); r->data.assign(d.store.begin(),d.store.end()); }
    void
    operator()(QDenseCplx const& d) { r->is_qn = 1; r->dtype = 1; blocksOf(d); cplx(d.store.data(),d.store.size()); }
    void
    operator()(DenseReal const& d) { r->dtype = 0; r->nelems = d.store.size(); r->data.assign(d.store.begin(),d.store.end()); }
    void
    operator()(DenseCplx const& d) { r->dtype = 1; r->nelems = d.store.size(); cplx(d.store.data(),d.store.size()); }
    void
    operator()(ScalarReal const& d) { r->dtype = 0; r->nelems = 1; r->data.assign(1,d.val); }
    void
    operator()(ScalarCplx const& d) { r->dtype = 1; r->nelems = 1; r->data = {d.val.real(),d.val.imag()}; }
    template<typename T>
    void
    blocksOf(QDense<T> const& d)
        {
        r->nblocks = d.offsets.size();
        r->nelems = d.store.size();
        for(auto const& bo : d.offsets)
            {
            for(int j = 0; j < r->order; ++j) r->blocks.push_back(bo.block[j]);
            r->offsets.push_back(bo.offset);
            }
        }
    void
    cplx(Cplx const* p, size_t n)
        {
        r->data.resize(2*n);
        std::memcpy(r->data.data(),p,sizeof(double)*2*n);
        }
    };

ref_result*
extract(ITensor const& T, Registry const& reg)
    {
    auto* r = new ref_result();
    r->order = T.order();
    for(auto const& I : T.inds())
        {
        auto it = reg.labelOf.find(I.id());
        r->labels.push_back(it == reg.labelOf.end() ? -1 : it->second);
        }
    if(T.store()) applyFunc(Extract{r},T.store());
    if(hasQNs(T) && T.store() && r->nblocks > 0)
        {
        auto q = flux(T);
        for(size_t n = 1; n <= QNSize(); ++n)
            {
            if(q.num(n)) r->flux.push_back(q.val(n));
            }
        }
    return r;
    }

} // namespace

extern "C" {

ref_result*
ref_contract(ref_tensor const* A, ref_tensor const* B)
    {
    Registry reg;
    auto a = makeTensor(*A,reg);
    auto b = makeTensor(*B,reg);
    auto c = a*b;
    return extract(c,reg);
    }

// best-of-reps wall time (seconds) of the reference's operator* on the given pair
double
ref_time_contract(ref_tensor const* A, ref_tensor const* B, int reps)
    {
    Registry reg;
    auto a = makeTensor(*A,reg);
    auto b = makeTensor(*B,reg);
    double best = 1e300;
    for(int i = 0; i < reps; ++i)
        {
        auto t0 = std::chrono::steady_clock::now();
        auto c = a*b;
        auto t1 = std::chrono::steady_clock::now();
        best = std::min(best,std::chrono::duration<double>(t1-t0).count());
        if(!c.store()) return -1.;
        }
    return best;
    }

// One effective-Hamiltonian product the way LocalOp::product does it (mps/localop.h:346-362):
// phip = phi*L; phip *= W1; phip *= W2; phip *= R.  Returns best-of-reps seconds for the chain and
// (optionally) the result of the last repetition.
double
ref_time_heff(ref_tensor const* phi, ref_tensor const* L, ref_tensor const* W1, ref_tensor const* W2,
              ref_tensor const* R, int reps, ref_result** out)
    {
    Registry reg;
    auto p = makeTensor(*phi,reg);
    auto l = makeTensor(*L,reg);
    auto w1 = makeTensor(*W1,reg);
    auto w2 = makeTensor(*W2,reg);
    auto r = makeTensor(*R,reg);
    double best = 1e300;
    ITensor res;
    for(int i = 0; i < reps; ++i)
        {
        auto t0 = std::chrono::steady_clock::now();
        auto phip = p*l;
        phip *= w1;
        phip *= w2;
        phip *= r;
        auto t1 = std::chrono::steady_clock::now();
        best = std::min(best,std::chrono::duration<double>(t1-t0).count());
        res = phip;
        }
    if(out) *out = extract(res,reg);
    return best;
    }

// permute T so that its indices appear in the order new_labels (reference ITensor::permute)
ref_result*
ref_permute(ref_tensor const* T, int64_t const* new_labels)
    {
    Registry reg;
    auto t = makeTensor(*T,reg);
    auto inds = std::vector<Index>();
    for(int j = 0; j < T->order; ++j)
        {
        auto I = reg.byLabel.at(new_labels[j]);
        for(auto const& J : t.inds()) if(J == I) { inds.push_back(J); break; }
        }
    t.permute(IndexSet(inds));
    return extract(t,reg);
    }

double
ref_time_permute(ref_tensor const* T, int64_t const* new_labels, int reps)
    {
    Registry reg;
    auto t0 = makeTensor(*T,reg);
    auto inds = std::vector<Index>();
    for(int j = 0; j < T->order; ++j)
        {
        auto I = reg.byLabel.at(new_labels[j]);
        for(auto const& J : t0.inds()) if(J == I) { inds.push_back(J); break; }
        }
    double best = 1e300;
    for(int i = 0; i < reps; ++i)
        {
        auto t = t0;
        auto c0 = std::chrono::steady_clock::now();
        t.permute(IndexSet(inds));
        auto c1 = std::chrono::steady_clock::now();
        best = std::min(best,std::chrono::duration<double>(c1-c0).count());
        }
    return best;
    }

// A += (alpha_re + i alpha_im) * B   (reference operator+= / daxpy path incl. permutation, block merge)
ref_result*
ref_pluseq(ref_tensor const* A, ref_tensor const* B, double alpha_re, double alpha_im)
    {
    Registry reg;
    auto a = makeTensor(*A,reg);
    auto b = makeTensor(*B,reg);
    if(alpha_im == 0.) a += alpha_re*b;
    else a += Cplx(alpha_re,alpha_im)*b;
    return extract(a,reg);
    }

double
ref_norm(ref_tensor const* T)
    {
    Registry reg;
    return norm(makeTensor(*T,reg));
    }

// blas_threads > 0: OpenBLAS threads; omp_threads > 0: threads of the reference's own OpenMP loop over C blocks
// (itensor/itdata/qutil.h:285-348; only in the -DITENSOR_USE_OMP build, libitref_omp.so). Returns 1 when that loop exists.
int
ref_set_threads(int blas_threads, int omp_threads)
    {
    if(blas_threads > 0) openblas_set_num_threads(blas_threads);
#if defined(_OPENMP) && defined(ITENSOR_USE_OMP)
    if(omp_threads > 0) omp_set_num_threads(omp_threads);
    return 1;
#else
    (void)omp_threads;
    return 0;
#endif
    }
int
ref_get_blas_threads(void) { return openblas_get_num_threads(); }
int
ref_get_omp_threads(void)
    {
#if defined(_OPENMP) && defined(ITENSOR_USE_OMP)
    return omp_get_max_threads();
#else
    return 0;
#endif
    }

// accessors for ref_result
int32_t ref_result_order(ref_result const* r) { return r->order; }
int32_t ref_result_dtype(ref_result const* r) { return r->dtype; }
int32_t ref_result_is_qn(ref_result const* r) { return r->is_qn; }
int64_t ref_result_nblocks(ref_result const* r) { return r->nblocks; }
int64_t ref_result_nelems(ref_result const* r) { return r->nelems; }
int32_t ref_result_nflux(ref_result const* r) { return (int32_t)r->flux.size(); }
void ref_result_labels(ref_result const* r, int64_t* o) { std::copy(r->labels.begin(),r->labels.end(),o); }
void ref_result_blocks(ref_result const* r, int32_t* o) { std::copy(r->blocks.begin(),r->blocks.end(),o); }
void ref_result_offsets(ref_result const* r, int64_t* o) { std::copy(r->offsets.begin(),r->offsets.end(),o); }
void ref_result_data(ref_result const* r, double* o) { std::copy(r->data.begin(),r->data.end(),o); }
void ref_result_flux(ref_result const* r, int32_t* o) { std::copy(r->flux.begin(),r->flux.end(),o); }
void ref_result_free(ref_result* r) { delete r; }

//
// Heisenberg-chain DMRG through the reference's own dmrg() (sample/dmrg.cc logic).
// spin2 = 1 -> SpinHalf, 2 -> SpinOne. Per-sweep schedule arrays have nsweeps entries.
// Outputs: energy after the last sweep, total wall seconds, max link dimension.
//
int
ref_dmrg_heisenberg(int N, int spin2, int conserve_qns, int nsweeps,
                    int const* maxdim, double const* cutoff, int const* niter, double const* noise,
                    double* energy_out, double* seconds_out, int* maxlink_out)
    {
    SiteSet sites;
    if(spin2 == 1) sites = SpinHalf(N,{"ConserveQNs=",conserve_qns != 0});
    else           sites = SpinOne(N,{"ConserveQNs=",conserve_qns != 0});
    auto ampo = AutoMPO(sites);
    for(auto j : range1(N-1))
        {
        ampo += 0.5,"S+",j,"S-",j+1;
        ampo += 0.5,"S-",j,"S+",j+1;
        ampo +=     "Sz",j,"Sz",j+1;
        }
    auto H = toMPO(ampo);
    auto state = InitState(sites);
    for(auto i : range1(N))
        {
        if(i%2 == 1) state.set(i,"Up");
        else         state.set(i,"Dn");
        }
    auto psi0 = MPS(state);
    auto sweeps = Sweeps(nsweeps);
    for(int s = 1; s <= nsweeps; ++s)
        {
        sweeps.setmaxdim(s,maxdim[s-1]);
        sweeps.setcutoff(s,cutoff[s-1]);
        sweeps.setniter(s,niter[s-1]);
        sweeps.setnoise(s,noise[s-1]);
        }
    auto t0 = std::chrono::steady_clock::now();
    auto [energy,psi] = dmrg(H,psi0,sweeps,{"Silent",true});
    auto t1 = std::chrono::steady_clock::now();
    *energy_out = energy;
    *seconds_out = std::chrono::duration<double>(t1-t0).count();
    *maxlink_out = maxLinkDim(psi);
    return 0;
    }

} // extern "C"
