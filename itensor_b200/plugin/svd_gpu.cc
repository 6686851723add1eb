//
// svd_gpu.cc — svdOrd2 for HBM-resident block-sparse tensors.
//
// The reference funnels every ITensor-level SVD (svd(), and MPS::svdBond when noise == 0) through
//     Spectrum svdOrd2(ITensor const& A, Index const& uI, Index const& vI, ITensor& U, ITensor& D, ITensor& V, Args)
// (itensor/svd.cc:429-463 -> svdImpl<T> :41-427). The plugin build compiles svd.cc UNMODIFIED but with
// -DsvdOrd2=svdOrd2_host, so the reference's implementation keeps existing under that name and the symbol the
// rest of the library calls (decomp.cc:242) resolves to the dispatcher at the bottom of this file:
//   * A stored as QDenseGPU<Real|Cplx>  -> svdBlocksGPU below: every block is factorised on the device straight
//     from A's buffer (itb_svd_batch_run: cuSOLVER polar / Jacobi SVD over several streams), only the singular
//     values come back to the host, the reference's own truncate() (decomp.cc:306-463) decides what is kept, and
//     the kept columns are copied device-to-device into new QDenseGPU U and V tensors. Nothing but O(m) numbers
//     crosses PCIe and U, V are born in HBM, so the next phi = A(b)*A(b+1) needs no upload.
//   * anything else (host storage, dense GPU storage, ComputeQNs / ShowEigs requests) -> svdOrd2_host.
// Index / QN bookkeeping of the result (new link indices, block lists, scale and sign conventions) follows
// svdImpl's QN branch (svd.cc:169-422) so that U, D, V are interchangeable with the reference's.
//
#include <algorithm>
#include <chrono>
#include <thread>
#include <cstdio>
#include <functional>
#include <vector>

#include "itensor/decomp.h"
#include "itensor/tensor/algs.h"
#include "itensor/itdata/qutil.h"
#include "itensor/util/print_macro.h"
#include "gpu_convert.h"
#include "itb200.h"

struct itb_ctx;

namespace itensor {

namespace gpu { itb_ctx* context(); }

// the reference's implementation (svd.cc compiled with the symbol renamed)
Spectrum
svdOrd2_host(ITensor const& A, Index const& uI, Index const& vI, ITensor & U, ITensor & D, ITensor & V, Args args);

namespace {

void
checkSvd(int rc, const char* what)
    {
    if(rc != ITB_OK) throw ITError(tinyformat::format("itensor_b200 (%s): %s",what,itb_last_error()));
    }

template<typename T> int32_t dtypeFor();
template<> int32_t dtypeFor<Real>() { return ITB_F64; }
template<> int32_t dtypeFor<Cplx>() { return ITB_C64; }

// ITB_PROFILE: wall time of the phases of the device svdOrd2, printed at exit
struct SvdProf
    {
    double secs[5] = {0,0,0,0,0}; long calls = 0;
    bool on = std::getenv("ITB_PROFILE") != nullptr;
    ~SvdProf()
        {
        if(!on || !calls) return;
        const char* names[5] = {"svdOrd2 launch device","svdOrd2 host blocks","svdOrd2 wait device","svdOrd2 truncate+index","svdOrd2 assemble U,V"};
        for(int i = 0; i < 5; ++i) std::fprintf(stderr,"[itensor_b200 profile] %-28s %10ld %12.4f\n",names[i],calls,secs[i]);
        }
    };
SvdProf& svdProf() { static SvdProf p; return p; }
struct Lap
    {
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void mark(int slot) { auto n = std::chrono::steady_clock::now(); svdProf().secs[slot] += std::chrono::duration<double>(n-t).count(); t = n; }
    };

struct BatchGuard
    {
    itb_svd_batch* b = nullptr;
    ~BatchGuard() { if(b) itb_svd_batch_destroy(b); }
    };

template<typename T>
Spectrum
svdBlocksGPU(ITensor const& A, QDenseGPU<T> const& d, Index const& uI, Index const& vI,
             ITensor & U, ITensor & D, ITensor & V, Args const& args)
    {
    auto do_truncate = args.getBool("Truncate");
    auto cutoff = args.getReal("Cutoff",MIN_CUT);
    auto maxdim = args.getInt("MaxDim",args.getInt("Maxm",MAX_DIM));
    auto mindim = args.getInt("MinDim",args.getInt("Minm",1));
    auto doRelCutoff = args.getBool("DoRelCutoff",true);
    auto absoluteCutoff = args.getBool("AbsoluteCutoff",false);
    auto litagset = getTagSet(args,"LeftTags","Link,U");
    auto ritagset = getTagSet(args,"RightTags","Link,V");
    if(litagset == ritagset) Error("In SVD, must specify different tags for the new left and right indices (with Args 'LeftTags' and 'RightTags')");
    if(dim(uI) == 0) throw ResultIsZero("dim(uI) == 0");
    if(dim(vI) == 0) throw ResultIsZero("dim(vI) == 0");

    auto const& is = A.inds();
    // stored blocks are (sector of is[0]) x (sector of is[1]) column-major matrices S; the matrix to factorise is
    // M = S when uI is the first index and M = S^T otherwise (GetBlocks::transpose, decomp.h:443-450)
    const bool transposed = (vI == is.front());
    const auto nb = long(d.offsets.size());
    if(nb == 0) throw ResultIsZero("IQTensor has no blocks");

    std::vector<int64_t> off(nb);
    std::vector<int32_t> mm(nb), nn(nb);
    std::vector<long> su(nb), sv(nb); // sector of uI / vI each block belongs to
    for(auto b : range(nb))
        {
        auto const& io = d.offsets[b];
        off[b] = io.offset;
        mm[b] = int32_t(is[0].blocksize0(io.block[0]));
        nn[b] = int32_t(is[1].blocksize0(io.block[1]));
        su[b] = transposed ? io.block[1] : io.block[0];
        sv[b] = transposed ? io.block[0] : io.block[1];
        }

    // Large blocks go to the device solvers, asynchronously; blocks too small to pay a cuSOLVER launch sequence
    // (min dim < ITB_SVD_DEVICE_MIN_N, default 160) are downloaded and factorised by the reference's own SVD()
    // on the host while the device works. Both produce S = Us diag(s) Vs^H of the STORED matrix.
    static const long dev_min = [] { auto* e = std::getenv("ITB_SVD_DEVICE_MIN_N"); return e ? std::atol(e) : 160l; }();
    // (until the background read of the cuSOLVER/cuBLAS objects has finished, every block stays on the host: a lazy
    // kernel load against a cold page cache costs far more than any of these factorisations)
    const bool dev_ready = itb_solver_ready() != 0;
    std::vector<long> dev_blocks, host_blocks, dev_slot(nb,-1);
    for(auto b : range(nb))
        {
        if(dev_ready && std::min(mm[b],nn[b]) >= dev_min) { dev_slot[b] = long(dev_blocks.size()); dev_blocks.push_back(b); }
        else host_blocks.push_back(b);
        }
    Lap lap;
    svdProf().calls += 1;
    // small blocks come to the host first (these synchronous copies must not queue behind the device solvers) ...
    auto hbuf = std::vector<std::vector<T>>(host_blocks.size());
    for(auto i : range(host_blocks.size()))
        {
        auto b = host_blocks[i];
        hbuf[i].resize(size_t(mm[b])*nn[b]);
        checkSvd(itb_memcpy_d2h(gpu::context(),hbuf[i].data(),static_cast<const char*>(d.buf.data())+size_t(off[b])*sizeof(T),hbuf[i].size()*sizeof(T)),"svd block download");
        }
    // ... then the device batch starts (asynchronous, one host thread per lane) ...
    BatchGuard batch;
    if(!dev_blocks.empty())
        {
        std::vector<int64_t> o; std::vector<int32_t> m2, n2;
        for(auto b : dev_blocks) { o.push_back(off[b]); m2.push_back(mm[b]); n2.push_back(nn[b]); }
        checkSvd(itb_svd_batch_run(gpu::context(),dtypeFor<T>(),int64_t(dev_blocks.size()),o.data(),m2.data(),n2.data(),d.buf.data(),&batch.b),"svd batch");
        }
    lap.mark(0);
    std::vector<long> first(nb+1,0);
    for(auto b : range(nb)) first[b+1] = first[b] + std::min(mm[b],nn[b]);
    auto sval = std::vector<Real>(size_t(first[nb]));
    auto Uh = std::vector<Mat<T>>(nb);
    auto Vh = std::vector<Mat<T>>(nb);
    // ... and the host factorises its share meanwhile (no CUDA calls in this loop)
    for(auto i : range(host_blocks.size()))
        {
        auto b = host_blocks[i];
        auto S = makeMatRef(hbuf[i].data(),hbuf[i].size(),mm[b],nn[b]);
        Vector dv;
        SVD(S,Uh[b],dv,Vh[b],args);
        for(auto j : range(dv.size())) sval[first[b]+j] = dv(j);
        }
    lap.mark(1);
    if(batch.b)
        {
        long ndev = 0;
        for(auto b : dev_blocks) ndev += first[b+1]-first[b];
        auto dvals = std::vector<Real>(size_t(ndev));
        checkSvd(itb_svd_batch_values(batch.b,dvals.data()),"svd values");
        long p = 0;
        for(auto b : dev_blocks)
            for(auto i : range(first[b+1]-first[b])) sval[first[b]+i] = dvals[p++];
        }

    lap.mark(2);
    // density-matrix eigenvalues of all blocks, largest first, handed to the reference's truncate()
    auto all = std::vector<Real>(sval.size());
    for(auto i : range(sval.size())) all[i] = sval[i]*sval[i];
    std::sort(all.begin(),all.end(),std::greater<Real>{});
    auto probs = Vector(std::move(all),VecRange{sval.size()});
    long keep_total = long(probs.size());
    Real truncerr = 0, cut_lo = -1, cut_hi = -1;
    int ndegen = 1;
    if(do_truncate)
        {
        std::tie(truncerr,cut_lo,cut_hi,ndegen) = truncate(probs,maxdim,mindim,cutoff,absoluteCutoff,doRelCutoff,args);
        keep_total = long(probs.size());
        }

    // how many singular values each block keeps: everything above the upper cut, then (up to ndegen) members of
    // the degenerate multiplet sitting between the two cuts, never more than keep_total overall
    auto kept = std::vector<long>(nb,0);
    long taken = 0;
    for(auto b : range(nb))
        {
        auto nsv = first[b+1]-first[b];
        auto const* s = sval.data()+first[b];
        long k = 0;
        if(!do_truncate) k = nsv;
        else
            {
            for(; k < nsv && taken+k < keep_total && s[k]*s[k] > cut_hi; ++k) { }
            for(; ndegen > 0 && k < nsv && taken+k < keep_total && s[k]*s[k] > cut_lo; ++k) --ndegen;
            }
        kept[b] = k;
        taken += k;
        }

    auto Lq = Index::qnstorage{};
    auto Rq = Index::qnstorage{};
    for(auto b : range(nb))
        {
        if(kept[b] == 0) continue;
        Lq.emplace_back(qn(uI,1+su[b]),kept[b]);
        Rq.emplace_back(qn(vI,1+sv[b]),kept[b]);
        }
    if(Lq.empty()) throw ResultIsZero("svd: no singular values kept");
    auto L = Index(std::move(Lq),uI.dir(),litagset);
    auto R = Index(std::move(Rq),vI.dir(),ritagset);
    auto Uis = IndexSet(uI,dag(L));
    auto Dis = IndexSet(L,R);
    auto Vis = IndexSet(vI,dag(R));

    // block structure of U and V exactly as QDense<T>(is,QN()) would allocate it; storage zeroed in HBM
    BlockOffsets Uoff, Voff;
    long Usize = 0, Vsize = 0;
    std::tie(Uoff,Usize) = getBlockOffsets(Uis,QN());
    std::tie(Voff,Vsize) = getBlockOffsets(Vis,QN());
    lap.mark(3);
    auto Ug = QDenseGPU<T>(Uoff,size_t(Usize));
    auto Vg = QDenseGPU<T>(Voff,size_t(Vsize));
    Ug.buf.zero();
    Vg.buf.zero();
    auto Dstore = QDiagReal(Dis);

    long n = 0;
    for(auto b : range(nb))
        {
        if(kept[b] == 0) continue;
        auto k = int32_t(kept[b]);
        auto ublk = Block(2); ublk[0] = su[b]; ublk[1] = n;
        auto vblk = Block(2); vblk[0] = sv[b]; vblk[1] = n;
        auto uo = offsetOf(Ug.offsets,ublk);
        auto vo = offsetOf(Vg.offsets,vblk);
        if(uo < 0 || vo < 0) Error("svd (QDenseGPU): block of U or V missing");
        auto* udst = static_cast<char*>(Ug.buf.data()) + size_t(uo)*sizeof(T);
        auto* vdst = static_cast<char*>(Vg.buf.data()) + size_t(vo)*sizeof(T);
        // S = Us diag(s) Vs^H. The reference stores U = U_M and V = conj(V_M) (svd.cc:211-213) for M = U_M s V_M^H:
        //   M = S   : U_M = Us,       V_M = Vs        -> U block = Us,       V block = conj(Vs)
        //   M = S^T : U_M = conj(Vs), V_M = conj(Us)  -> U block = conj(Vs), V block = Us
        if(dev_slot[b] >= 0)
            {
            auto q = dev_slot[b];
            if(!transposed)
                {
                checkSvd(itb_svd_batch_copy_u(batch.b,q,k,udst),"svd copy U");
                checkSvd(itb_svd_batch_copy_v(batch.b,q,k,vdst,1),"svd copy V");
                }
            else
                {
                checkSvd(itb_svd_batch_copy_v(batch.b,q,k,udst,1),"svd copy U");
                checkSvd(itb_svd_batch_copy_u(batch.b,q,k,vdst),"svd copy V");
                }
            }
        else
            {
            // host-factorised block: upload the kept leading columns (column-major, hence contiguous)
            auto& Us = Uh[b];
            auto& Vs = Vh[b];
            if(isCplx(Vs)) conjugate(Vs);
            auto const& forU = transposed ? Vs : Us;
            auto const& forV = transposed ? Us : Vs;
            checkSvd(itb_memcpy_h2d(gpu::context(),udst,forU.data(),size_t(nrows(forU))*k*sizeof(T)),"svd upload U");
            checkSvd(itb_memcpy_h2d(gpu::context(),vdst,forV.data(),size_t(nrows(forV))*k*sizeof(T)),"svd upload V");
            }
        auto dblk = Block(2); dblk[0] = n; dblk[1] = n;
        auto pD = getBlock(Dstore,Dis,dblk);
        auto const* s = sval.data()+first[b];
        for(auto i : range(k)) pD.data()[i] = std::max(s[i],0.);
        ++n;
        }

    lap.mark(4);
    // D carries the scale of A; U and V are unit-scale (sign convention of svd.cc:392-396)
    Real signfix = (A.scale().sign() == -1) ? -1. : +1.;
    U = ITensor(Uis,std::move(Ug));
    D = ITensor(Dis,std::move(Dstore),A.scale()*signfix);
    V = ITensor(Vis,std::move(Vg),LogNum{signfix});
    if(A.scale().isFiniteReal()) probs *= sqr(A.scale().real0());
    else println("Warning: scale not finite real after svd");
    return Spectrum(std::move(probs),{"Truncerr",truncerr});
    }

// dense (no-QN) branch of svdImpl (svd.cc:88-166) on a DenseGPU tensor: one device SVD, truncate() on the host,
// kept columns copied into new DenseGPU factors
template<typename T>
Spectrum
svdDenseGPU(ITensor const& A, DenseGPU<T> const& d, Index const& uI, Index const& vI,
            ITensor & U, ITensor & D, ITensor & V, Args const& args)
    {
    auto do_truncate = args.getBool("Truncate");
    auto cutoff = args.getReal("Cutoff",MIN_CUT);
    auto maxdim = args.getInt("MaxDim",args.getInt("Maxm",MAX_DIM));
    auto mindim = args.getInt("MinDim",args.getInt("Minm",1));
    auto doRelCutoff = args.getBool("DoRelCutoff",true);
    auto absoluteCutoff = args.getBool("AbsoluteCutoff",false);
    auto litagset = getTagSet(args,"LeftTags","Link,U");
    auto ritagset = getTagSet(args,"RightTags","Link,V");
    if(litagset == ritagset) Error("In SVD, must specify different tags for the new left and right indices (with Args 'LeftTags' and 'RightTags')");

    auto const& is = A.inds();
    const bool transposed = !(uI == is.front()); // stored matrix S is dim(is[0]) x dim(is[1]); M = S or S^T (decomp.cc:70-79)
    int64_t off = 0;
    int32_t mm = int32_t(dim(is[0])), nn = int32_t(dim(is[1]));
    BatchGuard batch;
    checkSvd(itb_svd_batch_run(gpu::context(),dtypeFor<T>(),1,&off,&mm,&nn,d.buf.data(),&batch.b),"svd (dense)");
    auto l = long(std::min(mm,nn));
    auto sval = std::vector<Real>(size_t(l));
    checkSvd(itb_svd_batch_values(batch.b,sval.data()),"svd values");

    auto probs = Vector(l);
    for(auto j : range(l)) probs(j) = sval[j]*sval[j];
    Real truncerr = 0, cut_lo = -1, cut_hi = -1;
    int ndegen = 1;
    long k = l;
    if(do_truncate)
        {
        std::tie(truncerr,cut_lo,cut_hi,ndegen) = truncate(probs,maxdim,mindim,cutoff,absoluteCutoff,doRelCutoff,args);
        k = long(probs.size());
        }
    auto uL = Index(k,litagset);
    auto vL = Index(k,ritagset);
    auto Ug = DenseGPU<T>(size_t(dim(uI))*size_t(k));
    auto Vg = DenseGPU<T>(size_t(dim(vI))*size_t(k));
    // S = Us diag(s) Vs^H; the reference keeps U = U_M and V = conj(V_M) of M = U_M s V_M^H (svd.cc:100-103)
    if(!transposed)
        {
        checkSvd(itb_svd_batch_copy_u(batch.b,0,int32_t(k),Ug.buf.data()),"svd copy U");
        checkSvd(itb_svd_batch_copy_v(batch.b,0,int32_t(k),Vg.buf.data(),1),"svd copy V");
        }
    else
        {
        checkSvd(itb_svd_batch_copy_v(batch.b,0,int32_t(k),Ug.buf.data(),1),"svd copy U");
        checkSvd(itb_svd_batch_copy_u(batch.b,0,int32_t(k),Vg.buf.data()),"svd copy V");
        }
    Real signfix = (A.scale().sign() == -1) ? -1 : +1;
    D = ITensor({uL,vL},Diag<Real>{sval.begin(),sval.begin()+k},A.scale()*signfix);
    U = ITensor({uI,uL},std::move(Ug),LogNum(signfix));
    V = ITensor({vI,vL},std::move(Vg));
    auto DD = Vector(k);
    for(auto j : range(k)) DD(j) = sval[j]*sval[j];
#ifdef USESCALE
    if(A.scale().isFiniteReal()) DD *= sqr(A.scale().real0());
    else println("Warning: scale not finite real after svd");
#endif
    return Spectrum(std::move(DD),{"Truncerr",truncerr});
    }

struct EighGuard
    {
    itb_eigh_batch* b = nullptr;
    ~EighGuard() { if(b) itb_eigh_batch_destroy(b); }
    };

// ITB_PROFILE: wall time of the phases of the device diag_hermitian, printed at exit
struct EighProf
    {
    double secs[4] = {0,0,0,0}; long calls = 0, dev_blocks = 0, host_blocks = 0;
    bool on = std::getenv("ITB_PROFILE") != nullptr;
    ~EighProf()
        {
        if(!on || !calls) return;
        const char* names[4] = {"diagH launch device","diagH host blocks","diagH wait device","diagH truncate+assemble"};
        for(int i = 0; i < 4; ++i) std::fprintf(stderr,"[itensor_b200 profile] %-28s %10ld %12.4f\n",names[i],calls,secs[i]);
        std::fprintf(stderr,"[itensor_b200 profile] %-28s %10ld device blocks, %ld host blocks\n","diagH blocks",dev_blocks,host_blocks);
        }
    };
EighProf& eighProf() { static EighProf p; return p; }

// QN branch of diagHImpl (hermitian.cc:180-426) on a QDenseGPU tensor: every diagonal block is diagonalised on the device
// straight from H's buffer (itb_eigh_batch_run: cuSOLVER syevd / heevd of -block over several streams), only the
// eigenvalues come back, the reference's truncate() decides what is kept, and the kept eigenvectors are copied device to
// device into a new QDenseGPU U. The density matrix never leaves HBM (the reference path downloads it through GetBlocks).
template<typename T>
Spectrum
eighBlocksGPU(ITensor H, QDenseGPU<T> const& d, ITensor & U, ITensor & D, Args const& args)
    {
    // argument defaults exactly as diagHImpl (hermitian.cc:60-92)
    auto origdim = dim(H.inds().front());
    auto cutoff = args.getReal("Cutoff",0.);
    auto maxdim = args.getInt("MaxDim",args.getInt("Maxm",origdim));
    auto mindim = args.getInt("MinDim",args.getInt("Minm",1));
    auto def_do_trunc = args.defined("Cutoff") || args.defined("MaxDim") || args.defined("Maxm");
    auto do_truncate = args.getBool("Truncate",def_do_trunc);
    auto doRelCutoff = args.getBool("DoRelCutoff",true);
    auto absoluteCutoff = args.getBool("AbsoluteCutoff",false);
    auto itagset = getTagSet(args,"Tags","Link");
    if(!do_truncate) maxdim = origdim;

    auto i1 = H.inds().front();
    auto i2 = H.inds().back();
    auto ai = (primeLevel(i1) < primeLevel(i2)) ? i1 : i2;
    auto pdiff = std::abs(primeLevel(i1)-primeLevel(i2));
    // the matrix of a block is M = S (stored, column-major) when ai is the first index and M = S^T = conj(S) otherwise
    // (GetBlocks::transpose, decomp.h:443-450); eigenvectors of conj(S) are the conjugates of those of S
    const bool transposed = !(ai == H.inds().front());
    auto const& is = H.inds();
    const auto nb = long(d.offsets.size());
    if(nb == 0) Error("No blocks in IQTensor svd");

    std::vector<int64_t> off(nb);
    std::vector<int32_t> nn(nb);
    std::vector<long> sa(nb); // sector of ai
    for(auto b : range(nb))
        {
        auto const& io = d.offsets[b];
        auto r = is[0].blocksize0(io.block[0]), c = is[1].blocksize0(io.block[1]);
        if(r != c) Error("diag_hermitian (QDenseGPU): non-square block");
        off[b] = io.offset;
        nn[b] = int32_t(r);
        sa[b] = transposed ? io.block[1] : io.block[0];
        }
    // blocks too small to pay a cuSOLVER launch sequence are diagonalised by the reference's own diagHermitian on the host
    // while the device works on the others
    static const long dev_min = [] { auto* e = std::getenv("ITB_EIGH_DEVICE_MIN_N"); return e ? std::atol(e) : 96l; }();
    const bool dev_ready = itb_solver_ready() != 0;
    std::vector<long> dev_blocks, host_blocks, dev_slot(nb,-1);
    for(auto b : range(nb))
        {
        if(dev_ready && nn[b] >= dev_min) { dev_slot[b] = long(dev_blocks.size()); dev_blocks.push_back(b); }
        else host_blocks.push_back(b);
        }
    Lap lap;
    auto& prof = eighProf();
    auto mark = [&](int slot) { auto n = std::chrono::steady_clock::now(); prof.secs[slot] += std::chrono::duration<double>(n-lap.t).count(); lap.t = n; };
    prof.calls += 1; prof.dev_blocks += long(dev_blocks.size()); prof.host_blocks += long(host_blocks.size());
    auto hbuf = std::vector<std::vector<T>>(host_blocks.size());
    for(auto i : range(host_blocks.size()))
        {
        auto b = host_blocks[i];
        hbuf[i].resize(size_t(nn[b])*nn[b]);
        checkSvd(itb_memcpy_d2h(gpu::context(),hbuf[i].data(),static_cast<const char*>(d.buf.data())+size_t(off[b])*sizeof(T),hbuf[i].size()*sizeof(T)),"eigh block download");
        }
    EighGuard batch;
    // (cuSOLVER's syevd/heevd synchronise their caller inside every call: itb_eigh_batch_run drives each solver lane from its
    // own host thread and returns at once, so this thread diagonalises the small blocks with host LAPACK meanwhile)
    if(!dev_blocks.empty())
        {
        std::vector<int64_t> o; std::vector<int32_t> n2;
        for(auto b : dev_blocks) { o.push_back(off[b]); n2.push_back(nn[b]); }
        checkSvd(itb_eigh_batch_run(gpu::context(),dtypeFor<T>(),int64_t(dev_blocks.size()),o.data(),n2.data(),d.buf.data(),1,&batch.b),"eigh batch");
        }
    mark(0);
    std::vector<long> first(nb+1,0);
    for(auto b : range(nb)) first[b+1] = first[b] + nn[b];
    auto eig = std::vector<Real>(size_t(first[nb])); // per block, largest first
    auto Uh = std::vector<Mat<T>>(nb);
    auto hostBlock = [&](size_t i)
        {
        auto b = host_blocks[i];
        auto S = makeMatRef(hbuf[i].data(),hbuf[i].size(),nn[b],nn[b]);
        Vector dv;
        // the reference's call is diagHermitian(M,UU,d); conjugate(UU) with M = S or transpose(S): done on S here, the
        // conjugations are applied when the columns are placed (below), identically for host and device blocks
        diagHermitian(S,Uh[b],dv);
        for(auto j : range(dv.size())) eig[first[b]+j] = dv(j);
        };
    // (one after the other: handing these small LAPACK calls to several host threads was tried — 2.1 -> 0.3 s in the Hubbard
    // ramp — and dropped: the OpenBLAS this build links is not safe under concurrent callers, dsyev returned info != 0 in one
    // of the GPU-box runs)
    for(auto i : range(host_blocks.size())) hostBlock(i);
    mark(1);
    if(batch.b)
        {
        long ndev = 0;
        for(auto b : dev_blocks) ndev += nn[b];
        auto w = std::vector<Real>(size_t(ndev));
        checkSvd(itb_eigh_batch_values(batch.b,w.data()),"eigh values");
        long p = 0;
        for(auto b : dev_blocks)
            for(auto i : range(nn[b])) eig[first[b]+i] = -w[p++]; // ascending eigenvalues of -block
        }
    mark(2);

    auto alleig = eig;
    std::sort(alleig.begin(),alleig.end(),std::greater<Real>{});
    auto probs = Vector(std::move(alleig),VecRange{eig.size()});
    long m = long(probs.size());
    Real truncerr = 0, docut_lower = -1, docut_upper = -1;
    int ndegen = 0;
    if(do_truncate)
        {
        std::tie(truncerr,docut_lower,docut_upper,ndegen) = truncate(probs,maxdim,mindim,cutoff,absoluteCutoff,doRelCutoff,args);
        m = long(probs.size());
        }
    if(m > maxdim) Error("m > maxdim");

    // how many eigenvectors each block keeps (hermitian.cc:325-370)
    auto kept = std::vector<long>(nb,0);
    long total_m = 0;
    for(auto b : range(nb))
        {
        auto const* e = eig.data()+first[b];
        long this_m = 0;
        if(do_truncate)
            {
            while(this_m < nn[b] && total_m < m && e[this_m] > docut_upper) { ++this_m; ++total_m; }
            while(ndegen > 0 && this_m < nn[b] && total_m < m && e[this_m] > docut_lower) { ++this_m; ++total_m; --ndegen; }
            }
        else { this_m = nn[b]; total_m += this_m; }
        kept[b] = this_m;
        }
    Index::qnstorage iq;
    for(auto b : range(nb)) if(kept[b] > 0) iq.emplace_back(qn(ai,1+sa[b]),kept[b]);
    bool nothing = iq.empty();
    if(nothing) { iq.emplace_back(qn(ai,1+sa[0]),1l); }
    auto dI = Index(std::move(iq),-ai.dir(),itagset);
    auto Uis = IndexSet(dag(ai),dag(dI));
    auto Dis = IndexSet(prime(dI,pdiff),dag(dI));
    BlockOffsets Uoff;
    long Usize = 0;
    std::tie(Uoff,Usize) = getBlockOffsets(Uis,QN());
    auto Ug = QDenseGPU<T>(Uoff,size_t(Usize));
    Ug.buf.zero();
    auto Dstore = QDiagReal(Dis);
    long n = 0;
    for(auto b : range(nb))
        {
        if(kept[b] == 0) continue;
        auto k = int32_t(kept[b]);
        auto ublk = Block(2); ublk[0] = sa[b]; ublk[1] = n;
        auto uo = offsetOf(Ug.offsets,ublk);
        if(uo < 0) Error("diag_hermitian (QDenseGPU): block of U missing");
        auto* udst = static_cast<char*>(Ug.buf.data()) + size_t(uo)*sizeof(T);
        // stored U block = conj(eigenvectors of M): M = S -> conj(W_S); M = S^T = conj(S) -> conj(conj(W_S)) = W_S
        if(dev_slot[b] >= 0) checkSvd(itb_eigh_batch_copy_vectors(batch.b,dev_slot[b],k,udst,transposed ? 0 : 1),"eigh copy U");
        else
            {
            auto& W = Uh[b];
            if(isCplx(W) && !transposed) conjugate(W);
            checkSvd(itb_memcpy_h2d(gpu::context(),udst,W.data(),size_t(nn[b])*k*sizeof(T)),"eigh upload U");
            }
        auto dblk = Block(2); dblk[0] = n; dblk[1] = n;
        auto pD = getBlock(Dstore,Dis,dblk);
        auto const* e = eig.data()+first[b];
        for(auto i : range(k)) pD.data()[i] = e[i];
        ++n;
        }
    if(Uh.size()) gpu::synchronize(); // the host eigenvector matrices must outlive their uploads
    mark(3);
    U = ITensor(Uis,std::move(Ug));
    D = ITensor(Dis,std::move(Dstore),H.scale());
    if(H.scale().isTooBigForReal()) println("scale too big, omitting from reported eigenvalues");
    else probs *= H.scale().real0();
    return Spectrum(std::move(probs),{"Truncerr",truncerr});
    }

struct EighDispatch
    {
    ITensor const& H; ITensor& U; ITensor& D; Args const& args; Spectrum& spec; bool& done;
    void operator()(QDenseGPUReal const& d) { spec = eighBlocksGPU<Real>(H,d,U,D,args); done = true; }
    void operator()(QDenseGPUCplx const& d) { spec = eighBlocksGPU<Cplx>(H,d,U,D,args); done = true; }
    template<typename S> void operator()(S const&) { }
    };

// find out whether A is one of the HBM-resident block-sparse storage types
struct SvdDispatch
    {
    ITensor const& A; Index const& uI; Index const& vI; ITensor& U; ITensor& D; ITensor& V; Args const& args;
    Spectrum& spec; bool& done;
    void operator()(QDenseGPUReal const& d) { spec = svdBlocksGPU<Real>(A,d,uI,vI,U,D,V,args); done = true; }
    void operator()(QDenseGPUCplx const& d) { spec = svdBlocksGPU<Cplx>(A,d,uI,vI,U,D,V,args); done = true; }
    void operator()(DenseGPUReal const& d) { spec = svdDenseGPU<Real>(A,d,uI,vI,U,D,V,args); done = true; }
    void operator()(DenseGPUCplx const& d) { spec = svdDenseGPU<Cplx>(A,d,uI,vI,U,D,V,args); done = true; }
    template<typename S> void operator()(S const&) { }
    };

} // namespace

// ITB_SPECTRUM_LOG=<file>: every svdOrd2 (host or device route) appends one JSON line with its truncation error and kept
// spectrum: the per-scale / per-bond comparison points of the parity runs (TRG: two factorisations per scale)
static void
logSpectrum(Spectrum const& spec)
    {
    static FILE* f = [] { auto* e = std::getenv("ITB_SPECTRUM_LOG"); return e ? std::fopen(e,"w") : nullptr; }();
    if(!f) return;
    std::fprintf(f,"{\"truncerr\": %.17g, \"eigs\": [",spec.truncerr());
    bool first = true;
    for(auto const& e : spec.eigsKept()) { std::fprintf(f,"%s%.17g",first ? "" : ",",e); first = false; }
    std::fprintf(f,"]}\n");
    std::fflush(f);
    }

static Spectrum
svdOrd2Dispatch(ITensor const& A, Index const& uI, Index const& vI, ITensor & U, ITensor & D, ITensor & V, Args args);

Spectrum
svdOrd2(ITensor const& A, Index const& uI, Index const& vI, ITensor & U, ITensor & D, ITensor & V, Args args)
    {
    auto spec = svdOrd2Dispatch(A,uI,vI,U,D,V,args);
    logSpectrum(spec);
    return spec;
    }

static Spectrum
svdOrd2Dispatch(ITensor const& A, Index const& uI, Index const& vI, ITensor & U, ITensor & D, ITensor & V, Args args)
    {
    static const bool device_svd = [] { auto* e = std::getenv("ITB_SVD_DEVICE"); return !(e && std::atoi(e) == 0); }();
    const bool plain = !args.getBool("ComputeQNs",false) && !args.getBool("ShowEigs",false);
    if(device_svd && plain && A.store() && A.order() == 2 && onGPU(A))
        {
        auto a = args;
        if(!a.defined("MaxDim") && a.defined("Maxm")) a.add("MaxDim",a.getInt("Maxm"));
        if(!a.defined("Truncate")) a.add("Truncate",a.defined("Cutoff") || a.defined("MaxDim"));
        Spectrum spec;
        bool done = false;
        applyFunc(SvdDispatch{A,uI,vI,U,D,V,a,spec,done},A.store());
        if(done) return spec;
        }
    // the reference's path (host storage; QDenseGPU reaches its per-block loops through GetBlocks views; dense GPU
    // storage has no raw host view, so it is downloaded)
    if(A.store() && onGPU(A) && !hasQNs(A)) return svdOrd2_host(toCPU(A),uI,vI,U,D,V,args);
    return svdOrd2_host(A,uI,vI,U,D,V,args);
    }

// diag_hermitian (hermitian.cc:429-440) likewise: hermitian.cc is compiled with -Ddiag_hermitian=diag_hermitian_host.
// Block-sparse GPU tensors go through the reference's per-block loop (GetBlocks views; blocks >= 256 are diagonalised
// by cuSOLVER behind the LAPACK boundary). A dense GPU tensor has no raw host view (ToMatRefc is private to
// decomp.cc), so it is diagonalised from a host copy. The eigenvectors return to HBM in both cases.
Spectrum
diag_hermitian_host(ITensor H, ITensor & U, ITensor & D, Args const& args);

Spectrum
diag_hermitian(ITensor H, ITensor & U, ITensor & D, Args const& args)
    {
    static const bool device_eigh = [] { auto* e = std::getenv("ITB_EIGH_DEVICE"); return !(e && std::atoi(e) == 0); }();
    if(device_eigh && H.store() && onGPU(H) && hasQNs(H) && H.order() == 2 && !args.getBool("ComputeQNs",false) && !args.getBool("ShowEigs",false))
        {
        // block-sparse, HBM-resident: batched device eigh, U born in HBM (sign handling as diagHImpl, hermitian.cc:209)
        if(H.scale().sign() < 0) H.scaleTo(H.scale()*(-1));
        Spectrum spec;
        bool done = false;
        applyFunc(EighDispatch{H,U,D,args,spec,done},H.store());
        if(done) return spec;
        }
    if(H.store() && onGPU(H))
        {
        // dense: no raw host view of GPU storage exists, diagonalise a host copy; block-sparse: the reference's loop
        // reads GetBlocks views of a host snapshot. Either way the eigenvectors go (back) to HBM at once, so that the
        // products that follow in denmatDecomp (decomp.h:400-418: cmb * dag(U), U * AAc) find both operands there.
        auto spec = hasQNs(H) ? diag_hermitian_host(H,U,D,args) : diag_hermitian_host(toCPU(H),U,D,args);
        U = toGPU(U);
        return spec;
        }
    return diag_hermitian_host(H,U,D,args);
    }

} //namespace itensor
