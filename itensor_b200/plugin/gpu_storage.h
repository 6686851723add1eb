//
// gpu_storage.h — DenseGPU<T> / QDenseGPU<T>: ITensor storage types whose element data lives in B200 HBM.
//
// Drop-in through the reference's own storage plugin surface: the types are appended to StorageTypes
// (itensor/itdata/storage_types.h:57-74, see INTEGRATION.md for the 3-line patch that
// itensor_b200/plugin/Makefile applies to a build-directory copy) and every operation is a free
// `doTask(Task, Storage...)` overload found by the dispatch in itensor/itdata/dotask.h:283-626.
// All Index / QN / IndexSet bookkeeping stays in the reference's host code; these overloads only
//   (1) turn IndexSet + BlockOffsets into the integer tables of include/itb200.h, and
//   (2) call the extern "C" CUDA layer (libitb200.so). There is no CPU arithmetic in this file:
//       if the CUDA library cannot create a context the first GPU task throws ITError.
//
// Hot-path tasks run on the device:
//   Contract  (QDenseGPU x QDenseGPU, all four real/complex pairings; a host QDense operand is
//              uploaded on the fly)                    replaces itensor/itdata/qdense.cc:671-747
//   Order     (permute, fills in all flux-allowed blocks) replaces qdense.cc:847-893
//   PlusEQ    (permuting accumulate, block-list merge, real->complex promotion) replaces qdense.cc:515-668
//   NormNoScale, Mult<Real|Cplx>, Fill, Conj, MakeCplx, TakeReal/TakeImag, GetElt
//   + integer-only tasks answered from the host-side block list: CalcDiv, NNZBlocks, NNZ, IsEmpty, ...
// Everything else (combiners, Diag, SVD/eigh, element-wise apply/generate/set, I/O) is NOT on the hot
// path (SURVEY §8f "next"): those overloads move the tensor to a host QDense/Dense and hand it to the
// reference's own implementation, so results of svdBond etc. are host tensors that re-enter the device
// the next time they meet a GPU tensor in a contraction.
//
#ifndef ITENSOR_B200_GPU_STORAGE_H
#define ITENSOR_B200_GPU_STORAGE_H

#include <memory>

#include "itensor/itdata/dense.h"
#include "itensor/itdata/qdense.h"
#include "itensor/itdata/task_types.h"

struct itb_ctx;

namespace itensor {

namespace gpu {

// process-wide context on device $ITB_DEVICE (default 0); throws ITError when no device is usable
itb_ctx* context();
void synchronize();
long launchCount();

// Multi-GPU (one process per GPU, ITB_WORLD / ITB_RANK in the environment): world size and rank of this process
int world();
int rank();

// A row-sharded tensor that has not been re-replicated yet: this rank's buffer holds only the rows of one index it
// computed itself (one owner per piece of C, itensor/itdata/qutil.h:285-348, refined to row ranges). The
// all-gather that completes it runs the first time anybody needs the whole buffer (Buffer::data()); a contraction that
// leaves the sharded index uncontracted keeps working on the rows it owns instead (gpu_storage.cc, contractQ), which is
// how the four steps of LocalOp::product (localop.h:346-362) run without communication until H*phi is complete.
struct Pending;

// RAII device buffer from the context's caching pool; copies are deep (device-to-device)
class Buffer
    {
    void* p_ = nullptr;
    size_t bytes_ = 0;
    mutable std::shared_ptr<Pending> pending_; // set: only this rank's rows are valid until data() completes the buffer
    public:
    Buffer() { }
    explicit Buffer(size_t bytes);
    Buffer(Buffer const& o);
    Buffer(Buffer&& o) noexcept : p_(o.p_), bytes_(o.bytes_), pending_(std::move(o.pending_)) { o.p_ = nullptr; o.bytes_ = 0; }
    Buffer& operator=(Buffer const& o);
    Buffer& operator=(Buffer&& o) noexcept;
    ~Buffer();
    void* data() const;                       // the complete buffer (runs a pending all-gather first)
    void* rawData() const { return p_; }      // the buffer as it is (only own rows valid while pending() is set)
    std::shared_ptr<Pending> const& pending() const { return pending_; }
    void setPending(std::shared_ptr<Pending> p) const { pending_ = std::move(p); }
    size_t bytes() const { return bytes_; }
    void upload(void const* host, size_t bytes);
    void download(void* host, size_t bytes) const;
    void zero();
    };

} //namespace gpu

template<typename T>
class QDenseGPU
    {
    public:
    using value_type = T;

    BlockOffsets offsets; // same host bookkeeping as QDense<T>::offsets (sorted block -> element offset)
    gpu::Buffer buf;      // flat element data in HBM, laid out exactly like QDense<T>::store
    size_t n = 0;         // number of stored elements
    // host snapshot handed out as raw matrix views by GetBlocks (per-block eigh/SVD, decomp.cc:84-113);
    // refreshed on every GetBlocks call, never used as a compute fallback
    mutable std::shared_ptr<std::vector<T>> mirror;

    QDenseGPU() { }

    // uninitialised device storage with the given block structure
    QDenseGPU(BlockOffsets const& off, size_t size) : offsets(off), buf(size*sizeof(T)), n(size) { }

    // upload
    explicit QDenseGPU(QDense<T> const& h);

    // download
    QDense<T> toHost() const;

    size_t size() const { return n; }
    explicit operator bool() const { return n != 0 && !offsets.empty(); }
    };

template<typename T>
class DenseGPU
    {
    public:
    using value_type = T;

    gpu::Buffer buf;
    size_t n = 0;

    DenseGPU() { }
    explicit DenseGPU(size_t size) : buf(size*sizeof(T)), n(size) { }
    explicit DenseGPU(Dense<T> const& h);
    Dense<T> toHost() const;
    size_t size() const { return n; }
    explicit operator bool() const { return n != 0; }
    };

using QDenseGPUReal = QDenseGPU<Real>;
using QDenseGPUCplx = QDenseGPU<Cplx>;
using DenseGPUReal = DenseGPU<Real>;
using DenseGPUCplx = DenseGPU<Cplx>;

const char* typeNameOf(QDenseGPUReal const&);
const char* typeNameOf(QDenseGPUCplx const&);
const char* typeNameOf(DenseGPUReal const&);
const char* typeNameOf(DenseGPUCplx const&);

template<typename T> bool constexpr isReal(QDenseGPU<T> const&) { return std::is_same<T,Real>::value; }
template<typename T> bool constexpr isCplx(QDenseGPU<T> const&) { return std::is_same<T,Cplx>::value; }
template<typename T> bool constexpr isReal(DenseGPU<T> const&) { return std::is_same<T,Real>::value; }
template<typename T> bool constexpr isCplx(DenseGPU<T> const&) { return std::is_same<T,Cplx>::value; }

//
// ---------------------------------------------------------------------------------------------
//  QDenseGPU tasks
// ---------------------------------------------------------------------------------------------
//

// integer-only tasks: answered from the host block list, no device access
template<typename T> QN doTask(CalcDiv const& C, QDenseGPU<T> const& d);
template<typename T> long doTask(NNZBlocks, QDenseGPU<T> const& d) { return long(d.offsets.size()); }
template<typename T> long doTask(NNZ, QDenseGPU<T> const& d) { return long(d.n); }
template<typename T> bool doTask(IsEmpty, QDenseGPU<T> const& d) { return d.offsets.empty(); }
template<typename T> bool constexpr doTask(CheckComplex, QDenseGPU<T> const& d) { return isCplx(d); }
auto constexpr inline doTask(StorageType const&, QDenseGPUReal const&) ->StorageType::Type { return StorageType::QDenseReal; }
auto constexpr inline doTask(StorageType const&, QDenseGPUCplx const&) ->StorageType::Type { return StorageType::QDenseCplx; }

// device tasks
template<typename T> Real doTask(NormNoScale, QDenseGPU<T> const& d);
template<typename T> void doTask(Mult<Real> const& M, QDenseGPU<T>& d);
void doTask(Mult<Cplx> const& M, QDenseGPUReal const& d, ManageStore& m);
void doTask(Mult<Cplx> const& M, QDenseGPUCplx& d);
void doTask(MakeCplx const&, QDenseGPUCplx& d);
void doTask(MakeCplx const&, QDenseGPUReal const& d, ManageStore& m);
void doTask(Fill<Real> const& F, QDenseGPUReal& d);
void doTask(Fill<Cplx> const& F, QDenseGPUCplx& d);
void doTask(Fill<Real> const& F, QDenseGPUCplx const& d, ManageStore& m);
void doTask(Fill<Cplx> const& F, QDenseGPUReal const& d, ManageStore& m);
void inline doTask(Conj, QDenseGPUReal const&) { }
void doTask(Conj, QDenseGPUCplx& d);
void inline doTask(TakeReal, QDenseGPUReal const&) { }
void doTask(TakeReal, QDenseGPUCplx const& d, ManageStore& m);
void doTask(TakeImag, QDenseGPUReal& d);
void doTask(TakeImag, QDenseGPUCplx const& d, ManageStore& m);
Cplx doTask(GetElt& G, QDenseGPUReal const& d);
Cplx doTask(GetElt& G, QDenseGPUCplx const& d);
template<typename T> void doTask(Order const& O, QDenseGPU<T>& d);

template<typename VA, typename VB>
void doTask(Contract& Con, QDenseGPU<VA> const& A, QDenseGPU<VB> const& B, ManageStore& m);
template<typename VA, typename VB>
void doTask(Contract& Con, QDenseGPU<VA> const& A, QDense<VB> const& B, ManageStore& m);
template<typename VA, typename VB>
void doTask(Contract& Con, QDense<VA> const& A, QDenseGPU<VB> const& B, ManageStore& m);

template<typename TA, typename TB>
void doTask(PlusEQ const& P, QDenseGPU<TA> const& A, QDenseGPU<TB> const& B, ManageStore& m);
template<typename TA, typename TB>
void doTask(PlusEQ const& P, QDenseGPU<TA> const& A, QDense<TB> const& B, ManageStore& m);
template<typename TA, typename TB>
void doTask(PlusEQ const& P, QDense<TA> const& A, QDenseGPU<TB> const& B, ManageStore& m);

//
// Off-hot-path tasks: hand a host copy to the reference's own QDense implementation.
//
template<typename T> Cplx doTask(SumEls S, QDenseGPU<T> const& d) { return doTask(S,d.toHost()); }
template<typename T> void doTask(PrintIT& P, QDenseGPU<T> const& d) { doTask(P,d.toHost()); }
template<typename F, typename T>
void doTask(VisitIT<F>& V, QDenseGPU<T> const& d) { doTask(V,d.toHost()); }
template<typename F, typename T>
void doTask(ApplyIT<F>& A, QDenseGPU<T>& d) { auto h = d.toHost(); doTask(A,h); d = QDenseGPU<T>(h); }
template<typename T>
void doTask(SetElt<Real>& S, QDenseGPU<T>& d) { auto h = d.toHost(); doTask(S,h); d = QDenseGPU<T>(h); }
void inline doTask(SetElt<Cplx>& S, QDenseGPUCplx& d) { auto h = d.toHost(); doTask(S,h); d = QDenseGPUCplx(h); }
template<typename F>
void doTask(GenerateIT<F,Real>& G, QDenseGPUReal& d) { auto h = d.toHost(); doTask(G,h); d = QDenseGPUReal(h); }
template<typename F>
void doTask(GenerateIT<F,Cplx>& G, QDenseGPUCplx& d) { auto h = d.toHost(); doTask(G,h); d = QDenseGPUCplx(h); }
// results below are host storage: they come back to the device when they next meet a GPU tensor
template<typename V>
void doTask(RemoveQNs& R, QDenseGPU<V> const& d, ManageStore& m) { doTask(R,d.toHost(),m); }
template<typename T> bool doTask(IsDense, QDenseGPU<T> const&) { return true; }

// QN combiner (index fusion) on the device: block scatter/gather through the strided block-copy kernel
// (replaces combine()/uncombine(), itensor/itdata/qcombiner.cc:125-301); the result stays in HBM.
template<typename T> void doTask(Contract& C, QDenseGPU<T> const& d, QCombiner const& cmb, ManageStore& m);
template<typename T> void doTask(Contract& C, QCombiner const& cmb, QDenseGPU<T> const& d, ManageStore& m);

// per-block matrix views for the host-side eigh/SVD loops (decomp.cc:84-113): views into a host snapshot
template<typename T> struct GetBlocks;
template<typename T> struct Ord2Block;
template<typename T> std::vector<Ord2Block<T>> doTask(GetBlocks<T> const& G, QDenseGPU<T> const& d);

// QDiag x QDense (qdiag.cc:352-566): an order-2 diagonal tensor with a contracted index (svdBond's A *= D,
// mps_impl.h:63-64) is written out as block-diagonal QDense matrices and contracted on the device, so the site tensor
// stays in HBM; any other diagonal product runs the reference's host code on a host copy.
template<typename TA, typename TB> void doTask(Contract& C, QDenseGPU<TA> const& d, QDiag<TB> const& t, ManageStore& m);
template<typename TA, typename TB> void doTask(Contract& C, QDiag<TA> const& t, QDenseGPU<TB> const& d, ManageStore& m);
template<typename VA, typename VB>
void doTask(NCProd& P, QDenseGPU<VA> const& A, QDenseGPU<VB> const& B, ManageStore& m) { doTask(P,A.toHost(),B.toHost(),m); }

//
// ---------------------------------------------------------------------------------------------
//  DenseGPU tasks (one block, no QNs): same kernels through the same C ABI
// ---------------------------------------------------------------------------------------------
//
template<typename T> bool constexpr doTask(CheckComplex, DenseGPU<T> const& d) { return isCplx(d); }
auto constexpr inline doTask(StorageType const&, DenseGPUReal const&) ->StorageType::Type { return StorageType::DenseReal; }
auto constexpr inline doTask(StorageType const&, DenseGPUCplx const&) ->StorageType::Type { return StorageType::DenseCplx; }
template<typename T> Real doTask(NormNoScale, DenseGPU<T> const& d);
template<typename T> void doTask(Mult<Real> const& M, DenseGPU<T>& d);
void doTask(Mult<Cplx> const& M, DenseGPUReal const& d, ManageStore& m);
void doTask(Mult<Cplx> const& M, DenseGPUCplx& d);
void doTask(MakeCplx const&, DenseGPUCplx& d);
void doTask(MakeCplx const&, DenseGPUReal const& d, ManageStore& m);
void doTask(Fill<Real> const& F, DenseGPUReal& d);
void doTask(Fill<Cplx> const& F, DenseGPUCplx& d);
void doTask(Fill<Real> const& F, DenseGPUCplx const& d, ManageStore& m);
void doTask(Fill<Cplx> const& F, DenseGPUReal const& d, ManageStore& m);
void inline doTask(Conj, DenseGPUReal const&) { }
void doTask(Conj, DenseGPUCplx& d);
void inline doTask(TakeReal, DenseGPUReal const&) { }
void doTask(TakeReal, DenseGPUCplx const& d, ManageStore& m);
void doTask(TakeImag, DenseGPUReal& d);
void doTask(TakeImag, DenseGPUCplx const& d, ManageStore& m);
Cplx doTask(GetElt const& G, DenseGPUReal const& d);
Cplx doTask(GetElt const& G, DenseGPUCplx const& d);
template<typename T> void doTask(Order const& O, DenseGPU<T>& d);
template<typename VA, typename VB>
void doTask(Contract& Con, DenseGPU<VA> const& A, DenseGPU<VB> const& B, ManageStore& m);
template<typename VA, typename VB>
void doTask(Contract& Con, DenseGPU<VA> const& A, Dense<VB> const& B, ManageStore& m);
template<typename VA, typename VB>
void doTask(Contract& Con, Dense<VA> const& A, DenseGPU<VB> const& B, ManageStore& m);
template<typename TA, typename TB>
void doTask(PlusEQ const& P, DenseGPU<TA> const& A, DenseGPU<TB> const& B, ManageStore& m);
template<typename TA, typename TB>
void doTask(PlusEQ const& P, DenseGPU<TA> const& A, Dense<TB> const& B, ManageStore& m);
template<typename TA, typename TB>
void doTask(PlusEQ const& P, Dense<TA> const& A, DenseGPU<TB> const& B, ManageStore& m);

template<typename T> Cplx doTask(SumEls S, DenseGPU<T> const& d) { return doTask(S,d.toHost()); }
template<typename T> void doTask(PrintIT& P, DenseGPU<T> const& d) { doTask(P,d.toHost()); }
template<typename F, typename T>
void doTask(VisitIT<F>& V, DenseGPU<T> const& d) { doTask(V,d.toHost()); }
// element-wise apply / set (dense.h:166-182, dense.cc:48-75) on a host copy; the result is NEW device storage made through
// the ManageStore. (The reference overloads must not be handed the outer ManageStore: their m.modifyData(d) casts the
// managed object to ITWrap<Dense<T>>, which a DenseGPU<T> is not.)
template<typename F, typename T>
void
doTask(ApplyIT<F>& A, DenseGPU<T> const& d, ManageStore& m)
    {
    using new_type = ApplyIT_result_of<T,F>;
    auto h = d.toHost();
    if(switchesType<T>(A))
        {
        auto nh = Dense<new_type>(h.size());
        for(auto i : range(h)) A(h.store[i],nh.store[i]);
        m.makeNewData<DenseGPU<new_type>>(nh);
        }
    else
        {
        for(auto& el : h) A(el);
        m.makeNewData<DenseGPU<T>>(h);
        }
    }
template<typename E, typename T>
void
doTask(SetElt<E> const& S, DenseGPU<T> const& d, ManageStore& m)
    {
    auto h = d.toHost();
    if constexpr (std::is_same<E,Cplx>::value && std::is_same<T,Real>::value)
        {
        auto nh = DenseCplx(h.begin(),h.end());
        nh[offset(S.is,S.inds)] = S.elt;
        m.makeNewData<DenseGPUCplx>(nh);
        }
    else
        {
        h[offset(S.is,S.inds)] = S.elt;
        m.makeNewData<DenseGPU<T>>(h);
        }
    }
// Dense combiner on the device (combiner.cc:55-178): combining is a relabelling when the fused indices already sit
// together in combiner order, otherwise one device permute brings them to the front; uncombining is always a
// relabelling. The result stays in HBM; the decompositions that follow reach it through svdOrd2 / diag_hermitian
// (plugin/svd_gpu.cc), not through raw host views.
template<typename T> void doTask(Contract& C, DenseGPU<T> const& d, Combiner const& cmb, ManageStore& m);
template<typename T> void doTask(Contract& C, Combiner const& cmb, DenseGPU<T> const& d, ManageStore& m);
// Diag x Dense (diag.cc:128-207): a delta that replaces one index is metadata only (the storage is shared / copied
// on the device); everything else (scaling by a diagonal, traces) goes through the reference's host code.
template<typename TA, typename TB> void doTask(Contract& C, DenseGPU<TA> const& d, Diag<TB> const& t, ManageStore& m);
template<typename TA, typename TB> void doTask(Contract& C, Diag<TA> const& t, DenseGPU<TB> const& d, ManageStore& m);

// binary I/O (ITensor::write, LocalMPO disk spill): GPU storage is written in the host wire format
// (StorageType QDenseReal/... above), so files read back as ordinary host tensors.
template<typename T> void write(std::ostream& s, QDenseGPU<T> const& d) { write(s,d.toHost()); }
template<typename T> void write(std::ostream& s, DenseGPU<T> const& d) { write(s,d.toHost()); }

} //namespace itensor

#endif
