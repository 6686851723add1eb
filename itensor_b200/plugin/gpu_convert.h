//
// gpu_convert.h — move ITensors / MPS / MPO between host storage (Dense, QDense) and HBM-resident
// storage (DenseGPU, QDenseGPU). Uses applyFunc (itensor/itdata/applyfunc.h:16-98) to reach the typed
// storage without touching reference files. Other storage kinds (Diag, Combiner, Scalar) are left alone.
//
#ifndef ITENSOR_B200_GPU_CONVERT_H
#define ITENSOR_B200_GPU_CONVERT_H

#include "itensor/itensor.h"

namespace itensor {

namespace detail {
struct ToGPU
    {
    IndexSet const& is;
    ITensor& out;
    void operator()(QDenseReal const& d) { out = ITensor(is,QDenseGPUReal(d)); }
    void operator()(QDenseCplx const& d) { out = ITensor(is,QDenseGPUCplx(d)); }
    void operator()(DenseReal const& d) { out = ITensor(is,DenseGPUReal(d)); }
    void operator()(DenseCplx const& d) { out = ITensor(is,DenseGPUCplx(d)); }
    template<typename S> void operator()(S const&) { }
    };
struct ToCPU
    {
    IndexSet const& is;
    ITensor& out;
    void operator()(QDenseGPUReal const& d) { out = ITensor(is,d.toHost()); }
    void operator()(QDenseGPUCplx const& d) { out = ITensor(is,d.toHost()); }
    void operator()(DenseGPUReal const& d) { out = ITensor(is,d.toHost()); }
    void operator()(DenseGPUCplx const& d) { out = ITensor(is,d.toHost()); }
    template<typename S> void operator()(S const&) { }
    };
struct IsGPU
    {
    bool& yes;
    template<typename T> void operator()(QDenseGPU<T> const&) { yes = true; }
    template<typename T> void operator()(DenseGPU<T> const&) { yes = true; }
    template<typename S> void operator()(S const&) { }
    };
} //namespace detail

ITensor inline
toGPU(ITensor const& T)
    {
    if(!T.store()) return T;
    ITensor out = T;
    applyFunc(detail::ToGPU{T.inds(),out},T.store());
    return out;
    }

ITensor inline
toCPU(ITensor const& T)
    {
    if(!T.store()) return T;
    ITensor out = T;
    applyFunc(detail::ToCPU{T.inds(),out},T.store());
    return out;
    }

// (non-const lvalues would otherwise pick the MPS/MPO template below)
ITensor inline toGPU(ITensor & T) { return toGPU(static_cast<ITensor const&>(T)); }
ITensor inline toCPU(ITensor & T) { return toCPU(static_cast<ITensor const&>(T)); }

bool inline
onGPU(ITensor const& T)
    {
    bool yes = false;
    if(T.store()) applyFunc(detail::IsGPU{yes},T.store());
    return yes;
    }

// MPS / MPO: anything with length() and ref(j)
template<class MPSLike>
void
toGPU(MPSLike & psi)
    {
    for(auto j : range1(length(psi))) psi.ref(j) = toGPU(psi(j));
    }
// Re-pin one site tensor in HBM without disturbing the orthogonality bookkeeping (ref() widens the limits).
// svdBond's factors come out of the reference's host-side per-block loops as host storage; a DMRG observer
// calls this for sites b, b+1 after every bond update (dmrg.h:445 obs.measure) so that the next
// phi = A(b)*A(b+1), H_eff*phi and the Davidson vector algebra all run on device-resident data.
template<class MPSLike>
void
pinToGPU(MPSLike & psi, int j)
    {
    if(j < 1 || j > length(psi) || onGPU(psi(j))) return;
    auto l = psi.leftLim();
    auto r = psi.rightLim();
    psi.ref(j) = toGPU(psi(j));
    psi.leftLim(l);
    psi.rightLim(r);
    }

template<class MPSLike>
void
toCPU(MPSLike & psi)
    {
    for(auto j : range1(length(psi))) psi.ref(j) = toCPU(psi(j));
    }

} //namespace itensor

#endif
