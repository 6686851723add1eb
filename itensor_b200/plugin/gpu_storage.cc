//
// gpu_storage.cc — doTask overloads of DenseGPU<T> / QDenseGPU<T> (see gpu_storage.h).
//
// Host side only: label matching and result IndexSets come from the reference's own helpers
// (computeLabels, contractIS, calcDiv, getBlockOffsets), the block tables and all arithmetic from the
// C ABI in include/itb200.h. Citations name the reference routine each overload stands in for.
//
#include "itensor/itensor.h"
#include "itensor/decomp.h"
#include "itensor/itdata/qcombiner.h"
#include "itensor/itdata/qutil.h"
#include "itensor/tensor/contract.h"
#include "itensor/tensor/slicemat.h"

#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <unordered_map>

#include "itb200.h"

#include <chrono>
#include <fstream>
#include <map>
#include <optional>
#include <thread>

namespace itensor {

namespace gpu {

static itb_ctx* g_ctx = nullptr;

// ITB_PROFILE=1 (2: with device synchronisation): host wall time and call counts per plugin entry point, printed at exit (the storage-level
// counterpart of the reference's -DCOLLECT_TIMES section timers, itensor/util/timers.h)
struct Prof
    {
    struct Rec { double secs = 0; long calls = 0; };
    std::map<std::string,Rec> recs;
    bool on = false;
    bool sync = false; // ITB_PROFILE=2: synchronise the device around every entry so that GPU time lands in the entry that queued it
    Prof() { if(auto* e = std::getenv("ITB_PROFILE")) { on = std::atoi(e) != 0; sync = std::atoi(e) >= 2; } }
    ~Prof()
        {
        if(!on) return;
        std::fprintf(stderr,"[itensor_b200 profile] %-28s %10s %12s\n","entry","calls","seconds");
        for(auto& kv : recs) std::fprintf(stderr,"[itensor_b200 profile] %-28s %10ld %12.4f\n",kv.first.c_str(),kv.second.calls,kv.second.secs);
        }
    };
static Prof& prof() { static Prof p; return p; }
struct Scope
    {
    const char* name;
    std::chrono::steady_clock::time_point t0;
    bool on;
    explicit Scope(const char* n) : name(n), on(prof().on)
        {
        if(!on) return;
        if(prof().sync && g_ctx) itb_synchronize(g_ctx);
        t0 = std::chrono::steady_clock::now();
        }
    ~Scope()
        {
        if(!on) return;
        if(prof().sync && g_ctx) itb_synchronize(g_ctx);
        auto& r = prof().recs[name];
        r.secs += std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
        r.calls += 1;
        }
    };
#define ITB_SCOPE(name) gpu::Scope itb_scope_(name)

static void
check(int rc, const char* what)
    {
    if(rc != ITB_OK) throw ITError(tinyformat::format("itensor_b200 (%s): %s",what,itb_last_error()));
    }

itb_ctx*
context()
    {
    if(!g_ctx)
        {
        int dev = 0;
        if(auto* e = std::getenv("ITB_DEVICE")) dev = std::atoi(e);
        // One process per GPU: every rank must take the SAME numerical route for every block, or their states drift apart
        // by rounding, then their control flow (Davidson iterations, kept dimensions), and the next collective never
        // matches. The library warm-up makes "is the device solver ready yet" a matter of timing, so it is switched off
        // here: the device solvers are used from the first call on every rank (first-use stalls instead).
        if(std::getenv("ITB_WORLD") && std::atoi(std::getenv("ITB_WORLD")) > 1) setenv("ITB_WARM_LIBS","0",1);
        check(itb_ctx_create(dev,&g_ctx),"no CUDA device: DenseGPU/QDenseGPU have no CPU fallback");
        }
    return g_ctx;
    }

void
synchronize() { check(itb_synchronize(context()),"synchronize"); }

long
launchCount() { return g_ctx ? long(itb_launch_count(g_ctx)) : 0; }

//
// ---- multi-GPU: one process per GPU ---------------------------------------------------------------------------------
//
static int
envInt(const char* name, int dflt) { auto* e = std::getenv(name); return e ? std::atoi(e) : dflt; }
int
world() { static int w = std::max(1,envInt("ITB_WORLD",1)); return w; }
int
rank() { static int r = envInt("ITB_RANK",0); return r; }

// communicator: rank 0 creates the NCCL id and publishes it in the file $ITB_COMM_FILE (atomically, by rename), the
// other ranks poll for it; itb_comm_create is collective
static itb_comm*
comm()
    {
    static itb_comm* c = nullptr;
    if(c) return c;
    auto* path = std::getenv("ITB_COMM_FILE");
    if(!path) Error("itensor_b200: ITB_WORLD > 1 needs ITB_COMM_FILE (a path all ranks can see) for the communicator id");
    uint8_t id[ITB_COMM_ID_BYTES];
    auto file = std::string(path);
    if(rank() == 0)
        {
        check(itb_comm_unique_id(id),"comm id");
        auto tmp = file+".tmp";
        { std::ofstream f(tmp,std::ios::binary); f.write((const char*)id,sizeof(id)); }
        std::rename(tmp.c_str(),file.c_str());
        }
    else
        {
        for(int tries = 0;; ++tries)
            {
            std::ifstream f(file,std::ios::binary);
            if(f && f.read((char*)id,sizeof(id))) break;
            if(tries > 60000) Error("itensor_b200: timed out waiting for the communicator id file");
            std::this_thread::sleep_for(std::chrono::milliseconds(5));
            }
        }
    check(itb_comm_create(context(),world(),rank(),id,&c),"comm create");
    return c;
    }

// pack / scatter plans are shared between the structure cache and every pending tensor that still needs them (an
// environment tensor can stay row-sharded across many bonds, long after the cache has recycled its entry)
struct CopyPlan
    {
    itb_permute_plan* p = nullptr;
    explicit CopyPlan(itb_permute_plan* p_) : p(p_) { }
    CopyPlan(CopyPlan const&) = delete;
    CopyPlan& operator=(CopyPlan const&) = delete;
    ~CopyPlan() { if(p) itb_permute_plan_destroy(p); }
    };

struct Pending
    {
    std::shared_ptr<CopyPlan> pack;     // own rows of the tensor -> this rank's segment
    std::shared_ptr<CopyPlan> unpack;   // the other ranks' segments -> their places in the tensor
    long segDoubles = 0;                // padded segment length in doubles
    // what a follow-up contraction needs to keep working on the same rows
    Index shardIndex;
    int nsect = 0;
    std::vector<int64_t> lo, hi;        // [world*nsect]: rows of sector s owned by rank r = [lo[r*nsect+s], hi[r*nsect+s])
    };

static void
runPending(Pending const& P, void* p)
    {
    ITB_SCOPE("all-gather rows");
    auto* ctx = context();
    void* seg = nullptr;
    check(itb_malloc(ctx,size_t(P.segDoubles)*world()*sizeof(double),&seg),"malloc (gather segments)");
    auto* mine = static_cast<double*>(seg)+size_t(P.segDoubles)*rank();
    check(itb_permute_run(ctx,P.pack->p,p,mine,1.,0.,0),"pack rows");
    check(itb_comm_allgather(comm(),ctx,mine,seg,P.segDoubles),"all-gather");
    check(itb_permute_run(ctx,P.unpack->p,seg,p,1.,0.,0),"scatter rows");
    check(itb_free(ctx,seg),"free (gather segments)"); // stream-ordered pool: reused only behind the scatter
    }

Buffer::
Buffer(size_t bytes) : bytes_(bytes)
    {
    check(itb_malloc(context(),bytes ? bytes : 1,&p_),"malloc");
    }

Buffer::
Buffer(Buffer const& o) : bytes_(o.bytes_)
    {
    if(!o.p_) return;
    check(itb_malloc(context(),bytes_ ? bytes_ : 1,&p_),"malloc");
    check(itb_memcpy_d2d(context(),p_,o.data(),bytes_),"d2d"); // (o.data(): a row-sharded source is completed first)
    }

void* Buffer::
data() const
    {
    if(pending_)
        {
        auto p = std::move(pending_);
        pending_.reset();
        runPending(*p,p_);
        }
    return p_;
    }

Buffer& Buffer::
operator=(Buffer const& o)
    {
    if(this == &o) return *this;
    Buffer tmp(o);
    *this = std::move(tmp);
    return *this;
    }

Buffer& Buffer::
operator=(Buffer&& o) noexcept
    {
    if(this == &o) return *this;
    if(p_ && g_ctx) itb_free(g_ctx,p_);
    p_ = o.p_; bytes_ = o.bytes_; pending_ = std::move(o.pending_);
    o.p_ = nullptr; o.bytes_ = 0;
    return *this;
    }

Buffer::
~Buffer()
    {
    if(p_ && g_ctx) itb_free(g_ctx,p_);
    }

void Buffer::
upload(void const* host, size_t bytes) { if(bytes == bytes_) pending_.reset(); check(itb_memcpy_h2d(context(),data(),host,bytes),"h2d"); }

void Buffer::
download(void* host, size_t bytes) const { check(itb_memcpy_d2h(context(),host,data(),bytes),"d2h"); }

void Buffer::
zero() { pending_.reset(); check(itb_memset0(context(),p_,bytes_),"memset"); }

} //namespace gpu

using gpu::check;
using gpu::context;

const char* typeNameOf(QDenseGPUReal const&) { return "QDenseGPUReal"; }
const char* typeNameOf(QDenseGPUCplx const&) { return "QDenseGPUCplx"; }
const char* typeNameOf(DenseGPUReal const&) { return "DenseGPUReal"; }
const char* typeNameOf(DenseGPUCplx const&) { return "DenseGPUCplx"; }

template<typename T> int constexpr dtypeOf() { return std::is_same<T,Cplx>::value ? ITB_C64 : ITB_F64; }

//
// host <-> device
//
template<typename T>
QDenseGPU<T>::
QDenseGPU(QDense<T> const& h) : offsets(h.offsets), buf(h.store.size()*sizeof(T)), n(h.store.size())
    {
    ITB_SCOPE("upload QDense");
    // upload synchronously: the host vector may die right after this constructor
    buf.upload(h.store.data(),n*sizeof(T));
    gpu::synchronize();
    }
template QDenseGPU<Real>::QDenseGPU(QDense<Real> const&);
template QDenseGPU<Cplx>::QDenseGPU(QDense<Cplx> const&);

template<typename T>
QDense<T> QDenseGPU<T>::
toHost() const
    {
    ITB_SCOPE("download QDense");
    auto h = QDense<T>(undef,offsets,n);
    buf.download(h.store.data(),n*sizeof(T));
    return h;
    }
template QDense<Real> QDenseGPU<Real>::toHost() const;
template QDense<Cplx> QDenseGPU<Cplx>::toHost() const;

template<typename T>
DenseGPU<T>::
DenseGPU(Dense<T> const& h) : buf(h.store.size()*sizeof(T)), n(h.store.size())
    {
    buf.upload(h.store.data(),n*sizeof(T));
    gpu::synchronize();
    }
template DenseGPU<Real>::DenseGPU(Dense<Real> const&);
template DenseGPU<Cplx>::DenseGPU(Dense<Cplx> const&);

template<typename T>
Dense<T> DenseGPU<T>::
toHost() const
    {
    auto h = Dense<T>(undef,n);
    buf.download(h.store.data(),n*sizeof(T));
    return h;
    }
template Dense<Real> DenseGPU<Real>::toHost() const;
template Dense<Cplx> DenseGPU<Cplx>::toHost() const;

//
// IndexSet + BlockOffsets -> itb_tensor_desc (the only thing that crosses the C boundary besides pointers)
//
struct Desc
    {
    std::vector<int32_t> nsect, blocks;
    std::vector<int64_t> sect, offsets;
    itb_tensor_desc d;

    // block-sparse: sectors from Index::nblock()/blocksize0()
    Desc(IndexSet const& is, BlockOffsets const& off, size_t nelems, int dtype)
        {
        auto r = order(is);
        for(auto j : range(r))
            {
            nsect.push_back(is[j].nblock());
            for(auto b : range(is[j].nblock())) sect.push_back(is[j].blocksize0(b));
            }
        for(auto const& bo : off)
            {
            for(auto j : range(r)) blocks.push_back(int32_t(bo.block[j]));
            offsets.push_back(bo.offset);
            }
        finish(r,off.size(),nelems,dtype);
        }

    // dense: one sector per index, one block
    Desc(IndexSet const& is, size_t nelems, int dtype)
        {
        auto r = order(is);
        for(auto j : range(r))
            {
            nsect.push_back(1);
            sect.push_back(dim(is[j]));
            blocks.push_back(0);
            }
        offsets.push_back(0);
        finish(r,1,nelems,dtype);
        }

    void
    finish(long r, size_t nblocks, size_t nelems, int dtype)
        {
        if(nsect.empty()) nsect.push_back(0);
        if(sect.empty()) sect.push_back(0);
        if(blocks.empty()) blocks.push_back(0);
        if(offsets.empty()) offsets.push_back(0);
        d.order = int32_t(r);
        d.dtype = dtype;
        d.nsect = nsect.data();
        d.sect = sect.data();
        d.nblocks = int64_t(nblocks);
        d.blocks = blocks.data();
        d.offsets = offsets.data();
        d.nelems = int64_t(nelems);
        }

    void
    appendKey(std::string & k) const
        {
        k.append((const char*)&d.order,sizeof(int32_t)*2);
        k.append((const char*)&d.nblocks,sizeof(int64_t));
        k.append((const char*)nsect.data(),nsect.size()*sizeof(int32_t));
        k.append((const char*)sect.data(),sect.size()*sizeof(int64_t));
        k.append((const char*)blocks.data(),blocks.size()*sizeof(int32_t));
        // the plans bake the element offsets into every pair / tile record: a tensor with the same block list but another
        // layout (non-canonical offsets) must not reuse them
        k.append((const char*)offsets.data(),offsets.size()*sizeof(int64_t));
        k.append((const char*)&d.nelems,sizeof(int64_t));
        }
    };

//
// Plan cache: inside davidson the same few contractions / permuted adds repeat with identical block
// structure every iteration; the integer planning and table upload are done once per structure.
//
template<typename Plan, int (*Destroy)(Plan*)>
class PlanCache
    {
    std::unordered_map<std::string,Plan*> map_;
    std::deque<std::string> order_;
    size_t cap_;
    public:
    explicit PlanCache(size_t cap) : cap_(cap) { }
    Plan*
    find(std::string const& key)
        {
        auto it = map_.find(key);
        return it == map_.end() ? nullptr : it->second;
        }
    void
    insert(std::string const& key, Plan* p)
        {
        if(map_.size() >= cap_)
            {
            auto old = order_.front();
            order_.pop_front();
            auto it = map_.find(old);
            if(it != map_.end()) { Destroy(it->second); map_.erase(it); }
            }
        map_[key] = p;
        order_.push_back(key);
        }
    };

// Davidson repeats the same handful of structures many times at one bond; across bonds and sweeps the sector sizes
// keep changing, so a larger cache only grows the table pool: measured on the Hubbard 16x4 run, 4096 entries cost
// 25.9 s inside Contract against 13.9 s with 16 (every retained plan's device tables are a fresh cudaMalloc instead of
// a recycled pool block). 32 entries; ITB_PLAN_CACHE overrides.
static size_t
planCacheCap(size_t dflt)
    {
    if(auto* e = std::getenv("ITB_PLAN_CACHE")) return size_t(std::max(16l,std::atol(e)));
    return dflt;
    }
static PlanCache<itb_contract_plan,itb_contract_plan_destroy>&
contractCache() { static PlanCache<itb_contract_plan,itb_contract_plan_destroy> c(planCacheCap(32)); return c; }
static PlanCache<itb_permute_plan,itb_permute_plan_destroy>&
permuteCache() { static PlanCache<itb_permute_plan,itb_permute_plan_destroy> c(planCacheCap(32)); return c; }

// sliceIndex >= 0: the plan executes only rows [lo[s],hi[s]) of C's index sliceIndex (multi-GPU row sharding); returns
// nullptr when the planner cannot slice that index (ITB_ERR_UNSUPPORTED: the caller runs the contraction unsharded)
static itb_contract_plan*
getContractPlan(Desc const& dA, std::vector<int32_t> const& la, Desc const& dB, std::vector<int32_t> const& lb,
                int sliceIndex = -1, int64_t const* lo = nullptr, int64_t const* hi = nullptr, int nsect = 0)
    {
    std::string key;
    key.reserve(256);
    dA.appendKey(key);
    key.append((const char*)la.data(),la.size()*sizeof(int32_t));
    key.push_back('|');
    dB.appendKey(key);
    key.append((const char*)lb.data(),lb.size()*sizeof(int32_t));
    if(sliceIndex >= 0)
        {
        key.push_back('/');
        key.append((const char*)&sliceIndex,sizeof(int));
        key.append((const char*)lo,size_t(nsect)*sizeof(int64_t));
        key.append((const char*)hi,size_t(nsect)*sizeof(int64_t));
        }
    auto& cache = contractCache();
    if(auto* p = cache.find(key)) return p;
    itb_contract_plan* p = nullptr;
    gpu::Scope missScope("Contract: b' plan create (cache miss)");
    check(itb_contract_plan_create(&dA.d,la.data(),&dB.d,lb.data(),&p),"contract plan");
    if(sliceIndex >= 0)
        {
        auto rc = itb_contract_plan_set_index_slices(p,sliceIndex,lo,hi);
        if(rc == ITB_ERR_UNSUPPORTED) { itb_contract_plan_destroy(p); return nullptr; }
        check(rc,"contract plan slices");
        }
    cache.insert(key,p);
    return p;
    }

static itb_permute_plan*
getPermutePlan(Desc const& dS, Desc const& dD, std::vector<int32_t> const& perm)
    {
    std::string key;
    key.reserve(256);
    dS.appendKey(key);
    key.push_back('>');
    dD.appendKey(key);
    key.append((const char*)perm.data(),perm.size()*sizeof(int32_t));
    auto& cache = permuteCache();
    if(auto* p = cache.find(key)) return p;
    itb_permute_plan* p = nullptr;
    check(itb_permute_plan_create(&dS.d,&dD.d,perm.data(),&p),"permute plan");
    cache.insert(key,p);
    return p;
    }

static std::vector<int32_t>
toLabels(Labels const& L)
    {
    auto v = std::vector<int32_t>(L.size() > 0 ? L.size() : 1,0);
    for(auto i : range(L.size())) v[i] = int32_t(L[i]);
    return v;
    }

static std::vector<int32_t>
toPerm(Permutation const& P, long r)
    {
    auto v = std::vector<int32_t>(r > 0 ? r : 1,0);
    for(auto i : range(r)) v[i] = int32_t(P.dest(i));
    return v;
    }

// block list + sizes -> BlockOffsets (QDense::updateOffsets(is,blocks), qdense.cc:186-211)
static std::tuple<BlockOffsets,long>
offsetsFor(IndexSet const& is, Blocks const& blocks)
    {
    auto bofs = BlockOffsets();
    if(order(is)==0)
        {
        bofs.push_back(make_blof(Block(0),0));
        return std::make_tuple(bofs,1l);
        }
    long tot = 0;
    for(auto const& b : blocks)
        {
        long sz = 1;
        for(auto j : range(order(is))) sz *= is[j].blocksize0(b[j]);
        bofs.push_back(make_blof(b,tot));
        tot += sz;
        }
    return std::make_tuple(bofs,tot);
    }

//
// ---------------------------------------------------------------------------------------------
//  QDenseGPU
// ---------------------------------------------------------------------------------------------
//

// doTask(CalcDiv,QDense) qdense.cc:86-105
template<typename T>
QN
doTask(CalcDiv const& C, QDenseGPU<T> const& D)
    {
    if(order(C.is)==0 || D.offsets.empty()) return QN{};
    auto b = D.offsets.front().block;
    auto block_ind = Block(order(C.is));
    block_ind = b;
    return calcDiv(C.is,block_ind);
    }
template QN doTask(CalcDiv const&,QDenseGPU<Real> const&);
template QN doTask(CalcDiv const&,QDenseGPU<Cplx> const&);

// doTask(NormNoScale,QDense) qdense.cc:409-417
template<typename T>
Real
doTask(NormNoScale, QDenseGPU<T> const& d)
    {
    ITB_SCOPE("NormNoScale");
    double out = 0;
    check(itb_nrm2(context(),dtypeOf<T>(),int64_t(d.n),d.buf.data(),&out),"nrm2");
    return out;
    }
template Real doTask(NormNoScale,QDenseGPU<Real> const&);
template Real doTask(NormNoScale,QDenseGPU<Cplx> const&);

// doTask(Mult<Real>,QDense) qdense.cc:303-311
template<typename T>
void
doTask(Mult<Real> const& M, QDenseGPU<T>& d)
    {
    check(itb_scal(context(),dtypeOf<T>(),int64_t(d.n),d.buf.data(),M.x,0.),"scal");
    }
template void doTask(Mult<Real> const&,QDenseGPU<Real>&);
template void doTask(Mult<Real> const&,QDenseGPU<Cplx>&);

static void
realToCplx(gpu::Buffer const& src, gpu::Buffer & dst, size_t n)
    {
    check(itb_real_to_cplx(context(),int64_t(n),src.data(),dst.data()),"real_to_cplx");
    }

// doTask(Mult<Cplx>,QDenseReal const&,ManageStore&) qdense.cc:314-325
void
doTask(Mult<Cplx> const& M, QDenseGPUReal const& d, ManageStore& m)
    {
    auto* nd = m.makeNewData<QDenseGPUCplx>(d.offsets,d.n);
    realToCplx(d.buf,nd->buf,d.n);
    check(itb_scal(context(),ITB_C64,int64_t(d.n),nd->buf.data(),M.x.real(),M.x.imag()),"scal");
    }
void
doTask(Mult<Cplx> const& M, QDenseGPUCplx& d)
    {
    check(itb_scal(context(),ITB_C64,int64_t(d.n),d.buf.data(),M.x.real(),M.x.imag()),"scal");
    }

void
doTask(MakeCplx const&, QDenseGPUCplx&) { }
void
doTask(MakeCplx const&, QDenseGPUReal const& d, ManageStore& m)
    {
    auto* nd = m.makeNewData<QDenseGPUCplx>(d.offsets,d.n);
    realToCplx(d.buf,nd->buf,d.n);
    }

// doTask(Fill<T>,QDense) qdense.cc:356-373
void
doTask(Fill<Real> const& F, QDenseGPUReal& d)
    {
    check(itb_fill(context(),ITB_F64,int64_t(d.n),d.buf.data(),F.x,0.),"fill");
    }
void
doTask(Fill<Cplx> const& F, QDenseGPUCplx& d)
    {
    check(itb_fill(context(),ITB_C64,int64_t(d.n),d.buf.data(),F.x.real(),F.x.imag()),"fill");
    }
void
doTask(Fill<Real> const& F, QDenseGPUCplx const& d, ManageStore& m)
    {
    auto* nd = m.makeNewData<QDenseGPUReal>(d.offsets,d.n);
    check(itb_fill(context(),ITB_F64,int64_t(d.n),nd->buf.data(),F.x,0.),"fill");
    }
void
doTask(Fill<Cplx> const& F, QDenseGPUReal const& d, ManageStore& m)
    {
    auto* nd = m.makeNewData<QDenseGPUCplx>(d.offsets,d.n);
    check(itb_fill(context(),ITB_C64,int64_t(d.n),nd->buf.data(),F.x.real(),F.x.imag()),"fill");
    }

void
doTask(Conj, QDenseGPUCplx& d)
    {
    check(itb_conj(context(),int64_t(d.n),d.buf.data()),"conj");
    }

void
doTask(TakeReal, QDenseGPUCplx const& d, ManageStore& m)
    {
    auto* nd = m.makeNewData<QDenseGPUReal>(d.offsets,d.n);
    check(itb_take_part(context(),int64_t(d.n),d.buf.data(),nd->buf.data(),0),"take_real");
    }
void
doTask(TakeImag, QDenseGPUReal& d)
    {
    check(itb_fill(context(),ITB_F64,int64_t(d.n),d.buf.data(),0.,0.),"fill");
    }
void
doTask(TakeImag, QDenseGPUCplx const& d, ManageStore& m)
    {
    auto* nd = m.makeNewData<QDenseGPUReal>(d.offsets,d.n);
    check(itb_take_part(context(),int64_t(d.n),d.buf.data(),nd->buf.data(),1),"take_imag");
    }

// QDense::getEltBlockOffset qdense.h:455-495 -> one element read back from the device
template<typename T>
static Cplx
getEltGPU(GetElt& G, QDenseGPU<T> const& d)
    {
    ITB_SCOPE("GetElt");
    auto r = long(G.inds.size());
    double out[2] = {0.,0.};
    if(r == 0)
        {
        if(d.n == 0) return Cplx(0.,0.);
        check(itb_get_elt(context(),dtypeOf<T>(),d.buf.data(),0,out),"get_elt");
        return Cplx(out[0],out[1]);
        }
    long eoff = 0, estr = 1;
    auto block = Block(r);
    for(auto i = 0; i < r; ++i)
        {
        auto& I = G.is[i];
        long block_subind = 0, elt_subind = G.inds[i];
        while(elt_subind >= I.blocksize0(block_subind))
            {
            elt_subind -= I.blocksize0(block_subind);
            ++block_subind;
            }
        block[i] = block_subind;
        eoff += elt_subind*estr;
        estr *= I.blocksize0(block_subind);
        }
    auto boff = offsetOf(d.offsets,block);
    if(boff < 0) return Cplx(0.,0.);
    check(itb_get_elt(context(),dtypeOf<T>(),d.buf.data(),boff+eoff,out),"get_elt");
    return Cplx(out[0],out[1]);
    }
Cplx doTask(GetElt& G, QDenseGPUReal const& d) { return getEltGPU(G,d); }
Cplx doTask(GetElt& G, QDenseGPUCplx const& d) { return getEltGPU(G,d); }

// doTask(Order,QDense) / permuteQDense qdense.cc:847-893: the result holds EVERY flux-allowed block
template<typename T>
void
doTask(Order const& O, QDenseGPU<T>& dB)
    {
    ITB_SCOPE("Order");
    auto const& Ais = O.is1();
    auto r = order(Ais);
    auto bind = IndexSetBuilder(r);
    for(auto i : range(r)) bind.setIndex(O.perm().dest(i),Ais[i]);
    auto Bis = bind.build();
    auto div = doTask(CalcDiv{Ais},dB);
    auto [bofs,size] = getBlockOffsets(Bis,div);
    auto nB = QDenseGPU<T>(bofs,size);
    auto dS = Desc(Ais,dB.offsets,dB.n,dtypeOf<T>());
    auto dD = Desc(Bis,nB.offsets,nB.n,dtypeOf<T>());
    auto* plan = getPermutePlan(dS,dD,toPerm(O.perm(),r));
    check(itb_permute_run(context(),plan,dB.buf.data(),nB.buf.data(),1.,0.,0),"permute");
    dB = std::move(nB);
    }
template void doTask(Order const&,QDenseGPU<Real>&);
template void doTask(Order const&,QDenseGPU<Cplx>&);

//
// ---- multi-GPU row sharding of contractions -----------------------------------------------------------------
//
// cut the concatenated rows of all sectors into world() contiguous segments of equal weight (same rule as
// itensor_b200/shard.py row_partition): lo/hi[r*nsect+s] = rows of sector s owned by rank r
static void
rowPartition(std::vector<int64_t> const& sizes, std::vector<double> const& w, std::vector<int64_t>& lo, std::vector<int64_t>& hi)
    {
    const int W = gpu::world(), ns = int(sizes.size());
    std::vector<int64_t> start(ns+1,0);
    std::vector<double> cumw(ns+1,0.);
    for(int s = 0; s < ns; ++s) { start[s+1] = start[s]+sizes[s]; cumw[s+1] = cumw[s]+w[s]; }
    std::vector<int64_t> cuts(W+1,0);
    for(int g = 1; g < W; ++g)
        {
        double target = cumw[ns]*g/W;
        int s = 0;
        while(s+1 < ns && cumw[s+1] <= target) ++s;
        double per_row = sizes[s] > 0 ? w[s]/double(sizes[s]) : 0.;
        int64_t r = per_row > 0 ? int64_t(std::llround((target-cumw[s])/per_row)) : 0;
        if(r > 0 && r < sizes[s]) { int64_t ra = (r+4)/8*8; if(ra > 0 && ra < sizes[s]) r = ra; } // cuts on 8-row (DMMA fragment) boundaries
        r = std::min<int64_t>(std::max<int64_t>(r,0),sizes[s]);
        cuts[g] = std::max(start[s]+r,cuts[g-1]);
        }
    cuts[W] = start[ns];
    lo.assign(size_t(W)*ns,0); hi.assign(size_t(W)*ns,0);
    for(int g = 0; g < W; ++g)
        for(int s = 0; s < ns; ++s)
            {
            lo[size_t(g)*ns+s] = std::min(std::max(cuts[g]-start[s],int64_t(0)),sizes[s]);
            hi[size_t(g)*ns+s] = std::min(std::max(cuts[g+1]-start[s],int64_t(0)),sizes[s]);
            }
    }

// pack / scatter plans for the rows [lo,hi) of index j of a tensor with structure (is, off): rank r's rows go, block
// after block and each box contiguous, into segment r of a (world x segment) buffer
static std::shared_ptr<gpu::Pending>
makePending(IndexSet const& is, BlockOffsets const& off, int dtype, int j, Index const& shardIndex,
            std::vector<int64_t> const& lo, std::vector<int64_t> const& hi)
    {
    const int W = gpu::world(), r = order(is), ns = int(is[j].nblock());
    auto P = std::make_shared<gpu::Pending>();
    P->shardIndex = shardIndex; P->nsect = ns; P->lo = lo; P->hi = hi;
    std::string key;
    key.append((const char*)&dtype,sizeof(int)); key.append((const char*)&j,sizeof(int));
    for(auto i : range(r)) for(auto b : range(is[i].nblock())) { long e = is[i].blocksize0(b); key.append((const char*)&e,sizeof(long)); key.push_back(';'); }
    for(auto const& bo : off) { for(auto i : range(r)) { int32_t c = int32_t(bo.block[i]); key.append((const char*)&c,4); } key.append((const char*)&bo.offset,sizeof(bo.offset)); }
    key.append((const char*)lo.data(),lo.size()*sizeof(int64_t)); key.append((const char*)hi.data(),hi.size()*sizeof(int64_t));
    // segment length: the largest rank share, in elements
    std::vector<int64_t> seg(W,0);
    for(int g = 0; g < W; ++g)
        for(auto const& bo : off)
            {
            auto sct = bo.block[j];
            auto rows = hi[size_t(g)*ns+sct]-lo[size_t(g)*ns+sct];
            if(rows <= 0) continue;
            int64_t rest = 1;
            for(auto i : range(r)) if(int(i) != j) rest *= is[i].blocksize0(bo.block[i]);
            seg[g] += rows*rest;
            }
    int64_t segMax = 1;
    for(auto v : seg) segMax = std::max(segMax,v);
    P->segDoubles = segMax*(dtype == ITB_C64 ? 2 : 1);
    // small LRU of shared plans keyed on structure + partition (davidson gathers the same H*phi structure every iteration)
    static std::unordered_map<std::string,std::shared_ptr<gpu::CopyPlan>> cache;
    static std::deque<std::string> lru;
    auto build = [&](bool pack) -> std::shared_ptr<gpu::CopyPlan>
        {
        auto k = key; k.push_back(pack ? 'P' : 'U');
        auto hit = cache.find(k);
        if(hit != cache.end()) return hit->second;
        std::vector<itb_copy_item> items;
        for(int g = 0; g < W; ++g)
            {
            if(pack ? g != gpu::rank() : g == gpu::rank()) continue;
            int64_t run = pack ? 0 : int64_t(g)*segMax;
            for(auto const& bo : off)
                {
                auto sct = bo.block[j];
                auto a = lo[size_t(g)*ns+sct], e = hi[size_t(g)*ns+sct];
                if(e <= a) continue;
                itb_copy_item it;
                std::memset(&it,0,sizeof(it));
                it.n = r;
                int64_t full = 1, packed = 1;
                int64_t shift = 0;
                for(auto i : range(r))
                    {
                    int64_t ext = is[i].blocksize0(bo.block[i]);
                    int64_t box = int(i) == j ? e-a : ext;
                    it.ext[i] = box;
                    (pack ? it.sstr[i] : it.dstr[i]) = full;
                    (pack ? it.dstr[i] : it.sstr[i]) = packed;
                    if(int(i) == j) shift = a*full;
                    full *= ext; packed *= box;
                    }
                (pack ? it.s_off : it.d_off) = bo.offset+shift;
                (pack ? it.d_off : it.s_off) = run;
                run += packed;
                items.push_back(it);
                }
            }
        itb_permute_plan* p = nullptr;
        check(itb_blockcopy_plan_create(int64_t(items.size()),items.data(),dtype,dtype,&p),"row pack/scatter plan");
        auto sp = std::make_shared<gpu::CopyPlan>(p);
        if(lru.size() >= 64) { cache.erase(lru.front()); lru.pop_front(); }
        cache[k] = sp;
        lru.push_back(k);
        return sp;
        };
    P->pack = build(true);
    P->unpack = build(false);
    return P;
    }

// position of index I in an IndexSet, -1 if absent
static int
findIndexPos(IndexSet const& is, Index const& I)
    {
    for(auto i : range(order(is))) if(is[i] == I) return int(i);
    return -1;
    }

// doTask(Contract,QDense,QDense) qdense.cc:671-747 (+ getContractedOffsets, loopContractedBlocks, contract, gemm)
//
// With ITB_WORLD > 1 (one process per GPU, every process running the same program on the same data) large contractions
// are ROW-SHARDED: each rank computes the rows of one uncontracted index it owns and the result stays "pending" — it is
// only all-gathered when something needs the whole tensor (gpu::Buffer::data()). A contraction that finds a pending
// operand whose sharded index stays uncontracted simply continues on the same rows: phi*L -> *W1 -> *W2 -> *R
// (LocalOp::product, localop.h:346-362) therefore runs without communication and H*phi is re-replicated by ONE
// all-gather when davidson first touches it; environment tensors stay row-sharded from one bond to the next.
template<typename VA, typename VB>
static void
contractQ(Contract& Con,
          BlockOffsets const& Aoff, gpu::Buffer const& Abuf, size_t An,
          BlockOffsets const& Boff, gpu::Buffer const& Bbuf, size_t Bn,
          ManageStore& m)
    {
    using VC = common_type<VA,VB>;
    Labels Lind, Rind, Cind;
    ITB_SCOPE("Contract QDenseGPU");
    std::optional<gpu::Scope> sub; // finer host-side breakdown under ITB_PROFILE
    sub.emplace("Contract: a labels+desc");
    computeLabels(Con.Lis,order(Con.Lis),Con.Ris,order(Con.Ris),Lind,Rind);
    const bool sortResult = false;
    contractIS(Con.Lis,Lind,Con.Ris,Rind,Con.Nis,Cind,sortResult);

    auto dA = Desc(Con.Lis,Aoff,An,dtypeOf<VA>());
    auto dB = Desc(Con.Ris,Boff,Bn,dtypeOf<VB>());
    auto la = toLabels(Lind), lb = toLabels(Rind);
    sub.emplace("Contract: b plan lookup/create");
    auto* plan = getContractPlan(dA,la,dB,lb); // (pair enumeration + C structure; device tables are built on first run)
    sub.emplace("Contract: c result structure");
    struct { int32_t c_order = 0, c_dtype = 0; int64_t c_nblocks = 0, c_nelems = 0, npairs = 0; double flops = 0; } info;
    check(itb_contract_plan_shape(plan,&info.c_order,&info.c_dtype,&info.c_nblocks,&info.c_nelems,&info.npairs,&info.flops),"plan shape");
    auto rC = long(info.c_order);
    auto cb = std::vector<int32_t>(size_t(info.c_nblocks*rC)+1);
    auto co = std::vector<int64_t>(size_t(info.c_nblocks)+1);
    itb_contract_plan_c_blocks(plan,cb.data());
    itb_contract_plan_c_offsets(plan,co.data());
    auto Coffsets = BlockOffsets();
    Coffsets.reserve(info.c_nblocks);
    for(auto c : range(info.c_nblocks))
        {
        auto b = Block(rC);
        for(auto j : range(rC)) b[j] = cb[c*rC+j];
        Coffsets.push_back(make_blof(b,co[c]));
        }

    // ---- row sharding -------------------------------------------------------------------------------------------
    std::shared_ptr<gpu::Pending> next;
    itb_contract_plan* sliced = nullptr;
    if(gpu::world() > 1 && info.c_nblocks > 0 && rC > 0)
        {
        // sharding a contraction costs a second (sliced) plan and, sooner or later, an exchange that waits for the slowest rank:
        // measured on 2 GPUs, the Hubbard 16x4 ramp to maxdim 800 (contractions of 1e8-1e10 flops) ran 2x SLOWER with a
        // threshold of 2e8 (23.6 s vs 11.7 s on one GPU). Only contractions worth >= ~0.5 ms of device time are sharded.
        static const double min_flops = [] { auto* e = std::getenv("ITB_SHARD_MIN_FLOPS"); return e ? std::atof(e) : 1e10; }();
        auto pa = Abuf.pending(), pb = Bbuf.pending();
        if(pa && pb) { Bbuf.data(); pb.reset(); }       // two sharded operands: complete one of them
        auto& pend = pa ? pa : pb;
        int cpos = -1;
        std::vector<int64_t> lo, hi;
        Index shardIndex;
        if(pend)
            {
            // continue on the same rows if the sharded index survives this contraction, else complete the operand now
            cpos = findIndexPos(Con.Nis,pend->shardIndex);
            if(cpos >= 0) { lo = pend->lo; hi = pend->hi; shardIndex = pend->shardIndex; }
            else { (pa ? Abuf : Bbuf).data(); pend.reset(); }
            }
        if(cpos < 0 && !pend && info.flops >= min_flops)
            {
            // start sharding here: the primed index of largest dimension (the bra-side link of an effective-Hamiltonian
            // product survives the whole chain), else the largest index; weights = flops per row of every sector
            long best = -1;
            for(auto j : range(rC))
                {
                auto const& I = Con.Nis[j];
                if(dim(I) < 8*gpu::world()) continue;
                long score = dim(I)+(primeLevel(I) > 0 ? (1l << 40) : 0);
                if(score > best) { best = score; cpos = int(j); }
                }
            if(cpos >= 0)
                {
                shardIndex = Con.Nis[cpos];
                auto ns = shardIndex.nblock();
                std::vector<double> bf(size_t(info.c_nblocks),0.), w(size_t(ns),0.);
                check(itb_contract_plan_cblock_flops(plan,bf.data()),"cblock flops");
                for(auto c : range(info.c_nblocks)) w[size_t(cb[c*rC+cpos])] += bf[size_t(c)];
                std::vector<int64_t> sizes(size_t(ns),0);
                for(auto q : range(ns)) sizes[size_t(q)] = shardIndex.blocksize0(q);
                rowPartition(sizes,w,lo,hi);
                }
            }
        if(cpos >= 0)
            {
            auto ns = int(shardIndex.nblock());
            sliced = getContractPlan(dA,la,dB,lb,cpos,lo.data()+size_t(gpu::rank())*ns,hi.data()+size_t(gpu::rank())*ns,ns);
            if(sliced) next = makePending(Con.Nis,Coffsets,dtypeOf<VC>(),cpos,shardIndex,lo,hi);
            else if(pend) { (pa ? Abuf : Bbuf).data(); } // this step cannot be sliced on that index: complete the operand, run unsharded
            }
        }
    sub.emplace("Contract: d alloc result");
    auto* nd = m.makeNewData<QDenseGPU<VC>>(Coffsets,size_t(info.c_nelems));
    sub.emplace("Contract: e tables+launch");
    if(sliced)
        {
        // operands as they are: a pending operand contributes exactly the rows this rank owns
        check(itb_contract_run(context(),sliced,Abuf.rawData(),Bbuf.rawData(),nd->buf.rawData()),"contract (row-sharded)");
        nd->buf.setPending(next);
        }
    else check(itb_contract_run(context(),plan,Abuf.data(),Bbuf.data(),nd->buf.rawData()),"contract");
    }

template<typename VA, typename VB>
void
doTask(Contract& Con, QDenseGPU<VA> const& A, QDenseGPU<VB> const& B, ManageStore& m)
    {
    contractQ<VA,VB>(Con,A.offsets,A.buf,A.n,B.offsets,B.buf,B.n,m);
    }
template<typename VA, typename VB>
void
doTask(Contract& Con, QDenseGPU<VA> const& A, QDense<VB> const& B, ManageStore& m)
    {
    auto gB = QDenseGPU<VB>(B); // upload the host operand
    contractQ<VA,VB>(Con,A.offsets,A.buf,A.n,gB.offsets,gB.buf,gB.n,m);
    }
template<typename VA, typename VB>
void
doTask(Contract& Con, QDense<VA> const& A, QDenseGPU<VB> const& B, ManageStore& m)
    {
    auto gA = QDenseGPU<VA>(A);
    contractQ<VA,VB>(Con,gA.offsets,gA.buf,gA.n,B.offsets,B.buf,B.n,m);
    }
#define ITB_INST_CONTRACT(TA,TB) \
template void doTask(Contract&,QDenseGPU<TA> const&,QDenseGPU<TB> const&,ManageStore&); \
template void doTask(Contract&,QDenseGPU<TA> const&,QDense<TB> const&,ManageStore&); \
template void doTask(Contract&,QDense<TA> const&,QDenseGPU<TB> const&,ManageStore&);
ITB_INST_CONTRACT(Real,Real)
ITB_INST_CONTRACT(Real,Cplx)
ITB_INST_CONTRACT(Cplx,Real)
ITB_INST_CONTRACT(Cplx,Cplx)
#undef ITB_INST_CONTRACT

// QDiag written out as a block-sparse matrix: diagonal position i of an order-2 QDiag over (I0,I1) lies in block
// (sector of i in I0, sector of i in I1) at the in-sector positions; the blocks it touches become dense matrices.
template<typename T>
static QDenseGPU<T>
qdiagAsMatrix(QDiag<T> const& t, IndexSet const& tis)
    {
    auto n = std::min(dim(tis[0]),dim(tis[1]));
    struct Hit { long s0, s1, p0, p1; };
    auto hits = std::vector<Hit>(n);
    long s0 = 0, s1 = 0, b0 = 0, b1 = 0; // current sectors and their first positions
    for(auto i : range(n))
        {
        while(i >= b0+tis[0].blocksize0(s0)) { b0 += tis[0].blocksize0(s0); ++s0; }
        while(i >= b1+tis[1].blocksize0(s1)) { b1 += tis[1].blocksize0(s1); ++s1; }
        hits[i] = Hit{s0,s1,i-b0,i-b1};
        }
    // unique blocks in the reference's order (last index most significant, qdense.cc:60-65)
    auto blocks = std::vector<std::pair<long,long>>();
    for(auto& h : hits) blocks.emplace_back(h.s1,h.s0);
    std::sort(blocks.begin(),blocks.end());
    blocks.erase(std::unique(blocks.begin(),blocks.end()),blocks.end());
    auto offs = BlockOffsets();
    auto start = std::map<std::pair<long,long>,long>();
    long size = 0;
    for(auto& b : blocks)
        {
        auto blk = Block(2);
        blk[0] = b.second; blk[1] = b.first;
        offs.push_back(make_blof(blk,size));
        start[b] = size;
        size += tis[0].blocksize0(b.second)*tis[1].blocksize0(b.first);
        }
    auto h = QDense<T>();
    h.offsets = offs;
    h.store.assign(size_t(size),T(0.));
    for(auto i : range(n))
        {
        auto& x = hits[i];
        h.store[size_t(start[{x.s1,x.s0}] + x.p0 + x.p1*tis[0].blocksize0(x.s0))] = t.allSame() ? t.val : t.store[i];
        }
    return QDenseGPU<T>(h);
    }
template<typename TA, typename TB>
void
doTask(Contract& C, QDenseGPU<TA> const& d, QDiag<TB> const& t, ManageStore& m)
    {
    Labels Lind, Rind;
    computeLabels(C.Lis,order(C.Lis),C.Ris,order(C.Ris),Lind,Rind);
    if(order(C.Ris) == 2 && (Rind[0] < 0 || Rind[1] < 0))
        {
        auto g = qdiagAsMatrix(t,C.Ris);
        contractQ<TA,TB>(C,d.offsets,d.buf,d.n,g.offsets,g.buf,g.n,m);
        return;
        }
    doTask(C,d.toHost(),t,m);
    }
template<typename TA, typename TB>
void
doTask(Contract& C, QDiag<TA> const& t, QDenseGPU<TB> const& d, ManageStore& m)
    {
    Labels Lind, Rind;
    computeLabels(C.Lis,order(C.Lis),C.Ris,order(C.Ris),Lind,Rind);
    if(order(C.Lis) == 2 && (Lind[0] < 0 || Lind[1] < 0))
        {
        auto g = qdiagAsMatrix(t,C.Lis);
        contractQ<TA,TB>(C,g.offsets,g.buf,g.n,d.offsets,d.buf,d.n,m);
        return;
        }
    doTask(C,t,d.toHost(),m);
    }
#define ITB_INST_QDIAG(TA,TB) \
template void doTask(Contract&,QDenseGPU<TA> const&,QDiag<TB> const&,ManageStore&); \
template void doTask(Contract&,QDiag<TA> const&,QDenseGPU<TB> const&,ManageStore&);
ITB_INST_QDIAG(Real,Real)
ITB_INST_QDIAG(Real,Cplx)
ITB_INST_QDIAG(Cplx,Real)
ITB_INST_QDIAG(Cplx,Cplx)
#undef ITB_INST_QDIAG

// add(PlusEQ,QDense,QDense) qdense.cc:515-549:  A += alpha * permute(B)
template<typename TA, typename TB>
static void
addQ(PlusEQ const& P, IndexSet const& isA, QDenseGPU<TA>& A, IndexSet const& isB, QDenseGPU<TB> const& B,
     Permutation const& perm, Real alpha)
    {
    auto r = order(isA);
    if(std::is_same<TA,TB>::value && isTrivial(perm) && A.offsets.size() == B.offsets.size() && A.n == B.n)
        {
        bool same = true;
        for(auto i : range(A.offsets.size()))
            if(A.offsets[i].block != B.offsets[i].block) { same = false; break; }
        if(same) // daxpy fast path (qdense.cc:523-529)
            {
            check(itb_axpy(context(),dtypeOf<TA>(),int64_t(A.n),alpha,0.,B.buf.data(),A.buf.data()),"axpy");
            return;
            }
        }
    auto dS = Desc(isB,B.offsets,B.n,dtypeOf<TB>());
    auto dD = Desc(isA,A.offsets,A.n,dtypeOf<TA>());
    auto* plan = getPermutePlan(dS,dD,toPerm(perm,r));
    check(itb_permute_run(context(),plan,B.buf.data(),A.buf.data(),alpha,0.,1),"permute-accumulate");
    }

// doTask(PlusEQ,QDense,QDense,ManageStore&) qdense.cc:551-668 (block-list merge, real->complex promotion)
template<typename TA, typename TB>
static void
plusEqQ(PlusEQ const& P, QDenseGPU<TA> const& A, QDenseGPU<TB> const& B, ManageStore& m)
    {
    ITB_SCOPE("PlusEQ");
    if(B.n == 0) return;
    using TC = common_type<TA,TB>;
    auto r = order(P.is1());
    auto trivial = Permutation(r);

    if(r == 0)
        {
        if(isReal(A) && isCplx(B))
            {
            auto* nA = m.makeNewData<QDenseGPUCplx>(A.offsets,A.n);
            realToCplx(A.buf,nA->buf,A.n);
            check(itb_axpy(context(),ITB_C64,1,P.alpha(),0.,B.buf.data(),nA->buf.data()),"axpy");
            }
        else
            {
            auto* mA = m.modifyData(A);
            addQ(P,P.is1(),*mA,P.is2(),B,P.perm(),P.alpha());
            }
        return;
        }

    // permute and sort the blocks of B, then merge with A's (both sorted)
    auto Bblockps = Blocks(B.offsets.size(),Block(r));
    auto invperm = inverse(P.perm());
    for(auto ib : range(B.offsets.size()))
        for(auto i : range(r))
            Bblockps[ib][i] = B.offsets[ib].block[invperm.dest(i)];
    std::sort(Bblockps.begin(),Bblockps.end());
    auto Cblocks = Blocks();
    Cblocks.reserve(A.offsets.size()+B.offsets.size());
    size_t ia = 0, ib = 0;
    while(ia < A.offsets.size() && ib < B.offsets.size())
        {
        auto const& Ablock = A.offsets[ia].block;
        auto const& Bblockp = Bblockps[ib];
        if(Bblockp < Ablock) { Cblocks.push_back(Bblockp); ++ib; }
        else if(Ablock < Bblockp) { Cblocks.push_back(Ablock); ++ia; }
        else { Cblocks.push_back(Ablock); ++ia; ++ib; }
        }
    for(; ia < A.offsets.size(); ++ia) Cblocks.push_back(A.offsets[ia].block);
    for(; ib < B.offsets.size(); ++ib) Cblocks.push_back(Bblockps[ib]);

    if(A.offsets.size() < Cblocks.size())
        {
        // B has blocks A lacks: widen A's storage (zero-filled), copy A in, then accumulate B
        auto [bofs,size] = offsetsFor(P.is1(),Cblocks);
        auto* nA = m.makeNewData<QDenseGPU<TC>>(bofs,size_t(size));
        auto dS = Desc(P.is1(),A.offsets,A.n,dtypeOf<TA>());
        auto dD = Desc(P.is1(),nA->offsets,nA->n,dtypeOf<TC>());
        auto* plan = getPermutePlan(dS,dD,toPerm(trivial,r));
        check(itb_permute_run(context(),plan,A.buf.data(),nA->buf.data(),1.,0.,0),"widen");
        addQ(P,P.is1(),*nA,P.is2(),B,P.perm(),P.alpha());
        }
    else if(isReal(A) && isCplx(B))
        {
        auto* nA = m.makeNewData<QDenseGPUCplx>(A.offsets,A.n);
        realToCplx(A.buf,nA->buf,A.n);
        addQ(P,P.is1(),*nA,P.is2(),B,P.perm(),P.alpha());
        }
    else
        {
        auto* mA = m.modifyData(A);
        addQ(P,P.is1(),*mA,P.is2(),B,P.perm(),P.alpha());
        }
    }

// real += complex is rejected by the reference's Adder (qdense.cc:513); addQ is only instantiated for legal pairs
template<> void
addQ<Real,Cplx>(PlusEQ const&, IndexSet const&, QDenseGPU<Real>&, IndexSet const&, QDenseGPU<Cplx> const&, Permutation const&, Real)
    {
    Error("itensor_b200: cannot accumulate a complex tensor into real storage");
    }

template<typename TA, typename TB>
void
doTask(PlusEQ const& P, QDenseGPU<TA> const& A, QDenseGPU<TB> const& B, ManageStore& m) { plusEqQ(P,A,B,m); }
template<typename TA, typename TB>
void
doTask(PlusEQ const& P, QDenseGPU<TA> const& A, QDense<TB> const& B, ManageStore& m)
    {
    if(B.store.size() == 0) return;
    auto gB = QDenseGPU<TB>(B);
    plusEqQ(P,A,gB,m);
    }
template<typename TA, typename TB>
void
doTask(PlusEQ const& P, QDense<TA> const& A, QDenseGPU<TB> const& B, ManageStore& m)
    {
    // the accumulator is a host tensor: keep it on the host (reference path) with a downloaded B
    doTask(P,A,B.toHost(),m);
    }
#define ITB_INST_PLUSEQ(TA,TB) \
template void doTask(PlusEQ const&,QDenseGPU<TA> const&,QDenseGPU<TB> const&,ManageStore&); \
template void doTask(PlusEQ const&,QDenseGPU<TA> const&,QDense<TB> const&,ManageStore&); \
template void doTask(PlusEQ const&,QDense<TA> const&,QDenseGPU<TB> const&,ManageStore&);
ITB_INST_PLUSEQ(Real,Real)
ITB_INST_PLUSEQ(Real,Cplx)
ITB_INST_PLUSEQ(Cplx,Real)
ITB_INST_PLUSEQ(Cplx,Cplx)
#undef ITB_INST_PLUSEQ

//
// QCombiner on the device. The index bookkeeping mirrors combine()/uncombine() (qcombiner.cc:125-301):
// every stored block of d lands in (combine) / comes from (uncombine) a sub-range [start,end) of one
// sector of the combined index; each such move is one strided box for itb_blockcopy.
//
static void
runBlockCopy(std::vector<itb_copy_item> const& items, int dtype, void const* src, void* dst)
    {
    if(items.empty()) return;
    itb_permute_plan* plan = nullptr;
    check(itb_blockcopy_plan_create(int64_t(items.size()),items.data(),dtype,dtype,&plan),"blockcopy plan");
    auto rc = itb_permute_run(context(),plan,src,dst,1.,0.,0);
    itb_permute_plan_destroy(plan);
    check(rc,"blockcopy");
    }

template<typename T>
static void
combineGPU(QDenseGPU<T> const& d, QCombiner const& C, IndexSet const& dis, IndexSet const& Cis, IndexSet& Nis, ManageStore& m)
    {
    ITB_SCOPE("combine");
    auto dr = order(dis);
    auto ncomb = order(Cis)-1;
    auto nr = dr-ncomb+1;
    auto dperm = Labels(dr,-1);
    auto uncomb_dest = ncomb;
    for(auto i : range(dr))
        {
        auto jc = indexPosition(Cis,dis[i]);
        if(jc >= 0) dperm[i] = jc-1;
        else        dperm[i] = uncomb_dest++;
        }
    auto combined = [&dperm,ncomb](long i) { return dperm[i] < long(ncomb); };
    auto newind = IndexSetBuilder(nr);
    newind.nextIndex(Cis[0]);
    for(auto i : range(dr)) if(!combined(i)) newind.nextIndex(dis[i]);
    Nis = newind.build();

    auto [bofs,size] = getBlockOffsets(Nis,doTask(CalcDiv{dis},d));
    auto& nd = *m.makeNewData<QDenseGPU<T>>(bofs,size_t(size));
    nd.buf.zero();

    auto items = std::vector<itb_copy_item>();
    items.reserve(d.offsets.size());
    auto nblock = Block(nr);
    auto cblock = Block(ncomb);
    for(auto const& io : d.offsets)
        {
        size_t nu = 1;
        for(auto i : range(dr))
            {
            if(combined(i)) cblock[dperm[i]] = io.block[i];
            else            nblock[nu++] = io.block[i];
            }
        size_t start = 0, end = 0;
        std::tie(nblock[0],start,end) = C.getBlockRange(cblock);
        auto doff = offsetOf(nd.offsets,nblock);
        if(doff < 0) Error("itensor_b200: combine target block missing");
        // destination strides of the new block (column-major) and of the fused group inside dim 0
        auto nstr = std::vector<long>(nr,1);
        for(auto j : range(1,nr)) nstr[j] = nstr[j-1]*Nis[j-1].blocksize0(nblock[j-1]);
        auto cstr = std::vector<long>(ncomb,1);
        for(auto p : range(1,ncomb)) cstr[p] = cstr[p-1]*Cis[p].blocksize0(cblock[p-1]);
        itb_copy_item it;
        std::memset(&it,0,sizeof(it));
        it.s_off = io.offset;
        it.d_off = doff+long(start);
        it.n = int32_t(dr);
        long sstr = 1;
        size_t un = 1;
        for(auto i : range(dr))
            {
            auto e = dis[i].blocksize0(io.block[i]);
            it.ext[i] = e;
            it.sstr[i] = sstr;
            it.dstr[i] = combined(i) ? cstr[dperm[i]] : nstr[un++];
            sstr *= e;
            }
        items.push_back(it);
        }
    runBlockCopy(items,dtypeOf<T>(),d.buf.data(),nd.buf.data());
    }

template<typename T>
static void
uncombineGPU(QDenseGPU<T> const& d, QCombiner const& C, IndexSet const& dis, IndexSet const& Cis, IndexSet& Nis, ManageStore& m)
    {
    ITB_SCOPE("uncombine");
    auto& cind = Cis[0];
    auto dr = order(dis);
    auto cr = order(Cis);
    auto ncomb = cr-1;
    auto nr = dr-1+ncomb;
    decltype(dr) jc = 0;
    auto newind = IndexSetBuilder(nr);
    for(auto n : range(dr))
        {
        if(dis[n] == cind)
            {
            jc = n;
            for(auto k : range(1,cr)) newind.nextIndex(Cis[k]);
            }
        else newind.nextIndex(dis[n]);
        }
    Nis = newind.build();

    auto [bofs,size] = getBlockOffsets(Nis,doTask(CalcDiv{dis},d));
    auto& nd = *m.makeNewData<QDenseGPU<T>>(bofs,size_t(size));
    nd.buf.zero();

    auto items = std::vector<itb_copy_item>();
    auto nblock = Block(nr);
    for(auto const& io : d.offsets)
        {
        auto n = io.block[jc];
        auto sstrs = std::vector<long>(dr,1);
        for(auto i : range(1,dr)) sstrs[i] = sstrs[i-1]*dis[i-1].blocksize0(io.block[i-1]);
        for(auto oo : range(C.store_))
            {
            auto& br = C.store_[oo];
            if(long(br.block) != long(n)) continue;
            auto o = oo;
            for(auto k : range(ncomb-1))
                {
                nblock[jc+k] = o % Cis[1+k].nblock();
                o = (o-nblock[jc+k])/Cis[1+k].nblock();
                }
            nblock[jc+ncomb-1] = o;
            for(auto k : range(jc)) nblock[k] = io.block[k];
            for(auto k : range(1+jc,dr)) nblock[ncomb+k-1] = io.block[k];
            auto doff = offsetOf(nd.offsets,nblock);
            if(doff < 0) Error("itensor_b200: uncombine target block missing");
            auto nstr = std::vector<long>(nr,1);
            for(auto j : range(1,nr)) nstr[j] = nstr[j-1]*Nis[j-1].blocksize0(nblock[j-1]);
            itb_copy_item it;
            std::memset(&it,0,sizeof(it));
            it.s_off = io.offset+long(br.start)*sstrs[jc];
            it.d_off = doff;
            int q = 0;
            for(auto i : range(dr))
                {
                if(i == jc)
                    {
                    long sub = sstrs[jc];
                    for(auto k : range(ncomb))
                        {
                        auto e = Cis[1+k].blocksize0(nblock[jc+k]);
                        it.ext[q] = e;
                        it.sstr[q] = sub;
                        it.dstr[q] = nstr[jc+k];
                        sub *= e;
                        ++q;
                        }
                    }
                else
                    {
                    auto ni = i < jc ? i : i+ncomb-1;
                    it.ext[q] = dis[i].blocksize0(io.block[i]);
                    it.sstr[q] = sstrs[i];
                    it.dstr[q] = nstr[ni];
                    ++q;
                    }
                }
            it.n = q;
            if(q > ITB_MAX_ORDER) Error("itensor_b200: uncombine order too large");
            items.push_back(it);
            }
        }
    runBlockCopy(items,dtypeOf<T>(),d.buf.data(),nd.buf.data());
    }

template<typename T>
void
doTask(Contract& C, QDenseGPU<T> const& d, QCombiner const& cmb, ManageStore& m)
    {
    if(hasIndex(C.Lis,C.Ris[0])) uncombineGPU(d,cmb,C.Lis,C.Ris,C.Nis,m);
    else                         combineGPU(d,cmb,C.Lis,C.Ris,C.Nis,m);
    }
template<typename T>
void
doTask(Contract& C, QCombiner const& cmb, QDenseGPU<T> const& d, ManageStore& m)
    {
    if(hasIndex(C.Ris,C.Lis[0])) uncombineGPU(d,cmb,C.Ris,C.Lis,C.Nis,m);
    else                         combineGPU(d,cmb,C.Ris,C.Lis,C.Nis,m);
    }
template void doTask(Contract&,QDenseGPU<Real> const&,QCombiner const&,ManageStore&);
template void doTask(Contract&,QDenseGPU<Cplx> const&,QCombiner const&,ManageStore&);
template void doTask(Contract&,QCombiner const&,QDenseGPU<Real> const&,ManageStore&);
template void doTask(Contract&,QCombiner const&,QDenseGPU<Cplx> const&,ManageStore&);

// doTask(GetBlocks,QDense) decomp.cc:84-113: matrix views of the blocks of an order-2 tensor for the per-block
// LAPACK loops of diagHImpl / svdImpl. The views point into a host snapshot of the device data.
template<typename T>
std::vector<Ord2Block<T>>
doTask(GetBlocks<T> const& G, QDenseGPU<T> const& d)
    {
    ITB_SCOPE("GetBlocks");
    if(G.is.order() != 2) Error("doTask(GetBlocks,QDenseGPU) only supports 2-index tensors");
    d.mirror = std::make_shared<std::vector<T>>(d.n);
    d.buf.download(d.mirror->data(),d.n*sizeof(T));
    auto res = std::vector<Ord2Block<T>>{d.offsets.size()};
    size_t n = 0;
    for(auto const& dio : d.offsets)
        {
        auto& R = res[n++];
        auto nrow = G.is[0].blocksize0(dio.block[0]);
        auto ncol = G.is[1].blocksize0(dio.block[1]);
        R.i1 = dio.block[0];
        R.i2 = dio.block[1];
        R.M = makeMatRef(d.mirror->data()+dio.offset,d.n-dio.offset,nrow,ncol);
        }
    if(G.transpose)
        {
        for(auto& R : res)
            {
            R.M = transpose(R.M);
            std::swap(R.i1,R.i2);
            }
        }
    return res;
    }
template std::vector<Ord2Block<Real>> doTask(GetBlocks<Real> const&,QDenseGPU<Real> const&);
template std::vector<Ord2Block<Cplx>> doTask(GetBlocks<Cplx> const&,QDenseGPU<Cplx> const&);

//
// ---------------------------------------------------------------------------------------------
//  DenseGPU
// ---------------------------------------------------------------------------------------------
//
template<typename T>
Real
doTask(NormNoScale, DenseGPU<T> const& d)
    {
    double out = 0;
    check(itb_nrm2(context(),dtypeOf<T>(),int64_t(d.n),d.buf.data(),&out),"nrm2");
    return out;
    }
template Real doTask(NormNoScale,DenseGPU<Real> const&);
template Real doTask(NormNoScale,DenseGPU<Cplx> const&);

template<typename T>
void
doTask(Mult<Real> const& M, DenseGPU<T>& d)
    {
    check(itb_scal(context(),dtypeOf<T>(),int64_t(d.n),d.buf.data(),M.x,0.),"scal");
    }
template void doTask(Mult<Real> const&,DenseGPU<Real>&);
template void doTask(Mult<Real> const&,DenseGPU<Cplx>&);

void
doTask(Mult<Cplx> const& M, DenseGPUReal const& d, ManageStore& m)
    {
    auto* nd = m.makeNewData<DenseGPUCplx>(d.n);
    realToCplx(d.buf,nd->buf,d.n);
    check(itb_scal(context(),ITB_C64,int64_t(d.n),nd->buf.data(),M.x.real(),M.x.imag()),"scal");
    }
void
doTask(Mult<Cplx> const& M, DenseGPUCplx& d)
    {
    check(itb_scal(context(),ITB_C64,int64_t(d.n),d.buf.data(),M.x.real(),M.x.imag()),"scal");
    }
void
doTask(MakeCplx const&, DenseGPUCplx&) { }
void
doTask(MakeCplx const&, DenseGPUReal const& d, ManageStore& m)
    {
    auto* nd = m.makeNewData<DenseGPUCplx>(d.n);
    realToCplx(d.buf,nd->buf,d.n);
    }
void
doTask(Fill<Real> const& F, DenseGPUReal& d) { check(itb_fill(context(),ITB_F64,int64_t(d.n),d.buf.data(),F.x,0.),"fill"); }
void
doTask(Fill<Cplx> const& F, DenseGPUCplx& d) { check(itb_fill(context(),ITB_C64,int64_t(d.n),d.buf.data(),F.x.real(),F.x.imag()),"fill"); }
void
doTask(Fill<Real> const& F, DenseGPUCplx const& d, ManageStore& m)
    {
    auto* nd = m.makeNewData<DenseGPUReal>(d.n);
    check(itb_fill(context(),ITB_F64,int64_t(d.n),nd->buf.data(),F.x,0.),"fill");
    }
void
doTask(Fill<Cplx> const& F, DenseGPUReal const& d, ManageStore& m)
    {
    auto* nd = m.makeNewData<DenseGPUCplx>(d.n);
    check(itb_fill(context(),ITB_C64,int64_t(d.n),nd->buf.data(),F.x.real(),F.x.imag()),"fill");
    }
void
doTask(Conj, DenseGPUCplx& d) { check(itb_conj(context(),int64_t(d.n),d.buf.data()),"conj"); }
void
doTask(TakeReal, DenseGPUCplx const& d, ManageStore& m)
    {
    auto* nd = m.makeNewData<DenseGPUReal>(d.n);
    check(itb_take_part(context(),int64_t(d.n),d.buf.data(),nd->buf.data(),0),"take_real");
    }
void
doTask(TakeImag, DenseGPUReal& d) { check(itb_fill(context(),ITB_F64,int64_t(d.n),d.buf.data(),0.,0.),"fill"); }
void
doTask(TakeImag, DenseGPUCplx const& d, ManageStore& m)
    {
    auto* nd = m.makeNewData<DenseGPUReal>(d.n);
    check(itb_take_part(context(),int64_t(d.n),d.buf.data(),nd->buf.data(),1),"take_imag");
    }

// doTask(GetElt,Dense) dense.cc:37-47
Cplx
doTask(GetElt const& g, DenseGPUReal const& d)
    {
    double out[2] = {0.,0.};
    check(itb_get_elt(context(),ITB_F64,d.buf.data(),offset(g.is,g.inds),out),"get_elt");
    return Cplx(out[0],0.);
    }
Cplx
doTask(GetElt const& g, DenseGPUCplx const& d)
    {
    double out[2] = {0.,0.};
    check(itb_get_elt(context(),ITB_C64,d.buf.data(),offset(g.is,g.inds),out),"get_elt");
    return Cplx(out[0],out[1]);
    }

// doTask(Order,Dense) / permuteDense dense.cc:417-439
template<typename T>
void
doTask(Order const& O, DenseGPU<T>& dA)
    {
    auto r = order(O.is1());
    auto nB = DenseGPU<T>(dA.n);
    auto dS = Desc(O.is1(),dA.n,dtypeOf<T>());
    auto dD = Desc(O.is2(),nB.n,dtypeOf<T>());
    auto* plan = getPermutePlan(dS,dD,toPerm(O.perm(),r));
    check(itb_permute_run(context(),plan,dA.buf.data(),nB.buf.data(),1.,0.,0),"permute");
    dA = std::move(nB);
    }
template void doTask(Order const&,DenseGPU<Real>&);
template void doTask(Order const&,DenseGPU<Cplx>&);

// Dense combiner (combiner.cc:55-178) with the tensor resident in HBM
template<typename T>
static void
combineDenseGPU(DenseGPU<T> const& d, IndexSet const& dis, IndexSet const& Cis, IndexSet& Nis, ManageStore& m)
    {
    auto const& cind = Cis[0];
    auto dr = long(order(dis));
    auto cr = long(order(Cis));
    auto jc = indexPosition(dis,cind);
    if(jc >= 0) // the tensor carries the combined index: put the original indices back in its place
        {
        auto nb = IndexSetBuilder(dr+cr-2);
        for(auto j : range(dr))
            {
            if(j == jc) { for(auto k : range(1,cr)) nb.nextIndex(Cis[k]); }
            else nb.nextIndex(dis[j]);
            }
        Nis = nb.build();
        return;
        }
    // positions of the indices to fuse
    auto pos = std::vector<long>(cr-1);
    for(auto k : range(1,cr))
        {
        pos[k-1] = indexPosition(dis,Cis[k]);
        if(pos[k-1] < 0)
            {
            println("IndexSet of dense tensor = \n",dis);
            println("IndexSet of combiner/delta = \n",Cis);
            Error("Combiner: missing index (no contracted indices in combiner-tensor product)");
            }
        }
    bool in_place = true;
    for(auto k : range(1,cr-1)) if(pos[k] != pos[k-1]+1) in_place = false;
    if(in_place)
        {
        auto nb = IndexSetBuilder(dr+2-cr);
        for(auto j : range(pos[0])) nb.nextIndex(dis[j]);
        nb.nextIndex(cind);
        for(auto j : range(pos[0]+cr-1,dr)) nb.nextIndex(dis[j]);
        Nis = nb.build();
        return;
        }
    // fused indices to the front in combiner order, the others behind them in their old order
    auto P = Permutation(dr);
    auto taken = std::vector<bool>(dr,false);
    long dest = 0;
    for(auto k : range(cr-1)) { P.setFromTo(pos[k],dest++); taken[pos[k]] = true; }
    auto nb = IndexSetBuilder(dr+2-cr);
    nb.nextIndex(cind);
    for(auto j : range(dr))
        if(!taken[j]) { P.setFromTo(j,dest++); nb.nextIndex(dis[j]); }
    Nis = nb.build();
    auto pb = IndexSetBuilder(dr);
    for(auto j : range(dr)) pb.setIndex(P.dest(j),dis[j]);
    auto pis = pb.build();
    auto* nd = m.makeNewData<DenseGPU<T>>(d.n);
    auto dS = Desc(dis,d.n,dtypeOf<T>());
    auto dD = Desc(pis,d.n,dtypeOf<T>());
    auto* plan = getPermutePlan(dS,dD,toPerm(P,dr));
    check(itb_permute_run(context(),plan,d.buf.data(),nd->buf.data(),1.,0.,0),"combiner permute");
    }
template<typename T>
void
doTask(Contract& C, DenseGPU<T> const& d, Combiner const& cmb, ManageStore& m)
    {
    combineDenseGPU(d,C.Lis,C.Ris,C.Nis,m);
    }
template<typename T>
void
doTask(Contract& C, Combiner const& cmb, DenseGPU<T> const& d, ManageStore& m)
    {
    combineDenseGPU(d,C.Ris,C.Lis,C.Nis,m);
    if(!m.newData()) m.assignPointerRtoL();
    }
template void doTask(Contract&,DenseGPU<Real> const&,Combiner const&,ManageStore&);
template void doTask(Contract&,DenseGPU<Cplx> const&,Combiner const&,ManageStore&);
template void doTask(Contract&,Combiner const&,DenseGPU<Real> const&,ManageStore&);
template void doTask(Contract&,Combiner const&,DenseGPU<Cplx> const&,ManageStore&);

template<typename VA, typename VB>
static void contractD(Contract& C, void const* Adata, size_t An, void const* Bdata, size_t Bn, ManageStore& m);

// An order-2 diagonal tensor with one index contracted scales the slices of the dense tensor along that index and
// renames it; with both contracted it takes a (weighted) partial trace (TRG's normalisation, sample/src/trg.h:58). On
// the device both are the dense contraction with the diagonal written out as a (tiny) matrix: the n^2 zeros cost
// nothing next to keeping the big operand in HBM.
template<typename T>
static bool
scalesOneIndex(IndexSet const& tis, Labels const& lab) { return order(tis) == 2 && (lab[0] < 0 || lab[1] < 0); }
template<typename T>
static DenseGPU<T>
diagAsMatrix(Diag<T> const& t, IndexSet const& tis)
    {
    auto n0 = size_t(dim(tis[0])), n1 = size_t(dim(tis[1]));
    auto h = Dense<T>(n0*n1,T(0.));
    for(auto i : range(std::min(n0,n1))) h.store[i+n0*i] = t.allSame() ? t.val : t.store[i];
    return DenseGPU<T>(h);
    }

// Diag x Dense (diag.cc:128-207): is the diagonal tensor a unit delta over two equal-size indices of which exactly
// one is contracted? Then the product only renames that index.
template<typename T>
static bool
renamesOneIndex(Diag<T> const& t, IndexSet const& tis, Labels const& lab)
    {
    if(order(tis) != 2 || !t.allSame() || !(t.val == 1.) || dim(tis[0]) != dim(tis[1])) return false;
    return (lab[0] < 0) != (lab[1] < 0);
    }
template<typename TA, typename TB>
void
doTask(Contract& C, DenseGPU<TA> const& d, Diag<TB> const& t, ManageStore& m)
    {
    Labels Lind, Rind;
    computeLabels(C.Lis,order(C.Lis),C.Ris,order(C.Ris),Lind,Rind);
    if(renamesOneIndex(t,C.Ris,Rind)) { contractISReplaceIndex(C.Lis,Lind,C.Ris,Rind,C.Nis); return; } // storage untouched
    if(scalesOneIndex<TB>(C.Ris,Rind))
        {
        auto g = diagAsMatrix(t,C.Ris);
        contractD<TA,TB>(C,d.buf.data(),d.n,g.buf.data(),g.n,m);
        return;
        }
    doTask(C,d.toHost(),t,m);
    }
template<typename TA, typename TB>
void
doTask(Contract& C, Diag<TA> const& t, DenseGPU<TB> const& d, ManageStore& m)
    {
    Labels Lind, Rind;
    computeLabels(C.Lis,order(C.Lis),C.Ris,order(C.Ris),Lind,Rind);
    if(renamesOneIndex(t,C.Lis,Lind))
        {
        contractISReplaceIndex(C.Ris,Rind,C.Lis,Lind,C.Nis);
        m.makeNewData<DenseGPU<TB>>(d); // the result owns a (device) copy of the dense operand, as diag.cc:192 does
        return;
        }
    if(scalesOneIndex<TA>(C.Lis,Lind))
        {
        auto g = diagAsMatrix(t,C.Lis);
        contractD<TA,TB>(C,g.buf.data(),g.n,d.buf.data(),d.n,m);
        return;
        }
    doTask(C,t,d.toHost(),m);
    }
#define ITB_INST_DIAG(TA,TB) \
template void doTask(Contract&,DenseGPU<TA> const&,Diag<TB> const&,ManageStore&); \
template void doTask(Contract&,Diag<TA> const&,DenseGPU<TB> const&,ManageStore&);
ITB_INST_DIAG(Real,Real)
ITB_INST_DIAG(Real,Cplx)
ITB_INST_DIAG(Cplx,Real)
ITB_INST_DIAG(Cplx,Cplx)
#undef ITB_INST_DIAG

// doTask(Contract,Dense,Dense) dense.cc:262-330
template<typename VA, typename VB>
static void
contractD(Contract& C, void const* Adata, size_t An, void const* Bdata, size_t Bn, ManageStore& m)
    {
    using VC = common_type<VA,VB>;
    Labels Lind, Rind, Nind;
    computeLabels(C.Lis,C.Lis.order(),C.Ris,C.Ris.order(),Lind,Rind);
    auto wanted = C.Nis; // a caller may prescribe the result index order (dense.cc:287-303)
    IndexSet natural;
    contractIS(C.Lis,Lind,C.Ris,Rind,natural,Nind,false);
    auto dA = Desc(C.Lis,An,dtypeOf<VA>());
    auto dB = Desc(C.Ris,Bn,dtypeOf<VB>());
    auto* plan = getContractPlan(dA,toLabels(Lind),dB,toLabels(Rind));
    auto rsize = size_t(dim(natural));
    auto* nd = m.makeNewData<DenseGPU<VC>>(rsize);
    check(itb_contract_run(context(),plan,Adata,Bdata,nd->buf.data()),"contract");
    if(!wanted)
        {
        C.Nis = natural;
        return;
        }
    // permute the natural-order result into the prescribed order
    auto r = order(natural);
    auto P = Permutation(r);
    calcPerm(natural,wanted,P);
    if(isTrivial(P)) return;
    auto out = DenseGPU<VC>(rsize);
    auto dS = Desc(natural,rsize,dtypeOf<VC>());
    auto dD = Desc(wanted,rsize,dtypeOf<VC>());
    auto* pplan = getPermutePlan(dS,dD,toPerm(P,r));
    check(itb_permute_run(context(),pplan,nd->buf.data(),out.buf.data(),1.,0.,0),"permute");
    *nd = std::move(out);
    }
template<typename VA, typename VB>
void
doTask(Contract& C, DenseGPU<VA> const& A, DenseGPU<VB> const& B, ManageStore& m)
    {
    contractD<VA,VB>(C,A.buf.data(),A.n,B.buf.data(),B.n,m);
    }
template<typename VA, typename VB>
void
doTask(Contract& C, DenseGPU<VA> const& A, Dense<VB> const& B, ManageStore& m)
    {
    auto gB = DenseGPU<VB>(B);
    contractD<VA,VB>(C,A.buf.data(),A.n,gB.buf.data(),gB.n,m);
    }
template<typename VA, typename VB>
void
doTask(Contract& C, Dense<VA> const& A, DenseGPU<VB> const& B, ManageStore& m)
    {
    auto gA = DenseGPU<VA>(A);
    contractD<VA,VB>(C,gA.buf.data(),gA.n,B.buf.data(),B.n,m);
    }
#define ITB_INST_CONTRACT(TA,TB) \
template void doTask(Contract&,DenseGPU<TA> const&,DenseGPU<TB> const&,ManageStore&); \
template void doTask(Contract&,DenseGPU<TA> const&,Dense<TB> const&,ManageStore&); \
template void doTask(Contract&,Dense<TA> const&,DenseGPU<TB> const&,ManageStore&);
ITB_INST_CONTRACT(Real,Real)
ITB_INST_CONTRACT(Real,Cplx)
ITB_INST_CONTRACT(Cplx,Real)
ITB_INST_CONTRACT(Cplx,Cplx)
#undef ITB_INST_CONTRACT

// doTask(PlusEQ,Dense,Dense) dense.cc:371-415
template<typename TA, typename TB>
static void
addD(PlusEQ const& P, DenseGPU<TA>& A, DenseGPU<TB> const& B)
    {
    if(std::is_same<TA,TB>::value && isTrivial(P.perm()))
        {
        check(itb_axpy(context(),dtypeOf<TA>(),int64_t(A.n),P.alpha(),0.,B.buf.data(),A.buf.data()),"axpy");
        return;
        }
    auto r = order(P.is1());
    auto dS = Desc(P.is2(),B.n,dtypeOf<TB>());
    auto dD = Desc(P.is1(),A.n,dtypeOf<TA>());
    auto* plan = getPermutePlan(dS,dD,toPerm(P.perm(),r));
    check(itb_permute_run(context(),plan,B.buf.data(),A.buf.data(),P.alpha(),0.,1),"permute-accumulate");
    }
template<> void
addD<Real,Cplx>(PlusEQ const&, DenseGPU<Real>&, DenseGPU<Cplx> const&)
    {
    Error("itensor_b200: cannot accumulate a complex tensor into real storage");
    }
template<typename TA, typename TB>
static void
plusEqD(PlusEQ const& P, DenseGPU<TA> const& A, DenseGPU<TB> const& B, ManageStore& m)
    {
    if(isReal(A) && isCplx(B))
        {
        auto* nA = m.makeNewData<DenseGPUCplx>(A.n);
        realToCplx(A.buf,nA->buf,A.n);
        addD(P,*nA,B);
        }
    else
        {
        auto* mA = m.modifyData(A);
        addD(P,*mA,B);
        }
    }
template<typename TA, typename TB>
void
doTask(PlusEQ const& P, DenseGPU<TA> const& A, DenseGPU<TB> const& B, ManageStore& m) { plusEqD(P,A,B,m); }
template<typename TA, typename TB>
void
doTask(PlusEQ const& P, DenseGPU<TA> const& A, Dense<TB> const& B, ManageStore& m)
    {
    auto gB = DenseGPU<TB>(B);
    plusEqD(P,A,gB,m);
    }
template<typename TA, typename TB>
void
doTask(PlusEQ const& P, Dense<TA> const& A, DenseGPU<TB> const& B, ManageStore& m)
    {
    doTask(P,A,B.toHost(),m);
    }
#define ITB_INST_PLUSEQ(TA,TB) \
template void doTask(PlusEQ const&,DenseGPU<TA> const&,DenseGPU<TB> const&,ManageStore&); \
template void doTask(PlusEQ const&,DenseGPU<TA> const&,Dense<TB> const&,ManageStore&); \
template void doTask(PlusEQ const&,Dense<TA> const&,DenseGPU<TB> const&,ManageStore&);
ITB_INST_PLUSEQ(Real,Real)
ITB_INST_PLUSEQ(Real,Cplx)
ITB_INST_PLUSEQ(Cplx,Real)
ITB_INST_PLUSEQ(Cplx,Cplx)
#undef ITB_INST_PLUSEQ

} //namespace itensor
