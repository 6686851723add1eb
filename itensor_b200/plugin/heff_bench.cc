//
// heff_bench.cc — bench.py's workload through the PLUGIN: LocalOp::product's four chained contractions
// (itensor/mps/localop.h:346-362: phip = phi*L; phip *= W1; phip *= W2; phip *= R) as ITensor::operator* on
// QDenseGPU storage, i.e. through the reference's own doTask dispatch (itensor.cc:935-981 -> dotask.h -> gpu_storage.cc).
//
//   heff_bench <m> <nsect> <steps> <warmup> [cpu]
// Tensors have the structure of itensor_b200/synth.py heff_chain(gaussian_sectors(m,nsect)): S=1/2 sites (Sz +1/-1, size 1),
// MPO links with sectors (0:3, -2:1, +2:1), MPS links on an even-spaced Sz ladder; every flux-0 block stored, values from
// a seeded generator. Prints one JSON line:
//   resident : operands already in HBM (toGPU once), per step 4 operator* calls, device synchronised per step
//   e2e      : per step toGPU of all five HOST ITensors (pageable std::vector storage, as a user holds them), the four
//              products, toCPU of H*phi — what switching storage types costs a caller who keeps tensors on the host
//   [cpu]    : the same products on host QDense storage (the reference's own path), for the ratio in the same process
//
#include <chrono>
#include <cmath>
#include <sstream>

#include "itensor/all.h"
#include "gpu_convert.h"

using namespace itensor;

static std::vector<long>
gaussianSectors(long m, int nsect, double sigma = 1.25, double tilt = 0.15)
    {
    // itensor_b200/synth.py gaussian_sectors
    std::vector<double> w(nsect);
    double sum = 0;
    for(int i = 0; i < nsect; ++i) { double x = i-(nsect-1)/2.0+tilt; w[i] = std::exp(-x*x/(2*sigma*sigma)); sum += w[i]; }
    std::vector<long> s(nsect);
    for(int i = 0; i < nsect; ++i) s[i] = std::max(1l,std::lrint(m*w[i]/sum));
    return s;
    }

static Index
linkIndex(std::vector<long> const& sizes, Arrow dir, const char* tag)
    {
    auto qns = Index::qnstorage(sizes.size());
    int n = int(sizes.size());
    for(int i = 0; i < n; ++i) qns[i] = std::make_pair(QN({"Sz",2*(i-n/2)}),sizes[i]);
    return Index(std::move(qns),dir,tag);
    }
static Index
siteIndex(const char* tag)
    {
    auto qns = Index::qnstorage(2);
    qns[0] = std::make_pair(QN({"Sz",1}),1l);
    qns[1] = std::make_pair(QN({"Sz",-1}),1l);
    return Index(std::move(qns),Out,tag);
    }
static Index
mpoLink(const char* tag)
    {
    auto qns = Index::qnstorage(3);
    qns[0] = std::make_pair(QN({"Sz",0}),3l);
    qns[1] = std::make_pair(QN({"Sz",-2}),1l);
    qns[2] = std::make_pair(QN({"Sz",2}),1l);
    return Index(std::move(qns),Out,tag);
    }

template<typename F>
static double
timeIt(int steps, int warmup, F&& f)
    {
    for(int i = 0; i < warmup; ++i) f();
    auto t0 = std::chrono::steady_clock::now();
    for(int i = 0; i < steps; ++i) f();
    return std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count()/steps;
    }

int
main(int argc, char* argv[])
    {
    if(argc < 5) { println("usage: heff_bench <m> <nsect> <steps> <warmup> [cpu]"); return 2; }
    long m = std::atol(argv[1]);
    int nsect = std::atoi(argv[2]), steps = std::atoi(argv[3]), warmup = std::atoi(argv[4]);
    bool alsoCPU = argc > 5 && std::string(argv[5]) == "cpu";
    seedRNG(7);
    auto sizes = gaussianSectors(m,nsect);
    auto l = linkIndex(sizes,In,"l"), r = linkIndex(sizes,Out,"r");
    auto s1 = siteIndex("s1"), s2 = siteIndex("s2");
    auto k0 = mpoLink("k0"), k1 = mpoLink("k1"), k2 = mpoLink("k2");
    auto zero = QN({"Sz",0});
    auto phi = randomITensor(zero,l,s1,s2,r);
    auto L = randomITensor(zero,dag(l),k0,prime(l));
    auto W1 = randomITensor(zero,dag(k0),dag(s1),prime(s1),k1);
    auto W2 = randomITensor(zero,dag(k1),dag(s2),prime(s2),k2);
    auto R = randomITensor(zero,dag(r),dag(k2),prime(r));

    // flops = sum over block pairs of 2*M*N*K, counted from the block structure with the reference's own
    // bookkeeping: nnz(A-block) * (uncontracted size of the B partner)
    auto product = [](ITensor const& p, ITensor const& a, ITensor const& b, ITensor const& c, ITensor const& d)
        {
        auto x = p*a;
        x *= b;
        x *= c;
        x *= d;
        return x;
        };
    auto gphi = toGPU(phi), gL = toGPU(L), gW1 = toGPU(W1), gW2 = toGPU(W2), gR = toGPU(R);
    ITensor res;
    auto tres = timeIt(steps,warmup,[&]{ res = product(gphi,gL,gW1,gW2,gR); gpu::synchronize(); });
    auto launches0 = gpu::launchCount();
    ITensor hres;
    auto te2e = timeIt(steps,warmup,[&]
        {
        auto a = toGPU(phi), b = toGPU(L), c = toGPU(W1), d = toGPU(W2), e = toGPU(R);
        hres = toCPU(product(a,b,c,d,e));
        });
    auto launches = gpu::launchCount()-launches0;
    double tcpu = 0, err = -1;
    if(alsoCPU)
        {
        ITensor cres;
        tcpu = timeIt(1,0,[&]{ cres = product(phi,L,W1,W2,R); });
        err = norm(cres-hres)/norm(cres);
        }
    std::stringstream js;
    js.precision(12);
    js << "{\"what\": \"H_eff*phi through ITensor::operator* on QDenseGPU storage (plugin)\", \"maxdim\": " << m << ", \"sectors\": [";
    for(size_t i = 0; i < sizes.size(); ++i) js << (i ? "," : "") << sizes[i];
    js << "], \"steps\": " << steps << ", \"resident_ms_per_step\": " << tres*1e3 << ", \"e2e_ms_per_step\": " << te2e*1e3
       << ", \"h2d_bytes_per_step\": " << 8*(nnz(phi)+nnz(L)+nnz(W1)+nnz(W2)+nnz(R)) << ", \"d2h_bytes_per_step\": " << 8*nnz(hres)
       << ", \"gpu_launches_per_e2e_step\": " << double(launches)/steps << ", \"result_on_gpu\": " << (onGPU(res) ? "true" : "false");
    if(alsoCPU) js << ", \"cpu_ms_per_step\": " << tcpu*1e3 << ", \"rel_err_vs_cpu\": " << err;
    js << "}";
    println(js.str());
    return 0;
    }
