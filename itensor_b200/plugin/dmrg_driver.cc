//
// dmrg_driver.cc — the reference's UNMODIFIED dmrg() (itensor/mps/dmrg.h) on host or HBM-resident storage.
//
//   dmrg_driver <model> <N> <qn|dense> <cpu|gpu> <maxdims> <cutoffs> <niters> <noises> [json-out] [options]
//     options : --save <prefix>   write the final MPS and its site set (<prefix>.psi / .sites, the reference's binary format,
//                                 GPU storage goes through write(ostream,QDenseGPU))
//               --load <prefix>   start from that MPS instead of the product / random state: lets host and HBM storage run
//                                 the SAME sweep from the SAME state (per-bond parity at large bond dimension without the
//                                 trajectory sensitivity of a whole unconverged ramp)
//               --bonds <file>    per-bond record of every half sweep: energy, truncation error, kept spectrum
//     model   : heis_half | heis_one        (sample/dmrg.cc Hamiltonian, Neel product start)
//               hubbard                     (sample/hubbard_2d.cc: Nx x Ny cylinder, t=1, U=8 or $HUBBARD_U, Nf and Sz
//                                            conserved, seeded randomMPS start; <N> is written NxxNy, e.g. 16x4)
//     lists   : comma separated, one entry per sweep (last entry repeats)
// Prints one JSON line: final energy, per-sweep energy / wall seconds / max truncation error / max link
// dimension, per-bond truncation errors of the last sweep and the kept density-matrix spectrum at the
// centre bond of the last sweep (the per-bond comparison points of the parity tests).
//
#include <chrono>
#include <fstream>
#include <sstream>

#include "itensor/all.h"
#include "gpu_convert.h"

using namespace itensor;

static std::vector<double>
parseList(std::string const& s)
    {
    std::vector<double> v;
    std::stringstream ss(s);
    std::string tok;
    while(std::getline(ss,tok,',')) v.push_back(std::stod(tok));
    return v;
    }

struct SweepRecord { double energy = 0, seconds = 0, maxtrunc = 0; int maxlink = 0; };
struct BondRecord { int sweep = 0, half = 0, bond = 0; double energy = 0, truncerr = 0; std::vector<double> spectrum; };

class TimingObserver : public DMRGObserver
    {
    MPS const& psi_;
    MPS* pin_ = nullptr; // GPU runs: site tensors are re-pinned in HBM after every bond update
    std::chrono::steady_clock::time_point t0_;
    double maxtrunc_ = 0;
    int N_;
    public:
    std::vector<SweepRecord> sweeps;
    std::vector<double> lastTrunc;   // per bond, last completed half sweep pair
    std::vector<double> centreSpec;  // kept spectrum at the centre bond, last sweep (right-to-left pass)
    std::vector<BondRecord> bonds;   // --bonds: every bond of every half sweep
    bool recordBonds = false;

    TimingObserver(MPS& psi, Args const& args, bool pin) : DMRGObserver(psi,args), psi_(psi), pin_(pin ? &psi : nullptr), N_(length(psi))
        {
        t0_ = std::chrono::steady_clock::now();
        }

    void
    measure(Args const& args = Args::global())
        {
        auto b = args.getInt("AtBond");
        auto ha = args.getInt("HalfSweep");
        auto terr = args.getReal("Truncerr",0.);
        if(pin_) { pinToGPU(*pin_,b); pinToGPU(*pin_,b+1); }
        if(recordBonds)
            {
            BondRecord r;
            r.sweep = args.getInt("Sweep",0); r.half = ha; r.bond = b;
            r.energy = args.getReal("Energy",0.); r.truncerr = terr;
            for(auto const& e : spectrum().eigsKept()) r.spectrum.push_back(e);
            bonds.push_back(std::move(r));
            }
        if(b == 1 && ha == 1) { maxtrunc_ = 0; lastTrunc.assign(2*(N_-1),0.); }
        maxtrunc_ = std::max(maxtrunc_,terr);
        auto slot = (ha == 1) ? (b-1) : (N_-1)+(N_-1-b);
        if(slot >= 0 && slot < int(lastTrunc.size())) lastTrunc[slot] = terr;
        if(b == N_/2 && ha == 2)
            {
            centreSpec.clear();
            for(auto const& e : spectrum().eigsKept()) centreSpec.push_back(e);
            }
        if(b == 1 && ha == 2)
            {
            auto now = std::chrono::steady_clock::now();
            SweepRecord r;
            r.energy = args.getReal("Energy");
            r.seconds = std::chrono::duration<double>(now-t0_).count();
            r.maxtrunc = maxtrunc_;
            r.maxlink = maxLinkDim(psi_);
            sweeps.push_back(r);
            t0_ = now;
            }
        }
    };

int
main(int argc, char* argv[])
    {
    if(argc < 9)
        {
        println("usage: dmrg_driver <heis_half|heis_one|hubbard> <N|NxxNy> <qn|dense> <cpu|gpu> <maxdims> <cutoffs> <niters> <noises> [json-out]");
        return 2;
        }
    auto model = std::string(argv[1]);
    int N = std::atoi(argv[2]);
    int Nx = 0, Ny = 0;
    if(model == "hubbard")
        {
        auto a = std::string(argv[2]);
        auto x = a.find('x');
        if(x == std::string::npos) { println("hubbard: <N> must be NxxNy"); return 2; }
        Nx = std::atoi(a.substr(0,x).c_str());
        Ny = std::atoi(a.substr(x+1).c_str());
        N = Nx*Ny;
        }
    bool qn = std::string(argv[3]) == "qn";
    bool useGPU = std::string(argv[4]) == "gpu";
    auto maxdim = parseList(argv[5]), cutoff = parseList(argv[6]), niter = parseList(argv[7]), noise = parseList(argv[8]);
    std::string jsonOut, savePrefix, loadPrefix, bondsFile;
    // one process per GPU (tools/run_ranks.sh): every rank runs the same program; ranks > 0 write their files with a suffix
    auto rankSuffix = std::string();
    if(auto* e = std::getenv("ITB_RANK")) if(std::atoi(e) > 0) rankSuffix = std::string(".rank")+e;
    for(int a = 9; a < argc; ++a)
        {
        auto opt = std::string(argv[a]);
        if(opt == "--save" && a+1 < argc) savePrefix = argv[++a];
        else if(opt == "--load" && a+1 < argc) loadPrefix = argv[++a];
        else if(opt == "--bonds" && a+1 < argc) bondsFile = argv[++a];
        else if(opt.rfind("--",0) != 0 && jsonOut.empty()) jsonOut = opt;
        else { println("dmrg_driver: unknown option ",opt); return 2; }
        }
    auto nsweep = int(maxdim.size());
    auto at = [](std::vector<double> const& v, int i) { return v[std::min<size_t>(i,v.size()-1)]; };

    SiteSet sites;
    if(!loadPrefix.empty())
        {
        // same site indices as the saved state (index ids are part of the file)
        if(model == "heis_half") sites = readFromFile<SpinHalf>(loadPrefix+".sites");
        else if(model == "hubbard") sites = readFromFile<Electron>(loadPrefix+".sites");
        else sites = readFromFile<SpinOne>(loadPrefix+".sites");
        }
    else if(model == "heis_half") sites = SpinHalf(N,{"ConserveQNs=",qn});
    else if(model == "hubbard") sites = Electron(N,{"ConserveQNs=",qn});
    else sites = SpinOne(N,{"ConserveQNs=",qn});
    auto ampo = AutoMPO(sites);
    if(model == "hubbard")
        {
        // sample/hubbard_2d.cc:19-35: nearest-neighbour hopping on a cylinder plus on-site repulsion
        Real U = 8., t = 1.;
        if(auto* e = std::getenv("HUBBARD_U")) U = std::atof(e);
        for(auto bnd : squareLattice(Nx,Ny,{"YPeriodic=",true}))
            for(auto* sp : {"up","dn"})
                {
                auto cd = std::string("Cdag")+sp, c = std::string("C")+sp;
                ampo += -t,cd,bnd.s1,c,bnd.s2;
                ampo += -t,cd,bnd.s2,c,bnd.s1;
                }
        for(auto j : range1(N)) ampo += U,"Nupdn",j;
        }
    else
        {
        for(auto j : range1(N-1))
            {
            ampo += 0.5,"S+",j,"S-",j+1;
            ampo += 0.5,"S-",j,"S+",j+1;
            ampo +=     "Sz",j,"Sz",j+1;
            }
        }
    auto H = toMPO(ampo);
    // The reference's RNG is seeded from time()+pid (detail/algs.h:87-93) and davidson RANDOMIZES the new basis vector when
    // Gram-Schmidt leaves less than 1e-10 of it (iterativesolvers.h:325-331), which happens at every bond of a nearly
    // converged sweep: two runs of the same binary then differ by ~1e-11..1e-9 in the kept spectra. A fixed seed makes runs
    // (and the host-vs-HBM storage comparison) reproducible; both storages draw the same numbers in the same order.
    seedRNG(1);
    auto state = InitState(sites);
    for(auto i : range1(N)) state.set(i,i%2==1 ? "Up" : "Dn");
    auto psi = MPS(state);
    if(model == "hubbard")
        {
        seedRNG(1); // randomMPS is unseeded otherwise (mps.cc:281-287)
        psi = randomMPS(state);
        }
    if(!loadPrefix.empty()) psi = readFromFile<MPS>(loadPrefix+".psi",sites);
    // DMRG_PERTURB=eps: multiply every element of the starting MPS by (1 + eps*u), u uniform in (-1,1) from a fixed-seed
    // generator of our own. Measures how far a rounding-sized change of the INPUT moves the per-bond record of a sweep,
    // i.e. the amplification any two arithmetically different implementations (summation order, FMA use) are subject to.
    if(auto* e = std::getenv("DMRG_PERTURB"))
        {
        Real eps = std::atof(e);
        unsigned long long st = 0x9E3779B97F4A7C15ull;
        auto next = [&st]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (st >> 11) * (1.0/9007199254740992.0); };
        for(auto j : range1(N))
            {
            auto l = psi.leftLim(), r = psi.rightLim();
            psi.ref(j).apply([&](Real x) { return x*(1.+eps*(2.*next()-1.)); });
            psi.leftLim(l); psi.rightLim(r);
            }
        }

    auto sweeps = Sweeps(nsweep);
    for(int s = 1; s <= nsweep; ++s)
        {
        sweeps.setmaxdim(s,int(at(maxdim,s-1)));
        sweeps.setcutoff(s,at(cutoff,s-1));
        sweeps.setniter(s,int(at(niter,s-1)));
        sweeps.setnoise(s,at(noise,s-1));
        }

    if(useGPU)
        {
        toGPU(H);
        toGPU(psi);
        }

#ifdef DRIVER_VERBOSE
    auto args = Args("Quiet",true);
#else
    auto args = Args("Silent",true);
#endif
    // disk spilling of the environments exactly as the reference does it (dmrg.h:390-401, localmpo.h:620-681): with
    // GPU storage the environment tensors go through write(ostream,QDenseGPU<T>) and come back as host tensors
    if(auto* e = std::getenv("DMRG_WRITE_DIM")) args.add("WriteDim",std::atoi(e));
    if(auto* e = std::getenv("DMRG_WRITE_DIR")) args.add("WriteDir",std::string(e));
    auto t0 = std::chrono::steady_clock::now();
    auto PH = LocalMPO(H,args);
    auto obs = TimingObserver(psi,args,useGPU);
    obs.recordBonds = !bondsFile.empty();
    auto energy = DMRGWorker(psi,PH,sweeps,obs,args);
    if(useGPU) gpu::synchronize();
    auto total = std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();

    std::stringstream js;
    js.precision(17);
    js << "{\"model\": \"" << model << "\", \"N\": " << N << ", \"qn\": " << (qn ? "true" : "false")
       << ", \"storage\": \"" << (useGPU ? "gpu" : "cpu") << "\", \"energy\": " << energy
       << ", \"total_seconds\": " << total << ", \"gpu_launches\": " << (useGPU ? gpu::launchCount() : 0) << ", \"sweeps\": [";
    for(size_t i = 0; i < obs.sweeps.size(); ++i)
        {
        auto& r = obs.sweeps[i];
        js << (i ? ", " : "") << "{\"energy\": " << r.energy << ", \"seconds\": " << r.seconds << ", \"maxtrunc\": " << r.maxtrunc
           << ", \"maxlink\": " << r.maxlink << "}";
        }
    js << "], \"last_sweep_truncerr\": [";
    for(size_t i = 0; i < obs.lastTrunc.size(); ++i) js << (i ? ", " : "") << obs.lastTrunc[i];
    js << "], \"centre_spectrum\": [";
    for(size_t i = 0; i < obs.centreSpec.size(); ++i) js << (i ? ", " : "") << obs.centreSpec[i];
    js << "]}";
    println(js.str());
    if(!jsonOut.empty())
        {
        std::ofstream f(jsonOut+rankSuffix);
        f << js.str() << "\n";
        }
    if(!bondsFile.empty())
        {
        std::ofstream f(bondsFile+rankSuffix);
        f.precision(17);
        f << "[";
        for(size_t i = 0; i < obs.bonds.size(); ++i)
            {
            auto& r = obs.bonds[i];
            f << (i ? ",\n" : "\n") << "{\"sweep\": " << r.sweep << ", \"half\": " << r.half << ", \"bond\": " << r.bond
              << ", \"energy\": " << r.energy << ", \"truncerr\": " << r.truncerr << ", \"spectrum\": [";
            for(size_t k = 0; k < r.spectrum.size(); ++k) f << (k ? "," : "") << r.spectrum[k];
            f << "]}";
            }
        f << "\n]\n";
        }
    if(!savePrefix.empty() && rankSuffix.empty())
        {
        // GPU-resident site tensors are written through write(ostream,QDenseGPU<T>) (gpu_storage.h): the file holds the
        // host wire format and reads back as ordinary host tensors
        writeToFile(savePrefix+".sites",sites);
        writeToFile(savePrefix+".psi",psi);
        }
    return 0;
    }
