//
// dense_ops_check.cc — host-logic check of the DenseGPU combiner / delta / diagonal-scaling overloads against the
// reference's own host implementations of the same ITensor products (run on the oracle-backed mock of the C ABI in
// `pytest -m "not gpu"`, and on the real library in `-m gpu`). Prints one line per case and exits non-zero on mismatch.
//
#include "itensor/all.h"
#include "gpu_convert.h"

using namespace itensor;

static int failures = 0;

static void
same(const char* what, ITensor const& host, ITensor const& dev, bool expect_gpu = true)
    {
    auto back = toCPU(dev);
    auto diff = norm(host-back);
    auto ok = diff <= 1e-12*std::max(1.,norm(host)) && hasSameInds(inds(host),inds(back)) && (!expect_gpu || onGPU(dev));
    printfln("%-44s |diff| %.2e  on_gpu %d  %s",what,diff,int(onGPU(dev)),ok ? "ok" : "FAIL");
    if(!ok) ++failures;
    }

int
main()
    {
    seedRNG(7);
    auto i = Index(3,"i"), j = Index(4,"j"), k = Index(5,"k"), l = Index(2,"l");
    for(int cplx = 0; cplx < 2; ++cplx)
        {
        auto T = cplx ? randomITensorC(i,j,k,l) : randomITensor(i,j,k,l);
        auto G = toGPU(T);
        // element access on GPU storage: set() / apply() (SetElt, ApplyIT) keep the tensor in HBM, also when the storage is
        // shared with a copy (copy-on-write) and when set() promotes real storage to complex
        {
        auto Th = T, Gs = G;
        auto Gshared = Gs; // second owner: the write below must not change it
        Th.set(i=2,j=3,k=4,l=1,0.75); Gs.set(i=2,j=3,k=4,l=1,0.75);
        same("set(real) on GPU storage",Th,Gs);
        same("  ... shared copy untouched",T,Gshared);
        if(!cplx) { Th.apply([](Real x) { return 2.*x+1.; }); Gs.apply([](Real x) { return 2.*x+1.; }); same("apply(real f)",Th,Gs); }
        auto Tc = T, Gc = G;
        Tc.set(i=1,j=1,k=2,l=2,Cplx(0.5,-1.5)); Gc.set(i=1,j=1,k=2,l=2,Cplx(0.5,-1.5));
        same("set(complex) (promotes real storage)",Tc,Gc);
        Tc.apply([](Cplx z) { return z*Cplx(0.,1.); }); Gc.apply([](Cplx z) { return z*Cplx(0.,1.); });
        same("apply(complex f)",Tc,Gc);
        }
        // combiner: fused indices adjacent and in order (relabelling), scattered (device permute), and uncombining
        {
        auto [C1,c1] = combiner(j,k);
        same("combine adjacent (j,k)",T*C1,G*C1);
        same("combine adjacent, combiner on the left",C1*T,C1*G);
        same("uncombine",(T*C1)*dag(C1),(G*C1)*dag(C1));
        auto [C2,c2] = combiner(l,i);
        same("combine scattered (l,i)",T*C2,G*C2);
        auto [C3,c3] = combiner(k,j,i);
        same("combine reversed (k,j,i)",T*C3,G*C3);
        same("uncombine after permuting combine",(T*C3)*dag(C3),(G*C3)*dag(C3));
        }
        // delta: index replacement (metadata only), both operand orders
        {
        auto jp = prime(j);
        same("delta renames j -> j'",T*delta(j,jp),G*delta(j,jp));
        same("delta on the left",delta(j,jp)*T,delta(j,jp)*G);
        }
        // diagonal tensor with one contracted index: scales slices and renames
        {
        auto m = Index(4,"m");
        auto D = ITensor(j,m);
        for(auto n : range1(4)) D.set(n,n,0.5+n);
        auto Dg = diagITensor(std::vector<Real>{1.5,2.5,3.5,4.5},j,m);
        same("diag scaling T*D",T*Dg,G*Dg);
        same("diag scaling D*T",Dg*T,Dg*G);
        same("diag vs dense matrix",T*D,G*Dg);
        }
        // partial trace over two indices of equal size, and the full trace down to a scalar
        {
        auto jj = Index(4,"jj");
        auto S = cplx ? randomITensorC(j,jj,k) : randomITensor(j,jj,k);
        same("partial trace with delta(j,jj)",S*delta(j,jj),toGPU(S)*delta(j,jj));
        same("partial trace, delta on the left",delta(j,jj)*S,delta(j,jj)*toGPU(S));
        auto S2 = cplx ? randomITensorC(j,jj) : randomITensor(j,jj);
        auto th = eltC(S2*delta(j,jj)), tg = eltC(toGPU(S2)*delta(j,jj));
        auto ok = std::abs(th-tg) <= 1e-13*std::max(1.,std::abs(th));
        printfln("%-44s |diff| %.2e  %s","full trace to a scalar",std::abs(th-tg),ok ? "ok" : "FAIL");
        if(!ok) ++failures;
        }
        }
    // QDiag x QDenseGPU (svdBond's A *= D): the diagonal factor of a host SVD against the device-resident factors
    {
    auto I = Index(QN({"Sz",-1}),3,QN({"Sz",0}),4,QN({"Sz",1}),2,Out,"I");
    auto J = Index(QN({"Sz",-1}),2,QN({"Sz",0}),5,QN({"Sz",1}),3,Out,"J");
    auto K = Index(QN({"Sz",-1}),2,QN({"Sz",1}),2,Out,"K");
    for(int cplx = 0; cplx < 2; ++cplx)
        {
        auto T = cplx ? randomITensorC(QN({"Sz",0}),I,dag(J),K) : randomITensor(QN({"Sz",0}),I,dag(J),K);
        ITensor U(I,K), D, V;
        svd(T,U,D,V,{"MaxDim",6});
        same("QDiag: V*D",V*D,toGPU(V)*D);
        same("QDiag: D*V",D*V,D*toGPU(V));
        same("QDiag: U*D",U*D,toGPU(U)*D);
        same("QDiag: U*D*V",U*D*V,toGPU(U)*D*toGPU(V));
        }
    }
    // svd / factor of a dense GPU tensor (device svdOrd2): reconstruction, spectrum and truncation vs the host run.
    // (skipped on the mock ABI, which has no device solver: ITB_SVD_DEVICE=0)
    if(!(std::getenv("ITB_SVD_DEVICE") && std::atoi(std::getenv("ITB_SVD_DEVICE")) == 0))
        {
        auto a = Index(6,"a"), b = Index(5,"b"), c = Index(7,"c"), e = Index(4,"e");
        for(int cplx = 0; cplx < 2; ++cplx)
            {
            auto T = cplx ? randomITensorC(a,b,c,e) : randomITensor(a,b,c,e);
            auto G = toGPU(T);
            for(int pass = 0; pass < 3; ++pass)
                {
                // pass 0: U = (a,b) leading indices; pass 1: U = (c,a) scattered (stored matrix is transposed / permuted);
                // pass 2: truncated to 9 singular values
                ITensor Uh(a,b), Dh, Vh, Ug(a,b), Dg, Vg;
                if(pass == 1) { Uh = ITensor(c,a); Ug = ITensor(c,a); }
                auto args = pass == 2 ? Args("MaxDim",9,"Cutoff",0.) : Args::global();
                auto sh = svd(T,Uh,Dh,Vh,args);
                auto sg = svd(G,Ug,Dg,Vg,args);
                auto rec_h = Uh*Dh*Vh, rec_g = toCPU(Ug)*Dg*toCPU(Vg);
                auto dspec = 0.;
                for(auto n : range1(std::min(sh.numEigsKept(),sg.numEigsKept()))) dspec = std::max(dspec,std::fabs(sh.eig(n)-sg.eig(n)));
                auto ok = onGPU(Ug) && onGPU(Vg) && sh.numEigsKept() == sg.numEigsKept() && dspec <= 1e-12*sh.eig(1)
                          && norm(rec_h-rec_g) <= 1e-11*norm(T) && std::fabs(sh.truncerr()-sg.truncerr()) <= 1e-12;
                printfln("svd dense %s pass %d: kept %d/%d  |dspec| %.1e  |rec_h-rec_g| %.1e  truncerr %.3e/%.3e  %s",cplx ? "cplx" : "real",pass,
                         sg.numEigsKept(),sh.numEigsKept(),dspec,norm(rec_h-rec_g),sg.truncerr(),sh.truncerr(),ok ? "ok" : "FAIL");
                if(!ok) ++failures;
                }
            }
        }
    if(failures) { printfln("%d case(s) FAILED",failures); return 1; }
    println("all dense-ops cases ok");
    return 0;
    }
