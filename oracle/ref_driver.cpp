// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Nothing in somar_b200/ may call this.
//
// ref_driver: stands in for AMRNSLevel::projectCorrect (reference
// src/Grade5_SOMAR/AMRNSLevelProject.cpp:247-373) and AMRNSLevel::validateOpsAndSolvers
// (src/Grade5_SOMAR/AMRNSLevelInit.cpp:590-690) so that the reference's OWN, UNMODIFIED
// Elliptic::PoissonOp / MGSolver / BiCGStabSolver / LevelHybridSolver classes (compiled from
// /root/reference by oracle/build_ref.sh) can be run on inputs prepared by the tests.
//
// It only (a) builds the single-level grids the way AnisotropicAMR::makeBaseLevelMesh does
// (src/Grade2_AnisotropicChombo/AnisotropicAMR.cpp:1461-1580), (b) constructs LevelGeometry,
// PoissonOp and LevelHybridSolver with the same arguments AMRNSLevel uses, (c) taps the
// depth-0 operator's virtual norm() so that the per-cycle residual norms are available in full
// precision (the reference only prints 7 digits), and (d) reads / writes flat binary files.
//
// Usage: somar_ref <deck> [key=value ...]
//   drv.mode  = solve | project | applyop | relax | transform | divgrad | vcycle
//   drv.map   = cartesian | stretched        drv.ampl = ax ay az (StretchedMap amplitudes)
//   drv.in    = <file>    raw little-endian doubles, global Fortran order over the domain box:
//                  solve:   rhs[nx*ny*nz]
//                  project: U0[(nx+1)*ny*nz] U1[nx*(ny+1)*nz] U2[nx*ny*(nz+1)]  (advecting velocity)
//                  applyop: phi[nx*ny*nz]
//                  relax:   phi[nx*ny*nz] rhs[nx*ny*nz]
//   drv.out   = <prefix>  writes <prefix>.bin (doubles) and <prefix>.txt (key = value lines)
//   drv.relaxIters = n (relax mode)
//   drv.reps = n          repeat the solve n times for timing (first result is dumped)
//   drv.useMGSolver = 1   call MGSolver directly instead of LevelHybridSolver (always MG mode)
//   drv.refSchedule = r0x r0y r0z r1x ... explicit MG refinement schedule (drv.useMGSolver=1)
#include <chrono>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <vector>

#include "BCTools.H"
#include "CFInterp.H"
#include "CartesianMap.H"
#include "DisjointBoxLayout.H"
#include "FluxBox.H"
#include "LevelData.H"
#include "LevelGeometry.H"
#include "AMRHybridSolver.H"
#include "LevelHybridSolver.H"
#include "MGSolver.H"
#include "ParmParse.H"
#include "PoissonOp.H"
#include "ProblemContext.H"
#include "StretchedMap.H"

using Elliptic::PoissonOp;
typedef LevelData<FArrayBox> LDFAB;

// -DSB_SHIM: the same driver with the two `using` lines of Grade5_SOMAR/AMRNSLevel.H:1448-1450 swapped -- the operator and the
// level solver are the B200 drop-ins of integration/ (C++ API -> C ABI -> CUDA); everything else, including the
// reference's own MGSolver when drv.useMGSolver=1, is the reference's code.  Built as somar_ref_b200.
#ifdef SB_SHIM
#include "B200LevelHybridSolver.H"
typedef somar_b200::B200PoissonOp         BasePoissonOp;
typedef somar_b200::B200LevelHybridSolver LevelSolverType;
#else
typedef Elliptic::PoissonOp         BasePoissonOp;
typedef Elliptic::LevelHybridSolver LevelSolverType;
#endif

static std::vector<double> g_norms;  // every depth-0 norm() result, in call order
static long g_relaxCalls = 0, g_relaxIters = 0, g_residualCalls = 0;

// Depth-0 tap.  MGSolver::define clones its top operator with newMGOperator(Unit)
// (MGSolverI.H:188), so the clone must be a TapOp too.
class TapOp : public BasePoissonOp
{
public:
    using BasePoissonOp::BasePoissonOp;
    TapOp(const TapOp& a_src) : BasePoissonOp(a_src) {}

    Real
    norm(const LDFAB& a_x, const int a_p, const Real a_powScale = 1.0) const override
    {
        const Real v = BasePoissonOp::norm(a_x, a_p, a_powScale);
        g_norms.push_back(v);
        return v;
    }

    void
    relax(LDFAB& a_cor, const LDFAB& a_res, const Real a_time, const int a_iters) const override
    {
        ++g_relaxCalls;
        g_relaxIters += a_iters;
        BasePoissonOp::relax(a_cor, a_res, a_time, a_iters);
    }

    Elliptic::MGOperator<LDFAB>*
    newMGOperator(const IntVect& a_refRatio) const override
    {
        if (a_refRatio == IntVect::Unit) return new TapOp(*this);
        return BasePoissonOp::newMGOperator(a_refRatio);
    }

    const LDFAB&   J() const { return m_J; }
    const LDFAB&   Dinv() const { return m_Dinv; }
    const FArrayBox& M(int d) const { return m_M[d]; }
    bool           hasNullSpace() const { return m_hasNullSpace; }
};


// Tap on the AMR levels' operators: AMRHybridSolver keeps its residual history private and prints 7 digits, so
// every AMRNormLevel result is recorded in call order.
static std::vector<double> g_amrNorms;
// (-DSB_SHIM: the levels' operators are B200PoissonOps, so the reference's own AMRHybridSolver -- and the LevelHybridSolvers it
// builds for its level solves -- drive the device operator through the virtual interface.)
class AmrTapOp : public BasePoissonOp
{
public:
    using BasePoissonOp::BasePoissonOp;
    Real
    AMRNormLevel(const LDFAB& a_res, const LDFAB* a_fineResPtr, const IntVect& a_refRatio, const int a_p) const override
    {
        const Real v = BasePoissonOp::AMRNormLevel(a_res, a_fineResPtr, a_refRatio, a_p);
        g_amrNorms.push_back(v);
        return v;
    }
};


static void
makeBaseGrids(Vector<Box>& a_grids, const ProblemDomain& a_domain, const IntVect& a_maxBGS,
              const IntVect& a_splitDirs, const int a_blockFactor)
{
    // Same arithmetic as AnisotropicAMR::makeBaseLevelMesh.
    const IntVect unsplit = IntVect::Unit - a_splitDirs;
    const IntVect bf      = a_blockFactor * a_splitDirs + unsplit;
    IntVect       maxBGS  = a_maxBGS;
    for (int d = 0; d < SpaceDim; ++d) {
        if (maxBGS[d] == 0 || maxBGS[d] > a_domain.size(d) || a_splitDirs[d] == 0)
            maxBGS[d] = a_domain.size(d);
    }
    Box blk = coarsen(a_domain.domainBox(), bf);
    if (refine(blk, bf) != a_domain.domainBox()) MayDay::Error("domain not coarsenable by blockFactor");
    for (int d = 0; d < SpaceDim; ++d) {
        if (a_splitDirs[d] == 1) continue;
        blk.shift(d, -blk.smallEnd(d));
        blk.setBig(d, blk.smallEnd(d));
    }
    IntVect num, base;
    for (int d = 0; d < SpaceDim; ++d) {
        int       nd = 1;
        const int sz = blk.size(d);
        if (a_splitDirs[d] == 0) {
            num[d] = 1; base[d] = sz;
        } else {
            const int bmax = maxBGS[d] / a_blockFactor;
            if (bmax <= 0) MayDay::Error("maxBaseGridSize < blockFactor");
            while (nd * bmax < sz) ++nd;
            num[d] = nd; base[d] = (sz + nd - 1) / nd;
        }
    }
    const Box b(IntVect::Zero, num - IntVect::Unit);
    for (BoxIterator bit(b); bit.ok(); ++bit) {
        const IntVect slo = a_splitDirs * (blk.smallEnd() + bit() * base);
        const IntVect shi = a_splitDirs * min(slo + base - IntVect::Unit, blk.bigEnd());
        const IntVect lo  = slo + unsplit * a_domain.domainBox().smallEnd();
        const IntVect hi  = shi + unsplit * a_domain.domainBox().bigEnd();
        Box g(lo, hi);
        g.refine(bf);
        a_grids.push_back(g);
    }
}


// global flat array (Fortran order over a_region) <-> LevelData
static void
scatter(LDFAB& a_dst, const double* a_src, const Box& a_region)
{
    const IntVect lo = a_region.smallEnd();
    const IntVect sz = a_region.size();
    for (DataIterator dit = a_dst.dataIterator(); dit.ok(); ++dit) {
        Box bx = a_dst.getBoxes()[dit];
        bx.convert(a_region.type());
        bx &= a_dst[dit].box();
        for (BoxIterator bit(bx); bit.ok(); ++bit) {
            const IntVect iv = bit() - lo;
            size_t idx = iv[0] + sz[0] * (size_t)(iv[1] D_TERM(, , +sz[1] * (size_t)iv[2]));
            a_dst[dit](bit(), 0) = a_src[idx];
        }
    }
}
static void
scatterFAB(FArrayBox& a_dst, const Box& a_valid, const double* a_src, const Box& a_region)
{
    const IntVect lo = a_region.smallEnd();
    const IntVect sz = a_region.size();
    Box bx = a_valid;
    bx.convert(a_region.type());
    bx &= a_dst.box();
    for (BoxIterator bit(bx); bit.ok(); ++bit) {
        const IntVect iv = bit() - lo;
        size_t idx = iv[0] + sz[0] * (size_t)(iv[1] D_TERM(, , +sz[1] * (size_t)iv[2]));
        a_dst(bit(), 0) = a_src[idx];
    }
}
static void
gatherFAB(std::vector<double>& a_dst, const FArrayBox& a_src, const Box& a_valid, const Box& a_region)
{
    const IntVect lo = a_region.smallEnd();
    const IntVect sz = a_region.size();
    Box bx = a_valid;
    bx.convert(a_region.type());
    bx &= a_src.box();
    for (BoxIterator bit(bx); bit.ok(); ++bit) {
        const IntVect iv = bit() - lo;
        size_t idx = iv[0] + sz[0] * (size_t)(iv[1] D_TERM(, , +sz[1] * (size_t)iv[2]));
        a_dst[idx] = a_src(bit(), 0);
    }
}
static std::vector<double>
gather(const LDFAB& a_src, const Box& a_region)
{
    std::vector<double> v(a_region.numPts(), 0.0);
    for (DataIterator dit = a_src.dataIterator(); dit.ok(); ++dit)
        gatherFAB(v, a_src[dit], a_src.getBoxes()[dit], a_region);
    return v;
}

struct OutFile {
    FILE*         bin;
    std::ofstream txt;
    size_t        off = 0;
    OutFile(const std::string& p) : bin(fopen((p + ".bin").c_str(), "wb")), txt(p + ".txt")
    {
        txt << std::setprecision(17) << std::scientific;
    }
    ~OutFile() { fclose(bin); }
    void
    put(const std::string& name, const std::vector<double>& v)
    {
        fwrite(v.data(), sizeof(double), v.size(), bin);
        txt << "array " << name << " " << off << " " << v.size() << "\n";
        off += v.size();
    }
    template <class T>
    void
    kv(const std::string& k, const T& v)
    {
        txt << k << " = " << v << "\n";
    }
};


int
main(int argc, char* argv[])
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s <deck> [key=value ...]\n", argv[0]);
        return 2;
    }
    ParmParse ppRoot(argc - 2, argv + 2, NULL, argv[1]);

    const ProblemContext* ctx = ProblemContext::getInstance();
    ParmParse             drv("drv");

    std::string mode = "solve", mapName = "cartesian", inFile, outPrefix = "ref_out";
    drv.query("mode", mode);
    drv.query("map", mapName);
    drv.query("in", inFile);
    drv.query("out", outPrefix);
    int reps = 1, useMGSolver = 0, relaxIters = 1;
    drv.query("reps", reps);
    drv.query("useMGSolver", useMGSolver);
    drv.query("relaxIters", relaxIters);

    const ProblemDomain& domain = ctx->base.domain;
    const Box            domBox = domain.domainBox();
    const RealVect       L      = ctx->base.L;

    // Geometry (exec/*/UserMain.cpp::oneTimeSetup).
    GeoSourceInterface* geoPtr = nullptr;
    if (mapName == "cartesian") {
        geoPtr = new CartesianMap();
    } else if (mapName == "stretched") {
        Vector<Real> vampl(SpaceDim, 0.0);
        drv.queryarr("ampl", vampl, 0, SpaceDim);
        const RealVect dx   = L / RealVect(ctx->base.nx);
        const RealVect xmin = RealVect(ctx->base.nxOffset) * dx;
        const RealVect xmax = xmin + L;
        geoPtr              = new StretchedMap(xmin, xmax, RealVect(vampl));
    } else {
        MayDay::Error("drv.map must be cartesian or stretched");
    }

    // Grids.
    Vector<Box> boxes;
    makeBaseGrids(boxes, domain, ctx->base.maxBaseGridSize, ctx->base.splitDirs, ctx->base.blockFactor);
    DisjointBoxLayout grids;
    grids.defineAndLoadBalance(boxes, nullptr, domain);

    LevelGeometry levGeo(domain, L, nullptr, geoPtr);
    levGeo.createMetricCache(grids);

    // Operator + solver, as AMRNSLevel::validateOpsAndSolvers does.
    std::shared_ptr<BCTools::BCFunction> bcPtr(new BCTools::HomogNeumBC);
    // drv.alpha / drv.beta: L = J (alpha + beta Lap) (PoissonOp::setAlphaAndBeta, PoissonOp.cpp:707-718); the projector uses 0, 1,
    // the implicit viscous / diffusive solves a Helmholtz form of the same operator
    Real opAlpha = 0.0, opBeta = 1.0;
    drv.query("alpha", opAlpha);
    drv.query("beta", opBeta);
    std::shared_ptr<TapOp> opPtr(new TapOp(levGeo, DisjointBoxLayout(), DisjointBoxLayout(), 1, bcPtr, opAlpha, opBeta, nullptr));

    OutFile out(outPrefix);
    out.kv("spacedim", SpaceDim);
    out.kv("numBoxes", boxes.size());

    {
        LayoutIterator lit = grids.layoutIterator();
        int            n   = 0;
        for (lit.reset(); lit.ok(); ++lit, ++n) {
            const Box& b = grids[lit];
            out.txt << "box " << n;
            for (int d = 0; d < SpaceDim; ++d) out.txt << " " << b.smallEnd(d);
            for (int d = 0; d < SpaceDim; ++d) out.txt << " " << b.bigEnd(d);
            out.txt << "\n";
        }
    }
    out.kv("hasNullSpace", (int)opPtr->hasNullSpace());
    out.kv("relaxMethod", ctx->proj.relaxMethod);

    // Metric dump (lets the tests pin the host-side geometry of the product).
    {
        out.put("J", gather(opPtr->J(), domBox));
        out.put("Dinv", gather(opPtr->Dinv(), domBox));
        for (int d = 0; d < SpaceDim; ++d) {
            const FArrayBox&    M = opPtr->M(d);
            std::vector<double> v(M.box().numPts() * 2);
            for (size_t i = 0; i < v.size(); ++i) v[i] = M.dataPtr(0)[i];
            out.put(std::string("M") + char('0' + d), v);
        }
        for (int d = 0; d < SpaceDim; ++d) {
            Box fcDom = surroundingNodes(domBox, d);
            std::vector<double> v(fcDom.numPts(), 0.0);
            for (DataIterator dit(grids); dit.ok(); ++dit)
                gatherFAB(v, opPtr->getFCJgup()[dit][d], grids[dit], fcDom);
            out.put(std::string("Jgup") + char('0' + d), v);
        }
    }

    // Read input.
    std::vector<double> in;
    if (!inFile.empty()) {
        FILE* f = fopen(inFile.c_str(), "rb");
        if (!f) MayDay::Error("cannot open drv.in");
        fseek(f, 0, SEEK_END);
        const long nb = ftell(f);
        fseek(f, 0, SEEK_SET);
        in.resize(nb / sizeof(double));
        if (fread(in.data(), sizeof(double), in.size(), f) != in.size()) MayDay::Error("short read");
        fclose(f);
    }
    const size_t N = domBox.numPts();

    if (mode == "amr") {
        // Composite solve over 2 or 3 AMR levels: AMRHybridSolver over the base level and
        // one refined rectangular patch per finer level, set up the way AMRNSLevel::validateOpsAndSolvers does
        // (Grade5_SOMAR/AMRNSLevelInit.cpp:617-676).
        //   drv.refRatio    = r0 r1 [r2]          refinement ratio of level 1
        //   drv.fineRegion  = lo... hi...          level-0-index box that level 1 covers
        //   drv.fineMaxBox  = n                     level-1 boxes are cut to at most n cells per direction (0: one box)
        //   drv.refRatio2 / drv.fineRegion2 (level-1 indices) / drv.fineMaxBox2: an optional level 2, likewise
        //   drv.in: rhs on level 0 over the domain box, then rhs on level l over refine(fineRegion_l), l = 1, 2
        std::vector<IntVect> vref;            // vref[l]: ratio between level l-1 and l
        std::vector<Box>     vregion;         // in level l index space
        std::vector<int>     vmaxBox;
        vref.push_back(IntVect::Unit); vregion.push_back(domBox); vmaxBox.push_back(0);
        for (int l = 1; l <= 2; ++l) {
            const std::string sfx = l == 1 ? "" : "2";
            Vector<int> vr(SpaceDim, 2), vreg;
            if (l == 2 && !drv.contains("fineRegion2")) break;
            drv.queryarr(("refRatio" + sfx).c_str(), vr, 0, SpaceDim);
            drv.getarr(("fineRegion" + sfx).c_str(), vreg, 0, 2 * SpaceDim);
            int fmb = 0;
            drv.query(("fineMaxBox" + sfx).c_str(), fmb);
            const IntVect ref(D_DECL(vr[0], vr[1], vr[2]));
            const Box     crseRegion(IntVect(D_DECL(vreg[0], vreg[1], vreg[2])),
                                     IntVect(D_DECL(vreg[SpaceDim], vreg[SpaceDim + 1], vreg[SpaceDim + 2])));
            if (!vregion.back().contains(crseRegion)) MayDay::Error("drv.fineRegion must lie inside the coarser level");
            vref.push_back(ref);
            vregion.push_back(refine(crseRegion, ref));
            vmaxBox.push_back(fmb);
        }
        const int nlev = (int)vref.size();
        std::vector<ProblemDomain>                  vdom(nlev);
        std::vector<DisjointBoxLayout>              vgrids(nlev);
        std::vector<std::shared_ptr<LevelGeometry>> vgeo(nlev);
        vdom[0] = domain; vgrids[0] = grids;
        std::vector<size_t> vnumBoxes(nlev, boxes.size());
        for (int l = 1; l < nlev; ++l) {
            vdom[l] = vdom[l - 1];
            vdom[l].refine(vref[l]);
            const Box&  fineRegion = vregion[l];
            Vector<Box> fineBoxes;
            IntVect     nb, sz;
            for (int d = 0; d < SpaceDim; ++d) {
                const int n = fineRegion.size(d);
                nb[d]       = (vmaxBox[l] > 0) ? (n + vmaxBox[l] - 1) / vmaxBox[l] : 1;
                if (n % nb[d]) MayDay::Error("drv.fineMaxBox must divide the refined region evenly");
                sz[d] = n / nb[d];
            }
            for (BoxIterator bit(Box(IntVect::Zero, nb - IntVect::Unit)); bit.ok(); ++bit) {
                const IntVect lo = fineRegion.smallEnd() + bit() * sz;
                fineBoxes.push_back(Box(lo, lo + sz - IntVect::Unit));
            }
            vgrids[l].defineAndLoadBalance(fineBoxes, nullptr, vdom[l]);
            vnumBoxes[l] = fineBoxes.size();
        }
        for (int l = 1; l < nlev; ++l) {
            vgeo[l].reset(new LevelGeometry(vdom[l], L, l == 1 ? &levGeo : vgeo[l - 1].get(), geoPtr));
            vgeo[l]->createMetricCache(vgrids[l]);
        }
        using OpType = Elliptic::AMRMGOperator<LDFAB>;
        Vector<std::shared_ptr<const OpType>> vOps(nlev);
        for (int l = 0; l < nlev; ++l) {
            const LevelGeometry& lg = l == 0 ? levGeo : *vgeo[l];
            vOps[l].reset(new AmrTapOp(lg, l + 1 < nlev ? vgrids[l + 1] : DisjointBoxLayout(), l > 0 ? vgrids[l - 1] : DisjointBoxLayout(), 1,
                                       bcPtr));
        }
        // level data
        std::vector<std::shared_ptr<LDFAB>> phi(nlev), rhs(nlev);
        std::vector<size_t>                 off(nlev + 1, 0);
        for (int l = 0; l < nlev; ++l) {
            phi[l].reset(new LDFAB(vgrids[l], 1, IntVect::Unit));
            rhs[l].reset(new LDFAB(vgrids[l], 1));
            for (DataIterator dit(vgrids[l]); dit.ok(); ++dit) { (*phi[l])[dit].setVal(0.0); (*rhs[l])[dit].setVal(0.0); }
            off[l + 1] = off[l] + vregion[l].numPts();
        }
        int compProjectEarly = 0;
        drv.query("compProject", compProjectEarly);
        if (!compProjectEarly && in.size() < off[nlev]) MayDay::Error("drv.in too short for the level data");
        auto tag = [&](const char* base, int l) { return std::string(base) + char('0' + l); };

        int cfOnly = 0;
        drv.query("cfInterpOnly", cfOnly);
        if (cfOnly) {
            // Known-answer hook for the coarse-fine ghost interpolation (PoissonOp::applyBCs with a coarse level ->
            // CFInterp::interpAtCFI -> MappedQuadCFInterp): drv.in = phi on every level; output = the finest level's
            // phi with its ghost layer ("fineWithGhosts"), and the same for level 1 of 3 ("midWithGhosts").
            for (int l = 0; l < nlev; ++l) scatter(*phi[l], in.data() + off[l], vregion[l]);
            for (int l = 1; l < nlev; ++l) {
                vOps[l]->applyBCs(*phi[l], &*phi[l - 1], 0.0, true, false);
                Box gb = vregion[l];
                gb.grow(1);
                std::vector<double> v(gb.numPts(), 0.0);
                // face ghosts only (edge and corner ghosts are not filled by the CF interpolation)
                for (DataIterator dit(vgrids[l]); dit.ok(); ++dit)
                    for (int d = 0; d < SpaceDim; ++d) {
                        Box b = vgrids[l][dit];
                        b.grow(d, 1);
                        gatherFAB(v, (*phi[l])[dit], b, gb);
                    }
                // valid data last, so that a ghost of one box never hides the valid value of its neighbour
                for (DataIterator dit(vgrids[l]); dit.ok(); ++dit) gatherFAB(v, (*phi[l])[dit], vgrids[l][dit], gb);
                out.put(l == nlev - 1 ? "fineWithGhosts" : "midWithGhosts", v);
            }
            return 0;
        }

        // Composite residual through the operators' public AMR interface (AMRMGOperator.H:107-175) and its norm
        // (AMRNormLevel, PoissonOp.cpp:1215-1286): an evaluation that does not go through AMRHybridSolver's bookkeeping.
        auto compResidual = [&](std::vector<std::shared_ptr<LDFAB>>& f, const std::string& name, bool dumpFields) {
            std::vector<std::shared_ptr<LDFAB>> res(nlev);
            for (int l = 0; l < nlev; ++l) res[l].reset(new LDFAB(vgrids[l], 1));
            for (int l = nlev - 1; l >= 0; --l) {
                if (l == nlev - 1 && l > 0) vOps[l]->AMRResidualNF(*res[l], *f[l], *f[l - 1], *rhs[l], vref[l], 0.0, true);
                else if (l == 0) vOps[l]->AMRResidualNC(*res[l], *f[l + 1], *f[l], *rhs[l], vref[l + 1], 0.0, true, *vOps[l + 1]);
                else vOps[l]->AMRResidual(*res[l], *f[l + 1], *f[l], *f[l - 1], *rhs[l], vref[l + 1], vref[l], 0.0, true, *vOps[l + 1]);
            }
            for (int l = 0; l < nlev; ++l) {
                const Real n = l == nlev - 1 ? vOps[l]->AMRNormLevel(*res[l], nullptr, IntVect::Unit, ctx->proj.normType)
                                             : vOps[l]->AMRNormLevel(*res[l], &*res[l + 1], vref[l + 1], ctx->proj.normType);
                out.kv(name + "Norm" + char('0' + l), n);
                if (dumpFields) out.put(name + char('0' + l), gather(*res[l], vregion[l]));
            }
        };

        int applyOnly = 0;
        drv.query("applyOnly", applyOnly);
        if (applyOnly) {
            // Known-answer hook for the composite operator: drv.in holds phi on every level (not a right-hand side);
            // out = -L[phi] per level with a zero right-hand side, i.e. the level operator with inhomogeneous
            // coarse-fine ghosts and / or the refluxed operator (PoissonOp.cpp:1156-1207, 1296-1440).
            for (int l = 0; l < nlev; ++l) scatter(*phi[l], in.data() + off[l], vregion[l]);
            compResidual(phi, "minusL", true);
            g_amrNorms.clear();
            return 0;
        }
        int compProject = 0;
        drv.query("compProject", compProject);
        if (compProject) {
            // The composite (sync) projection, AMRNSLevel::projectDownToThis (Grade5_SOMAR/AMRNSLevelProject.cpp:740-860),
            // restated around the reference's own operator calls (Grade5 is not part of this build): compDivergence with
            // refluxing per level, averageDown of the right-hand side (AMRNSLevelUtil.cpp:748-781 -> CFInterp::coarsen),
            // AMRHybridSolver::solve, compGradient with coarse-level CF BCs, vel -= grad, final compDivergence.
            //   drv.in: per level, the D face fields of the advecting velocity over the level's region.
            std::vector<std::shared_ptr<LevelData<FluxBox>>> vel(nlev), grad(nlev);
            size_t                                           o = 0;
            for (int l = 0; l < nlev; ++l) {
                vel[l].reset(new LevelData<FluxBox>(vgrids[l], 1));
                grad[l].reset(new LevelData<FluxBox>(vgrids[l], 1));
                for (int d = 0; d < SpaceDim; ++d) {
                    const Box fc = surroundingNodes(vregion[l], d);
                    if (in.size() < o + fc.numPts()) MayDay::Error("drv.in too short for the level velocities");
                    for (DataIterator dit(vgrids[l]); dit.ok(); ++dit) scatterFAB((*vel[l])[dit][d], vgrids[l][dit], in.data() + o, fc);
                    o += fc.numPts();
                }
            }
            std::vector<const PoissonOp*> pops(nlev);
            for (int l = 0; l < nlev; ++l) pops[l] = static_cast<const PoissonOp*>(&*vOps[l]);
            auto compDiv = [&](const char* name) {
                for (int l = 0; l < nlev; ++l) pops[l]->compDivergence(*rhs[l], *vel[l], l + 1 < nlev ? &*vel[l + 1] : nullptr);
                for (int l = 0; l < nlev; ++l) out.put(tag(name, l), gather(*rhs[l], vregion[l]));
            };
            compDiv("div_init");
            for (int l = nlev - 1; l > 0; --l) {
                CFInterp interp;
                interp.define(vgrids[l], (l == 0 ? levGeo : *vgeo[l]).getDXi(), vgrids[l - 1]);
                interp.coarsen(*rhs[l - 1], *rhs[l], false, nullptr);
            }
            for (int l = 0; l < nlev; ++l) out.put(tag("rhs", l), gather(*rhs[l], vregion[l]));
            Elliptic::AMRHybridSolver amr;
            amr.define(vOps, 0, nlev - 1, Elliptic::AMRHybridSolver::getDefaultOptions());
            Vector<LDFAB*>       vphi(nlev);
            Vector<const LDFAB*> vrhs(nlev);
            for (int l = 0; l < nlev; ++l) { vphi[l] = &*phi[l]; vrhs[l] = &*rhs[l]; }
            g_amrNorms.clear();
            Elliptic::SolverStatus st = amr.solve(vphi, vrhs, 0.0, true, true);
            out.put("amrLevelNorms", g_amrNorms);
            out.kv("status", st.getSolverStatus());
            for (int l = 0; l < nlev; ++l) {
                pops[l]->compGradient(*grad[l], *phi[l], l > 0 ? &*phi[l - 1] : nullptr, 0.0, true, false);
                for (DataIterator dit(vgrids[l]); dit.ok(); ++dit)
                    for (int d = 0; d < SpaceDim; ++d) (*vel[l])[dit][d].plus((*grad[l])[dit][d], -1.0);
            }
            for (int l = 0; l < nlev; ++l) {
                out.put(tag("phi", l), gather(*phi[l], vregion[l]));
                for (int d = 0; d < SpaceDim; ++d) {
                    const Box           fc = surroundingNodes(vregion[l], d);
                    std::vector<double> v(fc.numPts(), 0.0);
                    for (DataIterator dit(vgrids[l]); dit.ok(); ++dit) gatherFAB(v, (*vel[l])[dit][d], vgrids[l][dit], fc);
                    out.put(std::string("vel") + char('0' + l) + "_" + char('0' + d), v);
                }
            }
            compDiv("div_final");
            out.kv("numLevels", nlev);
            return 0;
        }
        for (int l = 0; l < nlev; ++l) scatter(*rhs[l], in.data() + off[l], vregion[l]);
        Elliptic::AMRHybridSolver amr;
        amr.define(vOps, 0, nlev - 1, Elliptic::AMRHybridSolver::getDefaultOptions());
        Vector<LDFAB*>       vphi(nlev);
        Vector<const LDFAB*> vrhs(nlev);
        for (int l = 0; l < nlev; ++l) { vphi[l] = &*phi[l]; vrhs[l] = &*rhs[l]; }
        g_amrNorms.clear();
        const auto t0 = std::chrono::high_resolution_clock::now();
        Elliptic::SolverStatus st = amr.solve(vphi, vrhs, 0.0, true, true);
        const auto t1 = std::chrono::high_resolution_clock::now();
        // AMRHybridSolver::computeAMRResidual calls AMRNormLevel once per level, lmin..lmax, per evaluation: the tap
        // holds the per-level norms of the initial residual and of every iteration (the solver keeps its history private)
        out.put("amrLevelNorms", g_amrNorms);
        for (int l = 0; l < nlev; ++l) out.put(tag("phi", l), gather(*phi[l], vregion[l]));
        {
            std::vector<std::shared_ptr<LDFAB>> zero(nlev);
            for (int l = 0; l < nlev; ++l) {
                zero[l].reset(new LDFAB(vgrids[l], 1, IntVect::Unit));
                for (DataIterator dit(vgrids[l]); dit.ok(); ++dit) (*zero[l])[dit].setVal(0.0);
            }
            compResidual(zero, "res_init", true);
            compResidual(phi, "res_final", true);
        }
        out.kv("status", st.getSolverStatus());
        out.kv("initResNorm", st.getInitResNorm());
        out.kv("finalResNorm", st.getFinalResNorm());
        out.kv("solveTime", std::chrono::duration<double>(t1 - t0).count());
        out.kv("numLevels", nlev);
        out.kv("numFineBoxes", vnumBoxes[nlev - 1]);
        return 0;
    }

    if (mode == "applyop") {
        LDFAB phi(grids, 1, IntVect::Unit), lhs(grids, 1);
        for (DataIterator dit(grids); dit.ok(); ++dit) phi[dit].setVal(0.0);
        scatter(phi, in.data(), domBox);
        opPtr->applyOp(lhs, phi, nullptr, 0.0, true, true);
        out.put("lhs", gather(lhs, domBox));
        g_norms.clear();
        out.kv("norm2", opPtr->norm(lhs, 2));
        out.kv("norm0", opPtr->norm(lhs, 0));
        return 0;
    }

    if (mode == "transform") {
        // AMRNSLevel::sendToAdvectingVelocity / sendToCartesianVelocity (Grade5_SOMAR/AMRNSLevelFill.cpp:
        // 194-280) -- Grade5 is not part of this build, so the two loops are restated here around the
        // reference's own GeoSourceInterface::fill_dxdXi and FArrayBox::mult / divide, statement for
        // statement.  drv.velGhost = ghost width of the FluxBox (AMRNSLevel's velocity carries 1).
        int velGhost = 1;
        drv.query("velGhost", velGhost);
        LevelData<FluxBox> vel(grids, 1, velGhost * IntVect::Unit);
        size_t off = 0;
        for (int d = 0; d < SpaceDim; ++d) {
            const Box fcDom = surroundingNodes(domBox, d);
            for (DataIterator dit(grids); dit.ok(); ++dit) {
                vel[dit][d].setVal(0.0);
                scatterFAB(vel[dit][d], grids[dit], in.data() + off, fcDom);
            }
            off += fcDom.numPts();
        }
        const GeoSourceInterface& geoSrc = levGeo.getGeoSource();
        const RealVect&           dXi    = levGeo.getDXi();
        auto dump = [&](const char* tag) {
            for (int d = 0; d < SpaceDim; ++d) {
                const Box           fcDom = surroundingNodes(domBox, d);
                std::vector<double> v(fcDom.numPts(), 0.0);
                for (DataIterator dit(grids); dit.ok(); ++dit) gatherFAB(v, vel[dit][d], grids[dit], fcDom);
                out.put(std::string(tag) + char('0' + d), v);
            }
        };
        for (DataIterator dit(grids); dit.ok(); ++dit) {
            for (int fcDir = 0; fcDir < SpaceDim; ++fcDir) {
                FArrayBox& velFAB = vel[dit][fcDir];
                FArrayBox  dxdXiFAB(velFAB.box(), 1);
                for (int offset = 1; offset < SpaceDim; ++offset) {
                    const int mu = (fcDir + offset) % SpaceDim;
                    geoSrc.fill_dxdXi(dxdXiFAB, 0, mu, dXi);
                    velFAB.mult(dxdXiFAB, 0, 0, 1);
                }
            }
        }
        dump("adv");
        for (DataIterator dit(grids); dit.ok(); ++dit) {
            for (int fcDir = 0; fcDir < SpaceDim; ++fcDir) {
                FArrayBox& velFAB = vel[dit][fcDir];
                FArrayBox  dxdXiFAB(velFAB.box(), 1);
                for (int offset = 1; offset < SpaceDim; ++offset) {
                    const int mu = (fcDir + offset) % SpaceDim;
                    geoSrc.fill_dxdXi(dxdXiFAB, 0, mu, dXi);
                    velFAB.divide(dxdXiFAB, 0, 0, 1);
                }
            }
        }
        dump("cart");
        return 0;
    }

    if (mode == "divgrad") {
        // Known-answer hooks for the face-centred operators: drv.in = phi over the domain box, then the D
        // face fields of an advecting velocity; out = levelGradient(phi) and levelDivergence(vel)
        // (PoissonOp.cpp:1486-1545, 1568-1610).
        LDFAB phi(grids, 1, IntVect::Unit), div(grids, 1);
        for (DataIterator dit(grids); dit.ok(); ++dit) phi[dit].setVal(0.0);
        scatter(phi, in.data(), domBox);
        LevelData<FluxBox> vel(grids, 1), grad(grids, 1);
        size_t off = N;
        for (int d = 0; d < SpaceDim; ++d) {
            const Box fcDom = surroundingNodes(domBox, d);
            for (DataIterator dit(grids); dit.ok(); ++dit) scatterFAB(vel[dit][d], grids[dit], in.data() + off, fcDom);
            off += fcDom.numPts();
        }
        opPtr->levelGradient(grad, phi, nullptr, 0.0, true, true);
        opPtr->levelDivergence(div, vel);
        out.put("div", gather(div, domBox));
        for (int d = 0; d < SpaceDim; ++d) {
            const Box           fcDom = surroundingNodes(domBox, d);
            std::vector<double> g(fcDom.numPts(), 0.0);
            for (DataIterator dit(grids); dit.ok(); ++dit) gatherFAB(g, grad[dit][d], grids[dit], fcDom);
            out.put(std::string("grad") + char('0' + d), g);
        }
        return 0;
    }

    if (mode == "relax") {
        LDFAB phi(grids, 1, IntVect::Unit), rhs(grids, 1);
        for (DataIterator dit(grids); dit.ok(); ++dit) phi[dit].setVal(0.0);
        scatter(phi, in.data(), domBox);
        scatter(rhs, in.data() + N, domBox);
        opPtr->relax(phi, rhs, 0.0, relaxIters);
        out.put("phi", gather(phi, domBox));
        return 0;
    }

    // Solver.
    // LevelHybridSolver keeps its residual-norm history (initial, then one entry per leptic order /
    // V-cycle, LevelHybridSolver.cpp:312-400) in a protected member; expose it.
    struct HybridPeek : LevelSolverType {
        const std::vector<Real>& resNorms() const { return m_resNorms; }
    } hybrid;
    Elliptic::MGSolver<LDFAB>            mg;
    Elliptic::LevelHybridSolver::Options hopt = Elliptic::LevelHybridSolver::getDefaultOptions();
    if (useMGSolver) {
        Vector<IntVect> sched;
        Vector<int>     vs;
        if (drv.contains("refSchedule")) {
            drv.queryarr("refSchedule", vs, 0, drv.countval("refSchedule"));
            for (size_t i = 0; i + SpaceDim <= vs.size(); i += SpaceDim)
                sched.push_back(IntVect(D_DECL(vs[i], vs[i + 1], vs[i + 2])));
        }
        mg.define(*opPtr, hopt.mgOptions, sched);
        const auto& o = mg.getOptions();
        out.kv("maxDepth", o.maxDepth);
    } else {
        hybrid.define(opPtr, hopt);
        struct Peek : Elliptic::LevelHybridSolver { using Elliptic::LevelHybridSolver::computeSolveMode; };
        std::shared_ptr<const Elliptic::MGOperator<LDFAB>> mgOpPtr = opPtr;
        out.kv("solveMode", (int)Peek::computeSolveMode(mgOpPtr, hopt));  // 1 MG, 2 Leptic, 3 Leptic_MG
#ifndef SB_SHIM
        out.kv("maxDepth", hybrid.getOptions().mgOptions.maxDepth);
#endif
    }

    LDFAB rhs(grids, 1), phi(grids, 1, IntVect::Unit);
    LevelData<FluxBox> vel, gradPhi;
    double initDivNorm = 0.0;

    if (mode == "vcycle") {
        // One outer-iteration body of MGSolver::vCycle (MGSolverI.H:342-347): preCond, then
        // vCycle_residualEq(cor, res, depth 0) -- the unit of the headline metric.  Timed with
        // std::chrono like the reference's own "Solve time" (AMRNSLevelProject.cpp:315-324).
        if (!useMGSolver) MayDay::Error("drv.mode=vcycle needs drv.useMGSolver=1");
        if (in.size() >= N) {
            scatter(rhs, in.data(), domBox);
        } else {
            // synthetic residual: smooth + oscillatory, zero mean is not required here
            for (DataIterator dit(grids); dit.ok(); ++dit)
                for (BoxIterator bit(grids[dit]); bit.ok(); ++bit) {
                    const IntVect& iv = bit();
                    rhs[dit](iv, 0) = std::sin(0.37 * iv[0]) * std::cos(0.23 * iv[1]) + 0.1 * ((iv[0] * 7 + iv[1] * 13 + iv[SpaceDim - 1] * 29) % 17 - 8);
                }
        }
        std::vector<double> times;
        for (int rep = 0; rep < reps; ++rep) {
            opPtr->preCond(phi, rhs, 0.0, 0);
            const auto t0 = std::chrono::high_resolution_clock::now();
            mg.vCycle_residualEq(phi, rhs, 0.0, 0);
            const auto t1 = std::chrono::high_resolution_clock::now();
            times.push_back(std::chrono::duration<double>(t1 - t0).count());
        }
        out.put("phi", gather(phi, domBox));
        out.put("vcycleTimes", times);
        out.kv("numCells", (double)N);
        return 0;
    }

    if (mode == "predict") {
        // AMRNSLevel::projectPredict (Grade5_SOMAR/AMRNSLevelProject.cpp:63-243) between the level's own BC fills, restated
        // around the reference's operator and solver calls (Grade5 is not part of this build): lagged-pressure correction,
        // and, if it raised |Div[vel]|, one projectCorrect (:247-373) under LevelHybridSolver's quick-and-dirty options.
        //   drv.in: the D face fields of the advecting velocity, then p over the domain box;  drv.projDt
        if (useMGSolver) MayDay::Error("drv.mode=predict uses the LevelHybridSolver");
        Real projDt = 0.5;
        drv.query("projDt", projDt);
        vel.define(grids, 1);
        gradPhi.define(grids, 1);
        LDFAB  p(grids, 1, IntVect::Unit), divVel(grids, 1);
        size_t off = 0;
        for (int d = 0; d < SpaceDim; ++d) {
            const Box fcDom = surroundingNodes(domBox, d);
            for (DataIterator dit(grids); dit.ok(); ++dit) scatterFAB(vel[dit][d], grids[dit], in.data() + off, fcDom);
            off += fcDom.numPts();
        }
        for (DataIterator dit(grids); dit.ok(); ++dit) p[dit].setVal(0.0);
        scatter(p, in.data() + off, domBox);
        const int normType = ctx->proj.normType;
        opPtr->levelDivergence(divVel, vel);
        const Real n0 = opPtr->BasePoissonOp::norm(divVel, normType);
        opPtr->levelGradient(gradPhi, p, nullptr, 0.0, false, false);
        for (DataIterator dit(grids); dit.ok(); ++dit)
            for (int d = 0; d < SpaceDim; ++d) vel[dit][d].plus(gradPhi[dit][d], -projDt);
        opPtr->levelDivergence(divVel, vel);
        const Real n1 = opPtr->BasePoissonOp::norm(divVel, normType);
        Real       n2 = -1.0;
        int        fallback = 0, fbStatus = -1;
        if (n1 > n0) {
            fallback = 1;
            const auto saveOpts = hybrid.getOptions();
            hybrid.modifyOptionsExceptMaxDepth(LevelSolverType::getQuickAndDirtyOptions());
            {   // projectCorrect
                LDFAB phiC(grids, 1, IntVect::Unit);
                opPtr->levelDivergence(divVel, vel);
                Elliptic::SolverStatus st = hybrid.solve(phiC, nullptr, divVel, 0.0, true, true);
                fbStatus = st.getSolverStatus();
                opPtr->levelGradient(gradPhi, phiC, nullptr, 0.0, true, true);
                for (DataIterator dit(grids); dit.ok(); ++dit) {
                    for (int d = 0; d < SpaceDim; ++d) vel[dit][d].plus(gradPhi[dit][d], -1.0);
                    p[dit].plus(phiC[dit], 1.0 / projDt);
                }
            }
            hybrid.modifyOptionsExceptMaxDepth(saveOpts);
            opPtr->levelDivergence(divVel, vel);
            n2 = opPtr->BasePoissonOp::norm(divVel, normType);
        }
        for (int d = 0; d < SpaceDim; ++d) {
            const Box           fcDom = surroundingNodes(domBox, d);
            std::vector<double> v(fcDom.numPts(), 0.0);
            for (DataIterator dit(grids); dit.ok(); ++dit) gatherFAB(v, vel[dit][d], grids[dit], fcDom);
            out.put(std::string("vel") + char('0' + d), v);
        }
        out.put("p", gather(p, domBox));
        out.kv("initDivNorm", n0);
        out.kv("laggedDivNorm", n1);
        out.kv("correctedDivNorm", n2);
        out.kv("usedFallback", fallback);
        out.kv("status", fbStatus);
        return 0;
    }

    if (mode == "project") {
        vel.define(grids, 1);
        gradPhi.define(grids, 1);
        size_t off = 0;
        for (int d = 0; d < SpaceDim; ++d) {
            const Box fcDom = surroundingNodes(domBox, d);
            for (DataIterator dit(grids); dit.ok(); ++dit)
                scatterFAB(vel[dit][d], grids[dit], in.data() + off, fcDom);
            off += fcDom.numPts();
        }
        opPtr->levelDivergence(rhs, vel);
        initDivNorm = opPtr->BasePoissonOp::norm(rhs, ctx->proj.normType);
        out.put("div", gather(rhs, domBox));
        out.kv("initDivNorm", initDivNorm);
    } else if (mode == "solve") {
        scatter(rhs, in.data(), domBox);
    } else {
        MayDay::Error("unknown drv.mode");
    }

    Elliptic::SolverStatus status;
    std::vector<double>    times;
    std::vector<double>    firstNorms;
    for (int rep = 0; rep < reps; ++rep) {
        g_norms.clear();
        g_relaxCalls = g_relaxIters = 0;
        const auto t0 = std::chrono::high_resolution_clock::now();
        if (useMGSolver) {
            status = mg.solve(phi, nullptr, rhs, 0.0, true, true);
        } else {
            status = hybrid.solve(phi, nullptr, rhs, 0.0, true, true);
        }
        const auto t1 = std::chrono::high_resolution_clock::now();
        times.push_back(std::chrono::duration<double>(t1 - t0).count());
        if (rep == 0) firstNorms = g_norms;
    }
    out.put("phi", gather(phi, domBox));
    out.put("norms", firstNorms);
    if (!useMGSolver) out.put("hybridNorms", hybrid.resNorms());
    out.put("solveTimes", times);
    out.kv("status", status.getSolverStatus());
    out.kv("initResNorm", status.getInitResNorm());
    out.kv("finalResNorm", status.getFinalResNorm());
#ifdef SB_SHIM
    if (!useMGSolver) { out.kv("maxDepth", hybrid.maxDepth()); out.kv("deviceSolveMs", hybrid.deviceMilliseconds()); }
    out.kv("impl", "b200-shim");
#endif
    out.kv("relaxCallsDepth0", g_relaxCalls);
    out.kv("relaxItersDepth0", g_relaxIters);

    if (mode == "project") {
        // AMRNSLevelProject.cpp:326-337, 352-357.
        opPtr->levelGradient(gradPhi, phi, nullptr, 0.0, true, true);
        for (DataIterator dit(grids); dit.ok(); ++dit)
            for (int d = 0; d < SpaceDim; ++d) vel[dit][d].plus(gradPhi[dit][d], -1.0);
        for (int d = 0; d < SpaceDim; ++d) {
            const Box           fcDom = surroundingNodes(domBox, d);
            std::vector<double> v(fcDom.numPts(), 0.0), g(fcDom.numPts(), 0.0);
            for (DataIterator dit(grids); dit.ok(); ++dit) {
                gatherFAB(v, vel[dit][d], grids[dit], fcDom);
                gatherFAB(g, gradPhi[dit][d], grids[dit], fcDom);
            }
            out.put(std::string("vel") + char('0' + d), v);
            out.put(std::string("grad") + char('0' + d), g);
        }
        opPtr->levelDivergence(rhs, vel);
        out.kv("finalDivNorm", opPtr->BasePoissonOp::norm(rhs, ctx->proj.normType));
        out.put("divAfter", gather(rhs, domBox));
    }
    return 0;
}
