// sb_host.h -- host-side mirror of the reference's operator / solver classes for the
// projection path.  Names follow the reference (PoissonOp, MGSolver, BiCGStabSolver,
// LevelHybridSolver); each method cites the code it mirrors.
#pragma once
#include <array>
#include <map>

#include "sb_core.h"
#include "sb_halo.h"
#include "sb_line_tma.h"

namespace sb {

struct Comm;  // NCCL wrapper (sb_comm.cpp)

struct Context {
    int          device = 0, rank = 0, nranks = 1;
    cudaStream_t st     = nullptr;
    cudaStream_t commSt = nullptr;          // halo exchanges that overlap interior work (multi-rank line relaxation)
    cudaEvent_t  evEdge = nullptr, evHalo = nullptr;
    cudaStream_t copySt[2] = {nullptr, nullptr};  // SB_STREAM_H2D, SB_STREAM_D2H (asynchronous uploads / downloads)
    cudaStream_t stream(int which);               // SB_STREAM_*; the copy streams are created on first use
    int*         fault  = nullptr;  // pinned, mapped: a kernel watchdog's last words (sb_line_tma.cu), readable after a failed launch
    double*      hpin   = nullptr;  // pinned host scratch (scalars coming back from reductions)
    size_t       hpinLen = 0;
    void*        scratch = nullptr;  // device scratch shared by all depths (line-relax workspace)
    size_t       scratchBytes = 0;
    Comm*        comm = nullptr;
    bool         peerHalo = true;   // SB_PEER_HALO=0 at creation: keep the NCCL face exchange inside the relaxations
    cudaEvent_t  evPost = nullptr;  // last peer post of a pass loop (joins commSt back into st)
    Context*     parent = nullptr;  // set on the single-rank view used by agglomerated MG depths: stream and profile are the parent's
    long long    launches0 = 0;
    // optional per-kernel timing (CUDA events on `st` around selected launches)
    bool         profiling = false;
    bool         phases = false;     // coarse profile: one record per phase of a V-cycle (relax / residual+restrict / prolong / agglomerated
                                     // depths), production launch path (CUDA graphs stay on)
    struct ProfRec { std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev; double ms = 0; long long count = 0; };
    std::map<std::string, ProfRec> prof;
    cudaEvent_t  tm0 = nullptr, tm1 = nullptr;  // sb_context_timer_start/stop
    void profBegin(const char* key, int depth, cudaEvent_t* e0);
    void profEnd(const char* key, int depth, cudaEvent_t e0);
    void profResolve();
    void phaseBegin(cudaEvent_t* e0);
    void phaseEnd(const char* key, int depth, cudaEvent_t e0);
    bool isProfiling() const { return parent ? parent->profiling : profiling; }

    Context(int dev, int rank, int nranks);
    explicit Context(Context& parent);  // single-rank view on the parent's device and stream
    ~Context();
    void* getScratch(size_t bytes);
    void  sync();
    // Comm::reduce (BaseTools/Comm.cpp:14-50)
    void allreduceSum(double* v, int n);
    void allreduceMax(double* v, int n);
};

struct MapSpec {
    int       kind = SB_MAP_CARTESIAN;
    double    xmin[3] = {0, 0, 0}, xmax[3] = {1, 1, 1}, ampl[3] = {0, 0, 0};
    sb_map_fn fn = nullptr;
    void*     user = nullptr;
    // GeoSourceInterface::interp
    void interp(std::vector<double>& x, const std::vector<double>& xi, int mu) const;
    // GeoSourceInterface::fill_physCoor(Vector<Real>&, mu, dXi, box): n points starting at index
    // lo, cell- (type 0) or node-centred (type 1), xi accumulated by repeated addition.
    std::vector<double> physCoor(int mu, double dXi, int lo, int n, int nodeType) const;
    // fill_dxdXi over n points starting at lo of the given centring (GeoSourceInterface.cpp:166-199).
    std::vector<double> dxdXi(int mu, double dXi, int lo, int n, int nodeType, double scale = 1.0) const;
};

struct Op;
struct CFLink;  // sb_amr.cpp

// A DisjointBoxLayout of another AMR level, as PoissonOp's constructor receives it (PoissonOp.cpp:33-43).
struct LevelGrids {
    std::vector<Box3> boxes;
    std::vector<int>  rank;
    Box3              domain{{0, 0, 0}, {-1, -1, -1}};
    bool closed() const { return !boxes.empty(); }
};

// Chombo's Copier at tile granularity (BoxTools/Copier.cpp): regions of a source array that land in a destination
// array, possibly on another rank and on another AMR level.  Every rank builds the same item list.
struct LevelCopier {
    struct Item { int srcRank, dstRank; Box3 box; size_t off; };
    std::vector<Item> items;
    size_t            sendLen = 0, recvLen = 0;  // staging (doubles) this rank needs
    double *          sendBuf = nullptr, *recvBuf = nullptr;
    // src[r] / dst[r]: the boxes (global indices) rank r holds / wants
    void define(int rank, const std::vector<std::vector<Box3>>& src, const std::vector<std::vector<Box3>>& dst);
    // dst(box) = src(box) (mode 0) or dst(box) += scale * src(box) (mode 1, the AddOp of AnisotropicFluxRegister.cpp:577-620)
    void exec(Context* ctx, const Lay& srcLay, const double* src, const Lay& dstLay, double* dst, int mode = 0, double scale = 1.0);
    ~LevelCopier();
};

struct Field {
    Op*     op;
    int     centering;  // SB_CELL or face dir
    double* d = nullptr;
    cudaEvent_t ready = nullptr;  // recorded on the compute stream after the zero fill of alloc(): the copy streams wait on it
    Field(Op* op, int centering);
    ~Field();
    Field(const Field&) = delete;
};

// Elliptic::PoissonOp at one MG depth (Grade3_Calculus/Elliptic/PoissonOp.{H,cpp}).
struct Op {
    Context* ctx;
    int      dim;
    Box3     domain;
    int      periodic[3];
    double   dXi[3];
    double   alpha, beta;
    int      relaxMethod;
    MapSpec  map;
    double   bcAlpha[3][2], bcBeta[3][2];
    int      depth = 0;
    // AMR (PoissonOp.cpp:83-120): the grids of the next coarser level, if any.  A refined level's boxes form a
    // rectangular patch that need not cover the domain; tile sides inside the domain without a neighbouring tile
    // are coarse-fine sides (SIDE_CF).
    LevelGrids crseGrids;
    bool     refined = false;        // m_crseAMRGrids.isClosed()
    int      crseRef[3] = {1, 1, 1};  // refinement ratio to the coarser AMR level (depth 0 only)
    double   amrCrseDXi[3] = {0, 0, 0};  // m_amrCrseDXi: the same at every MG depth (PoissonOp.cpp:345)
    Box3     patch;                  // bounding box of all boxes of this level (= domain on a base level)
    std::shared_ptr<CFLink> cf;      // coarse-fine interpolation, inter-level copiers, fine flux register (depth 0 of a refined level)
    bool     gsrbNatural = false;    // SB_GSRB_KERNEL=natural at creation: point GSRB stays on the natural layout (tests)
    bool     flatZ = false;  // horizontal-only operator of the leptic solver (PoissonOp.cpp:411-505): one layer, no vertical coupling

    std::vector<Box3> boxes;    // all ranks
    std::vector<int>  boxRank;
    std::vector<int>  local;    // indices into boxes owned by this rank
    std::vector<Box3> tiles;    // tile of every rank
    Box3              tile;
    Lay               lay;
    SideBC            side[3][2];
    bool              hasNullSpace = false;
    bool              finalized    = false;

    // device data
    double* J = nullptr;
    double* Dinv = nullptr;
    double* Jgup[3] = {nullptr, nullptr, nullptr};
    double* mtab = nullptr;  // mxl, mxr, myl, myr, mzl, mzr back to back
    double* loBC = nullptr;
    double* hiBC = nullptr;
    int*    boxLoHi = nullptr;  // [2][nlocal][3]
    double* redPartial = nullptr;
    double* redOut = nullptr;
    double* shiftBuf = nullptr;  // (sum, vol) of the last removeKernel
    int*    pivotFlag = nullptr;
    double* colTab = nullptr;    // [J(k) | Dinv(k)] when both are functions of the level only, bit for bit (coefUniform)
    bool    coefUniform = false;
    void    detectColumnCoefficients();
    double* lineTab = nullptr;   // [4][nz] tables of the shared-matrix line relaxation (s, f, g, MzR)
    bool    lineFast = false;
    // colour-split line relaxation (sb_line.cu): layout, tables [6][nz], scratch (cor0, cor1, res0, res1)
    SLay    slay;
    double* lineTabS = nullptr;
    bool    lineSplit = false;
    double* sp[4] = {nullptr, nullptr, nullptr, nullptr};
    // TMA-staged persistent line kernel (sb_line_tma.cu).  lineTma: use it for the shared-matrix case on this depth (large
    // enough, aligned); lineGeneral: horizontally varying metric -- the per-column factorisation is recomputed in the kernel
    // (tables [MzL | MzR], gstart) and the right-hand side is scaled by 1 / (beta J) cell by cell.
    bool    lineTma = false, lineGeneral = false, lineTmaAllowed = true, lineTmaForce = false, tmaMapsReady = false;
    LineTmaMap tmaOth[2], tmaRhs[2];   // tensor maps of sp[0], sp[1] (as the other colour) and sp[2], sp[3] (right-hand sides)
    double* lineTabG = nullptr;         // general: [MzL | MzR]
    double* gstart = nullptr;           // general: [NW - 1][ny][nx]
    double  lineSLo = 0.0, lineSHi = 0.0;
    // one colour pass on sp[], whichever kernel applies; fusedHalo (vertline_tma_k only): the pass also delivers its face
    // layers to the neighbouring tiles (haloLine) and publishes their arrival
    void    linePass(int pass, int region = 0, int nbMask = 0, bool fusedHalo = false);
    const double* splitResSrc = nullptr;  // natural-layout field whose split copy sp[2], sp[3] hold
    // point GSRB on colour-split storage with z ghosts (sb_line.cu: gsrb_split_k): cor, rhs, J, Dinv x 2 colours
    SLay    slayG;
    double* sg[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool    gsrbCoefSplit = false;            // sg[4..7] hold the current J / Dinv
    const double* splitResSrcG = nullptr;
    void   relaxGsrbSplit(double* cor, const double* res, int iters, bool resUnchanged, bool shift);
    // The pass loop of a line relaxation on a small depth is launch-latency bound (two tiny kernels
    // per colour pass): captured once per iteration count into a CUDA graph and replayed.
    struct RelaxGraph { cudaGraphExec_t exec = nullptr; long long kernels = 0; };
    std::map<int, RelaxGraph> relaxGraphs;
    int                       relaxCallsSeen = 0;
    void linePasses(int iters);  // the (ghost fill, colour pass) x 2 x iters loop on sp[]
    std::vector<double> hM[3];  // host copies of the 1-D tables over the whole domain (2*N_d)
    double* xbuf[3][2][2] = {};  // exchange buffers [dir][side][send/recv]
    // neighbour exchange of the split fields by stores into the neighbour's arrays (sb_halo.cu); null: NCCL send / recv.
    // haloLine owns sp[0], sp[1]; haloGsrb owns sg[0], sg[1].
    std::unique_ptr<PeerHalo> haloLine, haloGsrb;

    Op(Context* ctx, const sb_level_desc& d);
    Op(const Op& fine, const int ref[3]);  // coarsening ctor, PoissonOp.cpp:334-405
    Op(Context* single, const Op& dist);   // same depth, every box on the one rank of `single` (agglomeration); J/Jgup left to the caller
    struct HorizTag {};
    Op(const Op& full, HorizTag);          // PoissonOp::createHorizontalMGOperator (PoissonOp.cpp:1647-1686, 411-505)
    ~Op();
    Op& operator=(const Op&) = delete;

    void    setupLayout();
    void    fillMetricFromMap();          // LevelGeometry::createMetricCache, LevelGeometry.cpp:238-277
    void    cacheMatrixElements();        // PoissonOp.cpp:510-665
    void    buildLineTables(double sLo, double sHi);
    void    buildGeneralLine(double sLo, double sHi);
    bool    checkForNullSpace();          // PoissonOp.cpp:670-696
    void    finalize();                   // setAlphaAndBeta, PoissonOp.cpp:707-718
    Coef    coef() const;
    BoxList boxlist() const;
    int     nlocal() const { return (int)local.size(); }
    cudaStream_t st() const { return ctx->st; }

    // LevelOperator / MGOperator / StateOps surface
    void   applyBCs(double* phi, bool homog);
    void   applyBCsWithEdges(double* phi);
    void   exchange(double* phi);
    void   applyOp(double* lhs, double* phi, bool homog);
    void   residual(double* res, double* phi, const double* rhs, bool homog);
    // pre: work fused into the start of the relaxation -- RELAX_PRE_PRECOND: cor = res * Dinv first
    // (preCond(cor, res, 0)); RELAX_PRE_SHIFT: cor -= sum / vol of shiftBuf first (the tail of a
    // removeKernel deferred by MGProlong).  Both are applied even when iters == 0.
    enum { RELAX_PRE_NONE = 0, RELAX_PRE_PRECOND = 1, RELAX_PRE_SHIFT = 2 };
    void   relax(double* cor, const double* res, int iters, bool resUnchanged = false, int pre = RELAX_PRE_NONE);
    void   relaxLineSplit(double* cor, const double* res, int iters, bool resUnchanged, int pre);
    void   preCond(double* phi, const double* rhs, int relaxIters);
    bool   removeKernel(double* phi, bool defer = false);  // true: the subtraction is left to relax(RELAX_PRE_SHIFT)
    double norm(const double* x, int p, double powScale = 1.0);
    double dotProduct(const double* a, const double* b);
    void   incr(double* lhs, const double* x, double scale) { k::incr_valid(st(), lay, lhs, x, scale, -1); }
    void   axby(double* lhs, const double* x, const double* y, double a, double b) { k::axby_valid(st(), lay, lhs, x, y, a, b); }
    void   scale(double* lhs, double s) { k::scale_valid(st(), lay, lhs, s, -1); }
    void   setToZero(double* lhs) { k::fill(st(), lhs, lay.n, 0.0); }
    void   assignLocal(double* dst, const double* src) { k::copy_valid(st(), lay, dst, src); }
    void   MGRestrict(Op& crse, double* crseRes, const double* fineRes);
    bool   MGProlong(Op& crse, double* finePhi, double* crseCor, int order, bool deferKernel = false);  // returns removeKernel's
    void   levelDivergence(double* div, double* const vel[3]);
    void   levelGradient(double* const grad[3], double* phi, bool homog);
    // AMRNSLevel::sendToAdvectingVelocity (toAdvecting) / sendToCartesianVelocity, AMRNSLevelFill.cpp:194-280
    void   scaleVelocity(double* const vel[3], int ghost, bool toAdvecting);
    void   checkPivot();
    int    readPivotFlag(bool reset);
    double* alloc() const;

    // ---- AMRMGOperator surface (Elliptic/AMRMGOperator.H:43-218, PoissonOp.cpp:1156-1478); sb_amr.cpp ----
    void   defineCF();  // m_cfInterp.define(m_grids, m_dXi, m_crseAMRGrids) + copiers (PoissonOp.cpp:118-120)
    // PoissonOp::applyBCs(phi, crsePhiPtr, time, homogPhys, homogCFI) (PoissonOp.cpp:726-763)
    void   applyBCsAMR(double* phi, const Op* crseOp, const double* crsePhi, bool homogCFI);
    void   interpAtCFI(double* phi, const Op& crseOp, const double* crsePhi);  // CFInterp.cpp:387-416 -> MappedQuadCFInterp
    void   AMROperatorNF(double* lhs, double* phi, const Op& crseOp, const double* crsePhi);
    void   AMROperatorNC(double* lhs, Op& fineOp, double* finePhi, double* phi);
    void   AMROperator(double* lhs, Op& fineOp, double* finePhi, double* phi, const Op& crseOp, const double* crsePhi);
    // rhs - L[phi] for whichever neighbours exist (AMRMGOperator.H:107-175); fineOp / crseOp may be null
    void   AMRResidual(double* res, Op* fineOp, double* finePhi, double* phi, const Op* crseOp, const double* crsePhi,
                       const double* rhs);
    void   reflux(double* res, Op& fineOp, double* finePhi, const double* phi);          // PoissonOp.cpp:1333-1418
    void   refluxFlux(double* div, double* const flux[3], Op& fineOp, double* const fineFlux[3]);  // :1425-1477
    double AMRNormLevel(const double* res, const Op* fineOp, int p);                      // :1225-1286
    void   getFlux(double* const flux[3], const double* phi);                             // :1295-1326, all boxes
    void   compDivergence(double* div, double* const flux[3], Op* fineOp, double* const fineFlux[3]);  // :1618-1631
    // CFInterp::coarsen(crse, fine, harmonic = false, J = nullptr) (CFInterp.cpp:846-865): block average of this (fine)
    // level's data onto the cells of the coarser level it covers -- AMRNSLevel::averageDown (AMRNSLevelUtil.cpp:748-781)
    void   averageDownTo(Op& crseOp, double* crse, const double* fine);
};

// partial: the boxes form a rectangular patch inside the domain (refined AMR level) instead of covering it
void planDecomposition(const std::vector<Box3>& boxes, const std::vector<int>& boxRank, const Box3& domain,
                       const int periodic[3], int rank, int nranks, std::vector<Box3>& tiles, std::vector<int>& local,
                       SideBC side[3][2], bool partial = false);
std::vector<std::array<int, 3>> createMGRefScheduleBoxes(int dim, const Box3& domain, const double dXi[3],
                                                         const std::vector<Box3>& boxes, int maxDepth, bool horizStrategy,
                                                         bool doVertCoarsening);

// Stencil records of the quadratic coarse-fine ghost interpolation for one side of one fine box (sb_amr_plan.cpp)
void planCFStencils(const Box3& dom, const int periodic[3], const int ref[3], const std::vector<Box3>& fineBoxes, int box, int dir,
                    int side, std::vector<int>& cells, std::vector<double>& wFirst, std::vector<double>& wSecond,
                    std::vector<double>& wMixed, int dim = 3, const std::vector<Box3>* crseBoxes = nullptr);

// SemicoarseningStrategy / HorizCoarseningStrategy (Elliptic/MGCoarseningStrategy.cpp)
std::vector<std::array<int, 3>> createMGRefSchedule(const Op& top, int maxDepth, bool horizStrategy, bool doVertCoarsening);

struct SolverStatus {
    int    status = SB_STATUS_UNDEFINED;
    double initResNorm = -1.0, finalResNorm = -1.0;
    void clear() { status = SB_STATUS_UNDEFINED; initResNorm = finalResNorm = -1.0; }
};

// Elliptic::BiCGStabSolver<T> (Elliptic/LevelSolverI.H:252-536)
struct BiCGStabSolver {
    Op*               op = nullptr;
    sb_bottom_options opt;
    double *r = nullptr, *r_tilde = nullptr, *e = nullptr, *p = nullptr, *p_tilde = nullptr, *s_tilde = nullptr, *t = nullptr,
           *v = nullptr;
    int    lastIters = 0;
    void   define(Op* op);
    ~BiCGStabSolver();
    SolverStatus solve(double* phi, const double* rhs, bool homog, bool setPhiToZero, double convergenceMetric = -1.0);
};

// Elliptic::MGSolver<T> (Elliptic/MGSolverI.H)
struct MGSolver {
    sb_mg_options                    opt;
    std::vector<std::array<int, 3>>  refSchedule;
    std::vector<Op*>                 ops;        // ops[0] is the caller's top op
    std::vector<std::unique_ptr<Op>> owned;      // depths >= 1
    std::vector<double*>             tmpRes, cor, res;  // per-depth workspace (cor/res of depth d >= 1 are crseCor/crseRes)
    double *                         topRes = nullptr, *topCor = nullptr;
    std::unique_ptr<BiCGStabSolver>  bottom;
    SolverStatus                     status;
    std::vector<double>              absResNorms;  // history of the last solve
    int                              lastIters = 0;
    // Agglomeration: with more than one rank, depths >= aggDepth live on rank 0 only.
    // ops[aggDepth] stays distributed (restriction target / prolongation source); rank 0 also
    // owns aggTop (the same depth, all boxes) and `agg`, the solver of the remaining hierarchy.
    int                              aggDepth = -1;
    std::unique_ptr<Context>         aggCtx;
    std::unique_ptr<Op>              aggTop;
    std::unique_ptr<MGSolver>        agg;
    double *                         aggRes = nullptr, *aggCor = nullptr, *aggBuf = nullptr;
    void defineAgglomeration();
    void aggGather(const double* tileField, double* fullField, int centering, const Op& distOp);
    void aggScatter(double* tileField, const double* fullField, const Op& distOp);
    void checkPivotAll();
    bool needPivotCheck = false;
    int  localPivotFlag(bool reset);
    bool usesGeneralLineKernel() const;

    void define(Op& top, const sb_mg_options& opt, std::vector<std::array<int, 3>> sched, bool useBottomSolver = true);
    ~MGSolver();
    SolverStatus solve(double* phi, const double* rhs, bool homog, bool setPhiToZero, double convergenceMetric = -1.0);
    SolverStatus cycle(bool fmgMode, double* phi, const double* rhs, bool homog, bool setPhiToZero, double convergenceMetric);
    // corIsPreCond: cor still has to be set to preCond(res) -- the caller skipped that pass so that
    // the first relaxation can fuse it
    void vCycle_residualEq(double* cor, const double* res, int depth, bool corIsPreCond = false);
    // sb_tiny.cu: depths >= tinyTailStart() run as one single-CTA kernel; false: not applicable here
    int  tailStart = -2;
    double* tailOut = nullptr;  // device record of the last tail launch (bottom-solve status, time split)
    int  tinyTailStart();
    bool tinyTail(int depth, double* cor, const double* res, bool corIsPreCond);
    void fmg_residualEq(double* cor, const double* res, int depth);
    void modifyOptionsExceptMaxDepth(const sb_mg_options& o);
};

// Elliptic::LevelLepticSolver (Elliptic/LevelLepticSolver.cpp): vertical Neumann-Neumann line
// solves + one horizontal (flattened) MG solve per call, iterated over "orders".
struct LepticSolver {
    Op*                 op = nullptr;
    std::unique_ptr<Op> hOp;   // horizontal operator on the flattened grids
    MGSolver            hmg;   // m_horizSolverPtr
    // LevelLepticSolver::Options (LevelLepticSolver.cpp:22-60)
    double              absTol = 0, relTol = 0, hang = 0;
    int                 maxOrder = 0, normType = 2, maxDivergingOrders = 2;
    sb_mg_options       horizOptions;
    bool                horizRemoveAvg = true;
    std::vector<double> resNorms;
    double *corTotal = nullptr, *cor = nullptr, *rhsA = nullptr, *rhsB = nullptr, *gam = nullptr;  // on op
    double *excess = nullptr, *hiBC = nullptr, *hPhi = nullptr, *hRhs = nullptr, *ones = nullptr; // on hOp
    void define(Op& top, const sb_mg_options& proj);
    // LevelLepticSolver::modifyOptionsExceptMaxDepth (LevelLepticSolver.cpp:100-110): the leptic fields follow the
    // new (proj.*-shaped) options; the horizontal MGSolver keeps the options it was defined with, as in the reference
    void modifyOptionsExceptMaxDepth(const sb_mg_options& proj);
    ~LepticSolver();
    SolverStatus solve(double* phi, const double* rhs, bool homog, bool setPhiToZero);
    void computeVerticalExcess(const double* rhs);
    void verticalLineSolver(double* vertPhi, const double* vertRhs);
    SolverStatus horizontalSolver();
    void setZeroAvg(double* hphi);
};

// Elliptic::LevelHybridSolver (Elliptic/LevelHybridSolver.cpp).
struct HybridSolver {
    Op*                 op = nullptr;
    int                 mode = 0;
    bool                isHybrid = false;  // false: plain MGSolver behind the same handle
    MGSolver            mg;
    std::unique_ptr<LepticSolver> leptic;
    int                 maxSolverSwaps = 10;  // LevelHybridSolver.cpp:18
    double *            cor = nullptr, *res = nullptr, *localRes = nullptr;
    std::vector<double> resNorms;
    sb_mg_options       opt;
    // LevelHybridSolver::getQuickAndDirtyOptions (LevelHybridSolver.cpp:47-60) switched on / off around the fallback
    // projection of AMRNSLevel::projectPredict (AMRNSLevelProject.cpp:176-182): hybrid relTol 1e-2, one solver swap, the
    // MGSolver's quick-and-dirty options (MGSolverI.H:56-74); the leptic options stay (LevelLepticSolver.cpp:67-84)
    sb_mg_options       optDefine, savedOpt, savedMgOpt;
    int                 savedSwaps = 10;
    bool                quickAndDirty = false;
    void setQuickAndDirty(bool on);
    // the projection bracket on device-resident fields (sb_project_correct / sb_project_predict): persistent temporaries
    double *            pDiv = nullptr, *pPhi = nullptr, *pGrad[3] = {nullptr, nullptr, nullptr};
    void   projectTemps();
    // AMRNSLevel::projectCorrect (AMRNSLevelProject.cpp:247-373) / projectPredict (:63-243) without the level's own BC fills
    SolverStatus projectCorrect(double* const vel[3], double* p, double projDt, int velGhost, double* phiOut, double* initDivNorm,
                                double* finalDivNorm);
    SolverStatus projectPredict(double* const vel[3], double* p, double projDt, int velGhost, double norms[3], bool* usedFallback);
    static int computeSolveMode(const Op& op);  // LevelHybridSolver.cpp:457-498
    void define(Op& top, const sb_mg_options& o);
    ~HybridSolver();
    SolverStatus solve(double* phi, const double* rhs, bool homog, bool setPhiToZero, double convergenceMetric);
};

// Elliptic::AMRHybridSolver (Elliptic/AMRHybridSolver.cpp): composite solve over AMR levels lmin..lmax; the smoothers
// of its V-cycle are whole level solves (LevelHybridSolver), the levels talk through the operators' AMR interface.
struct AMRSolver {
    std::vector<Op*>                           ops;
    int                                        lbase = 0, lmin = 0, lmax = 0;
    sb_mg_options                              opt;
    std::vector<std::unique_ptr<HybridSolver>> hybrid;
    std::vector<double*>                       ve, vr, scratch;  // m_ve, m_vr, m_vScratchPhi (m_vResC lives in the level's CFLink)
    SolverStatus                               status;
    std::vector<double>                        absResNorms;
    int                                        lastIters = 0;
    void define(const std::vector<Op*>& ops, int lmin, int lmax, const sb_mg_options& o);
    ~AMRSolver();
    SolverStatus solve(std::vector<double*>& phi, const std::vector<const double*>& rhs, bool homog, bool setPhiToZero, double metric);
    void   amrVCycle_residualEq(std::vector<double*>& phi, std::vector<double*>& rhs, int lev);
    double computeAMRResidual(std::vector<double*>& res, const std::vector<double*>& phi, const std::vector<const double*>& rhs,
                              bool computeNorm);
    void   computeAMRResidualLevel(double* res, double* finePhi, double* phi, const double* crsePhi, const double* rhs, int lev);
};

}  // namespace sb

struct sb_amr_solver { sb::AMRSolver s; };
struct sb_context { sb::Context c; sb_context(int d, int r, int n) : c(d, r, n) {} };
struct sb_op { sb::Op* op; bool owned; };
struct sb_field { sb::Field f; sb_field(sb::Op* op, int c) : f(op, c) {} };
struct sb_solver { sb::HybridSolver s; };
