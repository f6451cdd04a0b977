// sb_leptic.cpp -- the leptic branch of the projection's level solver: Elliptic::LevelLepticSolver
// (reference Grade3_Calculus/Elliptic/LevelLepticSolver.cpp) and the horizontal-only operator it
// solves with (PoissonOp.cpp:411-505, 1647-1686).  LevelHybridSolver picks this branch when the
// grid is leptic: lepticity = min(dXi_x, dXi_y) / L_z > 0.2 (LevelHybridSolver.cpp:457-498).
//
// The reference re-grids to vertically spanning boxes (LepticBoxTools::createVerticalSolverGrids)
// and to flattened copies of them for the horizontal problem.  On this path a rank's tile always
// spans the vertical, and the box list is required to do so as well; the vertical grids are then
// the operator's own boxes and the horizontal grids their flattened images (vertical index 0,
// Subspace::flattenBox), owned by the same ranks.
#include <cmath>
#include <cstring>
#include <limits>

#include "sb_host.h"

namespace sb {

// ---------------------------------------------------------------------------------------------
// PoissonOp::createHorizontalMGOperator -> the "horizontal-only op" constructor
// (PoissonOp.cpp:411-505): flat grids, dXi_z = 1, J = dx/dXi * dy/dEta, Jg^{zz} = 0, vertical
// direction inactive (m_M[z] = 0, no vertical ghosts), alpha = 0, beta = 1, relaxation 6 -> 5.
Op::Op(const Op& f, HorizTag) : ctx(f.ctx)
{
    dim = f.dim; alpha = 0.0; beta = 1.0; map = f.map; depth = 0; flatZ = true;
    relaxMethod = f.relaxMethod == SB_RELAX_VERTLINE ? SB_RELAX_GSRB : f.relaxMethod;
    std::memcpy(periodic, f.periodic, sizeof(periodic));
    std::memcpy(bcAlpha, f.bcAlpha, sizeof(bcAlpha));
    std::memcpy(bcBeta, f.bcBeta, sizeof(bcBeta));
    std::memcpy(dXi, f.dXi, sizeof(dXi));
    dXi[2]      = 1.0;
    periodic[2] = 0;
    // a refined patch keeps its coarse-fine sides in the horizontal problem (PoissonOp.cpp:1649-1672, 469-479)
    refined = f.refined;
    for (int d = 0; d < 3; ++d) amrCrseDXi[d] = f.amrCrseDXi[d];
    amrCrseDXi[2] = 1.0;  // flat domains on both levels: ratio 1, dXi_z = 1
    domain      = f.domain;
    domain.lo[2] = domain.hi[2] = 0;
    boxRank = f.boxRank;
    boxes   = f.boxes;
    for (Box3& b : boxes) {
        if (b.lo[2] != f.domain.lo[2] || b.hi[2] != f.domain.hi[2])
            SB_FAIL("the leptic solver on the B200 path needs boxes that span the vertical (base.splitDirs = 1 1 0)");
        b.lo[2] = b.hi[2] = 0;
    }
    setupLayout();
    J = alloc();
    for (int d = 0; d < 3; ++d) Jgup[d] = alloc();
    // Metric: GeoSourceInterface::fill_dxdXi / fill_dXidx into ghost-free per-box holders (xi
    // accumulated from the box's own small end), vertical factors 1.
    const double one[2] = {1.0, 1.0};
    for (int lb = 0; lb < nlocal(); ++lb) {
        const Box3&         b = boxes[local[lb]];
        std::vector<double> tab;
        size_t              off[4];
        for (int mu = 0; mu < 2; ++mu) {
            const int           n = b.size(mu);
            std::vector<double> c, fc;
            if (dim == 2 && mu == 1) { c.assign(n, 1.0); fc.assign(n + 1, 1.0); }
            else { c = map.dxdXi(mu, dXi[mu], b.lo[mu], n, 0); fc = map.dxdXi(mu, dXi[mu], b.lo[mu], n + 1, 1); }
            off[mu] = tab.size(); tab.insert(tab.end(), c.begin(), c.end());
            off[2 + mu] = tab.size(); tab.insert(tab.end(), fc.begin(), fc.end());
        }
        const size_t o1 = tab.size();
        tab.insert(tab.end(), one, one + 2);
        double* dt = (double*)ctx->getScratch(tab.size() * sizeof(double));
        SB_CUDA(cudaMemcpyAsync(dt, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
        int blo[3], bhi[3];
        for (int i = 0; i < 3; ++i) { blo[i] = b.lo[i] - tile.lo[i]; bhi[i] = b.hi[i] - tile.lo[i]; }
        k::fill_metric_box(st(), lay, blo, bhi, dt + off[0], dt + off[1], dt + o1, dt + off[2], dt + off[3], dt + o1, J, Jgup[0],
                           Jgup[1], Jgup[2]);
        ctx->sync();
    }
    k::fill(st(), Jgup[2], lay.n, 0.0);  // m_Jgup[dit][SpaceDim - 1].setVal(0.0)
    finalize();                          // setAlphaAndBeta(0, 1): matrix elements + null-space check
}

// ---------------------------------------------------------------------------------------------
LepticSolver::~LepticSolver()
{
    for (double* q : {corTotal, cor, rhsA, rhsB, gam, excess, hiBC, hPhi, hRhs, ones})
        if (q) cudaFree(q);
}

// LevelLepticSolver::define (LevelLepticSolver.cpp:169-352) with getDefaultOptions (:22-60), whose
// leptic fields come from ProjectorParameters -- the same source as the MG options handed in.
void LepticSolver::define(Op& top, const sb_mg_options& proj)
{
    op = &top;
    if (top.periodic[2]) SB_FAIL("LevelLepticSolver: the vertical must not be periodic");
    if (top.lay.nz < 2) SB_FAIL("LevelLepticSolver: needs at least 2 cells in the vertical");
    absTol = proj.absTol; relTol = proj.relTol; maxOrder = proj.maxIters; normType = proj.normType; hang = proj.hang;
    maxDivergingOrders = 2;
    sb_mg_options& h = horizOptions;
    sb_mg_default_options(&h);
    h.absTol = 1.0e-15; h.relTol = 1.0e-15; h.convergenceMetric = -1.0;
    h.numSmoothDown = 4; h.numSmoothUp = 4; h.numSmoothBottom = 2; h.numSmoothPrecond = 2;
    h.prolongOrder = 1; h.prolongOrderFMG = 3; h.numSmoothUpFMG = 0;
    h.maxDepth = -1; h.numCycles = 1; h.maxIters = 20; h.hang = 0.01; h.normType = normType; h.verbosity = 0;
    h.bottom.absTol = 1.0e-15; h.bottom.relTol = 1.0e-15; h.bottom.small = 1.0e-30; h.bottom.hang = 0.01;
    h.bottom.maxIters = 80; h.bottom.maxRestarts = 5; h.bottom.normType = normType; h.bottom.verbosity = 0;
    h.bottom.numSmoothPrecond = 2; h.bottom.convergenceMetric = -1.0;

    hOp.reset(new Op(top, Op::HorizTag{}));
    // HorizCoarseningStrategy(m_L, doVertCoarsening = false) on the horizontal grids (:309-318)
    hmg.define(*hOp, h, createMGRefSchedule(*hOp, h.maxDepth, true, false), true);
    horizOptions = hmg.opt;
    // m_horizRemoveAvg (:327-348): the horizontal grids cover the horizontal domain
    long long pts = 0;
    for (const Box3& b : hOp->boxes) pts += b.numPts();
    horizRemoveAvg = pts == hOp->domain.numPts();

    corTotal = top.alloc(); cor = top.alloc(); rhsA = top.alloc(); rhsB = top.alloc(); gam = top.alloc();
    excess = hOp->alloc(); hiBC = hOp->alloc(); hPhi = hOp->alloc(); hRhs = hOp->alloc(); ones = hOp->alloc();
    k::fill(top.st(), ones, hOp->lay.n, 1.0);
}

void LepticSolver::modifyOptionsExceptMaxDepth(const sb_mg_options& proj)
{
    absTol = proj.absTol; relTol = proj.relTol; maxOrder = proj.maxIters; normType = proj.normType; hang = proj.hang;
}

// LevelLepticSolver::computeVerticalExcess (:711-770)
void LepticSolver::computeVerticalExcess(const double* rhs)
{
    k::vert_excess(op->st(), op->lay, hOp->lay, excess, hiBC, rhs, -op->dXi[2]);
}

// LevelLepticSolver::verticalLineSolver (:776-850)
void LepticSolver::verticalLineSolver(double* vertPhi, const double* vertRhs)
{
    k::tridiag_nn(op->st(), op->lay, hOp->lay, vertPhi, vertRhs, hiBC, op->Jgup[2], gam, op->dXi[2]);
}

// LevelLepticSolver::setZeroAvg (:1170-1225): phi -= (sum over boxes of FArrayBox::sum) / numPts
void LepticSolver::setZeroAvg(double* hphi)
{
    Op&          h   = *hOp;
    const double sum = h.dotProduct(hphi, ones);
    Context*     c   = h.ctx;
    c->hpin[0]       = sum;
    c->hpin[1]       = (double)h.domain.numPts();
    SB_CUDA(cudaMemcpyAsync(h.redOut, c->hpin, 2 * sizeof(double), cudaMemcpyHostToDevice, c->st));
    k::add_scalar_valid(h.st(), h.lay, hphi, h.redOut);
    c->sync();  // hpin is reused by the next reduction
}

// LevelLepticSolver::horizontalSolver (:856-892)
SolverStatus LepticSolver::horizontalSolver()
{
    SolverStatus st = hmg.solve(hPhi, hRhs, true, true);
    if (horizRemoveAvg) setZeroAvg(hPhi);
    return st;
}

// LevelLepticSolver::solve (:400-706)
SolverStatus LepticSolver::solve(double* a_phi, const double* a_rhs, bool homog, bool setPhiToZero)
{
    SolverStatus status;
    Op&          o = *op;
    Op&          h = *hOp;
    bool useExcess = true, useHorizPhi = true;  // m_doHorizSolve
    int  numDivergingOrders = 0;
    double* rhs    = rhsA;
    double* tmpRhs = rhsB;

    if (setPhiToZero) o.setToZero(a_phi);
    o.residual(rhs, a_phi, a_rhs, homog);
    o.setToZero(corTotal);
    resNorms.clear();
    resNorms.push_back(o.norm(rhs, normType));
    h.setToZero(hiBC);  // bdryData.setVal(0.0)

    for (int order = 0; order <= maxOrder; ++order) {
        if (order >= 1 && useExcess) h.incr(hiBC, excess, 1.0);   // bdryData.vertPlus(excess, 1.0, Side::Hi)
        if (useExcess) {
            computeVerticalExcess(rhs);
            if (order == 1) useExcess = false;
        }
        if (order == 0 && useExcess) h.incr(hiBC, excess, -1.0);

        verticalLineSolver(cor, rhs);

        if (useHorizPhi) {
            h.setToZero(hRhs);  // flatRhs = 0 (+ excess * (-1/H)), horizRhs = 0 + flatRhs
            if (useExcess) h.incr(hRhs, excess, -1.0 / ((double)o.domain.size(2) * o.dXi[2]));
            horizontalSolver();
            k::add_vertical_extrusion(o.st(), o.lay, h.lay, cor, hPhi);  // addHorizontalCorrection
        }

        o.incr(corTotal, cor, 1.0);
        o.residual(tmpRhs, cor, rhs, true);
        resNorms.push_back(o.norm(tmpRhs, normType));

        if (resNorms.back() <= absTol) { status.status = SB_STATUS_CONVERGED; break; }
        else if (resNorms.back() <= relTol * resNorms[0]) { status.status = SB_STATUS_CONVERGED; break; }
        else if (resNorms[order + 1] > resNorms[order]) {
            if (numDivergingOrders < maxDivergingOrders) ++numDivergingOrders;
            else { status.status = SB_STATUS_DIVERGED; break; }
        } else if (resNorms[order + 1] > (1.0 - hang) * resNorms[order]) { status.status = SB_STATUS_HANG; break; }
        else {
            numDivergingOrders = 0;
            if (order == maxOrder) status.status = SB_STATUS_MAXITERS;
        }
        std::swap(rhs, tmpRhs);
        useHorizPhi = false;
    }
    o.incr(a_phi, corTotal, 1.0);  // corTotal.addTo(a_phi)
    status.finalResNorm = resNorms.back();
    return status;
}

}  // namespace sb
