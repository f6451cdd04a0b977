// sb_mg.cpp -- the solver control flow of the projection, mirrored from the reference so that
// iteration counts and convergence decisions are identical: MGSolver<T> (Elliptic/MGSolverI.H),
// BiCGStabSolver<T> (Elliptic/LevelSolverI.H:252-536), the coarsening strategies
// (Elliptic/MGCoarseningStrategy.cpp) and LevelHybridSolver (Elliptic/LevelHybridSolver.cpp).
// Fields live on the device; only scalars (norms, dot products) come back to the host, exactly
// where the reference performs an MPI_Allreduce.
#include <cstdio>
#include <algorithm>
#include <cmath>
#include <limits>

#include "sb_comm.h"
#include "sb_host.h"

namespace sb {

// ---------------------------------------------------------------------------------------------
// Coarsening strategies.
namespace {
using IV = std::array<int, 3>;

bool coarsenableAll(const std::vector<Box3>& boxes, const IV& r)
{
    for (const Box3& b : boxes)
        if (!coarsenable(b, r.data())) return false;
    return true;
}
// SemicoarseningStrategy::measureAnisotropy (MGCoarseningStrategy.H:103-107)
double isoSemi(const double dx[3], int dim)
{
    if (dim == 2) { const double m = std::min(dx[0], dx[2]); return (dx[0] / m) * (dx[2] / m); }
    const double m = std::min(std::min(dx[0], dx[1]), dx[2]);
    return (dx[0] / m) * (dx[1] / m) * (dx[2] / m);
}
// HorizCoarseningStrategy::measureAnisotropy (MGCoarseningStrategy.H:153-156)
double isoHoriz(const double dx[3], int dim)
{
    if (dim == 2) { const double m = std::min(dx[0], dx[0]); return dx[0] / m; }
    const double m = std::min(dx[0], dx[1]);
    return (dx[0] / m) * (dx[1] / m);
}
// computeNextRefRatio (MGCoarseningStrategy.cpp:138-175, 281-310)
IV nextRef(const double dx[3], const std::vector<Box3>& grids, const std::vector<IV>& refList, bool horiz, int dim)
{
    const double isoLimiter = 0.75;
    auto         iso        = [&](const double* d) { return horiz ? isoHoriz(d, dim) : isoSemi(d, dim); };
    size_t       isoIdx     = std::numeric_limits<size_t>::max();
    double       isoVal     = iso(dx);
    const double lastIsoVal = isoVal;
    for (size_t cur = 0; cur < refList.size(); ++cur) {
        const IV&    r        = refList[cur];
        const double cdx[3]   = {dx[0] * r[0], dx[1] * r[1], dx[2] * r[2]};
        const double curIso   = iso(cdx);
        if (!coarsenableAll(grids, r)) continue;
        if (curIso < isoLimiter * lastIsoVal && isoIdx < refList.size()) continue;
        if (curIso <= isoVal) { isoIdx = cur; isoVal = curIso; }
    }
    if (isoIdx < refList.size()) return refList[isoIdx];
    return IV{1, 1, 1};
}
}  // namespace

std::vector<IV> createMGRefSchedule(const Op& top, int maxDepth, bool horizStrategy, bool doVertCoarsening)
{
    return createMGRefScheduleBoxes(top.dim, top.domain, top.dXi, top.boxes, maxDepth, horizStrategy, doVertCoarsening);
}
std::vector<IV> createMGRefScheduleBoxes(int dim, const Box3& domain, const double dXi[3], const std::vector<Box3>& boxes,
                                         int maxDepth, bool horizStrategy, bool doVertCoarsening)
{
    std::vector<IV> refList;
    if (!horizStrategy) {
        if (dim == 2) refList = {{2, 1, 1}, {1, 1, 2}, {2, 1, 2}};
        else refList = {{2, 1, 1}, {1, 2, 1}, {1, 1, 2}, {1, 2, 2}, {2, 1, 2}, {2, 2, 1}, {2, 2, 2}};
    } else {
        if (dim == 2) refList = {{2, 1, 1}};
        else refList = {{2, 1, 1}, {1, 2, 1}, {2, 2, 1}};
    }
    // curDx = L / N with L = N * dXi (MGSolverI.H:148-151, MGCoarseningStrategy.cpp:68)
    double curDx[3];
    for (int d = 0; d < 3; ++d) {
        const double N = (double)domain.size(d);
        curDx[d]       = (N * dXi[d]) / N;
    }
    std::vector<Box3> cur = boxes;  // minBoxSize = 1: coarsening by it is a no-op
    std::vector<IV>   sched;
    while (true) {
        if (maxDepth >= 0 && sched.size() == (size_t)maxDepth) break;
        IV r = nextRef(curDx, cur, refList, horizStrategy, dim);
        if (r == IV{1, 1, 1}) break;
        if (horizStrategy && doVertCoarsening && coarsenableAll(cur, IV{1, 1, 2})) r[2] = 2;
        for (Box3& b : cur) b = coarsen(b, r.data());
        for (int d = 0; d < 3; ++d) curDx[d] *= (double)r[d];
        sched.push_back(r);
    }
    sched.push_back(IV{1, 1, 1});
    return sched;
}

// ---------------------------------------------------------------------------------------------
// Options.
}  // namespace sb

extern "C" void sb_mg_default_options(sb_mg_options* o)
{
    // Grade5_SOMAR/ProjectorParameters.cpp:124-222 defaults, as MGSolverI.H:22-50 copies them.
    o->absTol = 1.0e-12; o->relTol = 1.0e-6; o->convergenceMetric = -1.0; o->hang = 1.0e-7;
    o->numSmoothDown = 16; o->numSmoothUp = 16; o->numSmoothBottom = 2; o->numSmoothPrecond = 2;
    o->prolongOrder = 1; o->prolongOrderFMG = 3; o->numSmoothUpFMG = 2;
    o->maxDepth = -1; o->numCycles = -1; o->maxIters = 10; o->normType = 2; o->verbosity = 0;
    o->bottom.absTol = 1.0e-6; o->bottom.relTol = 1.0e-4; o->bottom.small = 1.0e-30; o->bottom.hang = 1.0e-7;
    o->bottom.convergenceMetric = -1.0;
    o->bottom.maxIters = 80; o->bottom.maxRestarts = 5; o->bottom.normType = 2; o->bottom.verbosity = 0;
    o->bottom.numSmoothPrecond = 2;
}
extern "C" void sb_mg_quick_and_dirty_options(sb_mg_options* o)
{
    // MGSolverI.H:56-74
    sb_mg_default_options(o);
    o->absTol = 1.0e-300; o->relTol = 1.0e-300; o->numCycles = -1; o->maxIters = 1; o->verbosity = 0;
    o->bottom.absTol = 1.0e-300; o->bottom.relTol = 1.0e-300; o->bottom.verbosity = 0;
}

namespace sb {

// ---------------------------------------------------------------------------------------------
// BiCGStab bottom solver (LevelSolverI.H:252-536).
void BiCGStabSolver::define(Op* o)
{
    op = o;
    r = op->alloc(); r_tilde = op->alloc(); e = op->alloc(); p = op->alloc();
    p_tilde = op->alloc(); s_tilde = op->alloc(); t = op->alloc(); v = op->alloc();
}
BiCGStabSolver::~BiCGStabSolver()
{
    for (double* q : {r, r_tilde, e, p, p_tilde, s_tilde, t, v})
        if (q) cudaFree(q);
}
SolverStatus BiCGStabSolver::solve(double* phi, const double* rhs, bool homog, bool setPhiToZero, double a_convergenceMetric)
{
    SolverStatus st;
    st.status = SB_STATUS_UNDEFINED;
    Op& o     = *op;
    if (setPhiToZero) o.setToZero(phi);
    int recount = 0;
    o.residual(r, phi, rhs, homog);
    o.assignLocal(r_tilde, r);
    o.setToZero(e);
    o.setToZero(p_tilde);
    o.setToZero(s_tilde);
    int    i      = 0;
    double rho[4] = {0, 0, 0, 0};
    double norm[2];
    norm[0]              = o.norm(r, opt.normType);
    double initial_norm  = norm[0];
    double initial_rnorm = norm[0];
    norm[1]              = norm[0];
    st.initResNorm       = initial_norm;
    double alpha[2] = {0, 0}, beta[2] = {0, 0}, omega[2] = {0, 0};
    bool   init     = true;
    int    restarts = 0;
    if (opt.convergenceMetric > 0.0) initial_norm = opt.convergenceMetric;
    if (a_convergenceMetric > 0.0) initial_norm = a_convergenceMetric;
    const double smallReal = 1.0e4 * std::numeric_limits<double>::epsilon();

    while ((i < opt.maxIters && norm[0] > opt.absTol * norm[1]) && (norm[1] > 0)) {
        i++;
        norm[1] = norm[0]; alpha[1] = alpha[0]; beta[1] = beta[0]; omega[1] = omega[0];
        rho[3] = rho[2]; rho[2] = rho[1];
        rho[1] = o.dotProduct(r_tilde, r);
        if (std::abs(rho[1]) < smallReal) {  // RealCmp::isZero
            o.incr(phi, e, 1.0);
            st.finalResNorm = initial_norm;
            st.status       = SB_STATUS_SINGULAR;
            lastIters       = i;
            return st;
        }
        if (init) {
            o.assignLocal(p, r);
            init = false;
        } else {
            beta[1] = (rho[1] / rho[2]) * (alpha[1] / omega[1]);
            o.scale(p, beta[1]);
            o.incr(p, v, -beta[1] * omega[1]);
            o.incr(p, r, 1.0);
        }
        o.preCond(p_tilde, p, opt.numSmoothPrecond);
        o.applyOp(v, p_tilde, true);
        const double m = o.dotProduct(r_tilde, v);
        alpha[0]       = rho[1] / m;
        if (std::abs(m) > opt.small * std::abs(rho[1])) {
            o.incr(r, v, -alpha[0]);
            norm[0] = o.norm(r, opt.normType);
            o.incr(e, p_tilde, alpha[0]);
        } else {
            o.setToZero(r);
            norm[0] = 0.0;
        }
        if (norm[0] > opt.absTol * initial_norm && norm[0] > opt.relTol * initial_rnorm) {
            o.preCond(s_tilde, r, opt.numSmoothPrecond);
            o.applyOp(t, s_tilde, true);
            omega[0] = o.dotProduct(t, r) / o.dotProduct(t, t);
            o.incr(e, s_tilde, omega[0]);
            o.incr(r, t, -omega[0]);
            norm[0] = o.norm(r, opt.normType);
        }
        if (norm[0] <= opt.absTol * initial_norm || norm[0] <= opt.relTol * initial_rnorm) {
            st.finalResNorm = norm[0];
            st.status       = SB_STATUS_CONVERGED;
            break;
        }
        if (omega[0] == 0.0 || norm[0] > (1.0 - opt.hang) * norm[1]) {
            if (recount == 0) {
                recount = 1;
            } else {
                recount = 0;
                o.incr(phi, e, 1.0);
                if (restarts == opt.maxRestarts) {
                    st.finalResNorm = norm[0];
                    st.status       = SB_STATUS_MAXITERS;
                    lastIters       = i;
                    return st;
                }
                o.residual(r, phi, rhs, homog);
                norm[0] = o.norm(r, opt.normType);
                rho[1] = 0.0; rho[2] = 0.0; rho[3] = 0.0;
                alpha[0] = 0; beta[0] = 0; omega[0] = 0;
                o.assignLocal(r_tilde, r);
                o.setToZero(e);
                restarts++;
                init = true;
            }
        }
    }
    o.incr(phi, e, 1.0);
    st.finalResNorm = norm[0];
    lastIters       = i;
    return st;
}

// ---------------------------------------------------------------------------------------------
// MGSolver<T>::define (MGSolverI.H:140-204)
void MGSolver::define(Op& top, const sb_mg_options& a_opt, std::vector<IV> sched, bool useBottomSolver)
{
    opt = a_opt;
    tailStart = -2;
    if (sched.empty()) {
        if (top.relaxMethod == SB_RELAX_VERTLINE) refSchedule = createMGRefSchedule(top, opt.maxDepth, true, true);
        else refSchedule = createMGRefSchedule(top, opt.maxDepth, false, false);
    } else {
        refSchedule = sched;
    }
    if (refSchedule.empty() || refSchedule.back() != IV{1, 1, 1}) SB_FAIL("the MG ref schedule must end with (1,1,1)");
    opt.maxDepth = (int)refSchedule.size() - 1;
    // Agglomeration: the first depth >= 1 whose whole domain has at most SB_AGG_CELLS cells
    // (default 4 Mi; 0 disables) and everything below it run on rank 0 alone.  The arithmetic is
    // decomposition-independent (colours use global indices, ghosts are refreshed between
    // colours), so this changes where the work runs, not its result.
    aggDepth = -1;
    if (top.ctx->nranks > 1 && opt.maxDepth >= 1) {
        const char* e     = getenv("SB_AGG_CELLS");
        long long   limit = e ? atoll(e) : (4LL << 20);
        Box3        dom   = top.domain;
        for (int d = 1; d <= opt.maxDepth && limit > 0; ++d) {
            dom = coarsen(dom, refSchedule[d - 1].data());
            if (dom.numPts() <= limit) { aggDepth = d; break; }
        }
    }
    const int lastDist = aggDepth >= 0 ? aggDepth : opt.maxDepth;  // deepest depth held in ops[]
    ops.assign(lastDist + 1, nullptr);
    ops[0] = &top;
    for (int d = 1; d <= lastDist; ++d) {
        owned.emplace_back(new Op(*ops[d - 1], refSchedule[d - 1].data()));
        ops[d] = owned.back().get();
    }
    tmpRes.assign(lastDist + 1, nullptr);
    cor.assign(lastDist + 1, nullptr);
    res.assign(lastDist + 1, nullptr);
    for (int d = 0; d <= lastDist; ++d) {
        if (d != aggDepth) tmpRes[d] = ops[d]->alloc();
        if (d > 0) { cor[d] = ops[d]->alloc(); res[d] = ops[d]->alloc(); }
    }
    topRes = top.alloc();
    topCor = top.alloc();
    if (aggDepth >= 0) defineAgglomeration();
    else if (useBottomSolver) {
        bottom.reset(new BiCGStabSolver);
        bottom->define(ops[opt.maxDepth]);
        bottom->opt = opt.bottom;
    }
    status.clear();
    double need = usesGeneralLineKernel() ? 1.0 : 0.0;
    if (!top.ctx->parent) top.ctx->allreduceMax(&need, 1);  // (the agglomerated solver on rank 0 answers through its owner)
    needPivotCheck = need != 0.0;
}
void MGSolver::defineAgglomeration()
{
    Op&      dist = *ops[aggDepth];
    Context* ctx  = dist.ctx;
    // staging: every tile with room for its far faces
    auto ext = [](const Box3& t) { return (size_t)(t.size(0) + 1) * (t.size(1) + 1) * (t.size(2) + 1); };
    size_t n = ext(dist.tile);
    if (ctx->rank == 0) for (const Box3& t : dist.tiles) n += ext(t);
    SB_CUDA(cudaMalloc((void**)&aggBuf, n * sizeof(double)));
    if (ctx->rank == 0) {
        aggCtx.reset(new Context(*ctx));
        aggTop.reset(new Op(aggCtx.get(), dist));
    }
    const Lay* full = aggTop ? &aggTop->lay : nullptr;
    ctx->comm->gatherTiles(dist, dist.J, full, aggTop ? aggTop->J : nullptr, SB_CELL, aggBuf);
    for (int d = 0; d < 3; ++d) {
        if (dist.dim == 2 && d == 1) continue;
        ctx->comm->gatherTiles(dist, dist.Jgup[d], full, aggTop ? aggTop->Jgup[d] : nullptr, d, aggBuf);
    }
    ctx->sync();
    if (ctx->rank != 0) return;
    aggTop->cacheMatrixElements();
    aggTop->finalized = true;
    aggRes = aggTop->alloc();
    aggCor = aggTop->alloc();
    agg.reset(new MGSolver);
    sb_mg_options o = opt;
    o.maxDepth      = opt.maxDepth - aggDepth;
    agg->define(*aggTop, o, std::vector<IV>(refSchedule.begin() + aggDepth, refSchedule.end()), true);
    agg->bottom->opt = opt.bottom;
}
void MGSolver::aggGather(const double* tileField, double* fullField, int centering, const Op& distOp)
{
    distOp.ctx->comm->gatherTiles(distOp, tileField, aggTop ? &aggTop->lay : nullptr, fullField, centering, aggBuf);
}
void MGSolver::aggScatter(double* tileField, const double* fullField, const Op& distOp)
{
    distOp.ctx->comm->scatterTiles(distOp, tileField, aggTop ? &aggTop->lay : nullptr, fullField, aggBuf);
}
// Only the general line kernel (vertline_k) can raise the flag; whether any depth on any rank uses it is settled once
// in define(), so the shared-matrix path pays nothing here.  The flags of all depths (agglomerated ones included, which
// exist on rank 0 only) are combined over the ranks before anyone throws: every rank fails together instead of one rank
// leaving the others in the next collective.
int MGSolver::localPivotFlag(bool reset)
{
    int f = 0;
    for (Op* o : ops)
        if (o->relaxMethod == SB_RELAX_VERTLINE && !o->lineFast) f |= o->readPivotFlag(reset);
    if (agg) f |= agg->localPivotFlag(reset);
    return f;
}
bool MGSolver::usesGeneralLineKernel() const
{
    for (Op* o : ops)
        if (o->relaxMethod == SB_RELAX_VERTLINE && !o->lineFast) return true;
    return agg ? agg->usesGeneralLineKernel() : false;
}
void MGSolver::checkPivotAll()
{
    if (!needPivotCheck || ops.empty()) return;
    double f = (double)localPivotFlag(true);
    ops[0]->ctx->allreduceMax(&f, 1);
    if (f != 0.0)
        SB_FAIL("vertical line relaxation met a column where LAPACK dgtsv pivots or is singular (flag " + std::to_string((int)f) +
                "); the B200 path does not reproduce that branch");
}
MGSolver::~MGSolver()
{
    if (tailOut) cudaFree(tailOut);
    for (auto* q : tmpRes) if (q) cudaFree(q);
    for (auto* q : cor) if (q) cudaFree(q);
    for (auto* q : res) if (q) cudaFree(q);
    if (topRes) cudaFree(topRes);
    if (topCor) cudaFree(topCor);
    if (aggRes) cudaFree(aggRes);
    if (aggCor) cudaFree(aggCor);
    if (aggBuf) cudaFree(aggBuf);
    agg.reset();     // before the ops and the context it uses
    aggTop.reset();
    aggCtx.reset();
}
void MGSolver::modifyOptionsExceptMaxDepth(const sb_mg_options& o)
{
    const int old = opt.maxDepth;
    opt           = o;
    opt.maxDepth  = old;
    if (bottom) bottom->opt = o.bottom;
    tailStart = -2;  // numCycles / prolongOrder decide whether the single-kernel tail applies
    if (agg) { sb_mg_options a = o; a.maxDepth = agg->opt.maxDepth; agg->opt = a; agg->bottom->opt = o.bottom; agg->tailStart = -2; }
}

SolverStatus MGSolver::solve(double* phi, const double* rhs, bool homog, bool setPhiToZero, double metric)
{
    // MGSolverI.H:232-262
    return cycle(opt.numCycles < 0, phi, rhs, homog, setPhiToZero, metric);
}

// MGSolver<T>::vCycle (MGSolverI.H:265-430) and ::fmg (:434-612); the two outer loops differ only
// in the call that produces the correction.
SolverStatus MGSolver::cycle(bool fmgMode, double* phi, const double* rhs, bool homog, bool setPhiToZero, double a_metric)
{
    status.clear();
    absResNorms.clear();
    std::vector<double> relResNorms;
    Op&     op  = *ops[0];
    double* r   = topRes;
    double* c   = topCor;
    if (setPhiToZero) op.setToZero(phi);
    if (a_metric > 0.0) { absResNorms.push_back(a_metric); relResNorms.push_back(1.0); }
    else if (opt.convergenceMetric > 0.0) { absResNorms.push_back(opt.convergenceMetric); relResNorms.push_back(1.0); }

    op.residual(r, phi, rhs, homog);
    absResNorms.push_back(op.norm(r, opt.normType));
    status.initResNorm = absResNorms[0];
    relResNorms.push_back(absResNorms.back() / absResNorms[0]);
    lastIters = 0;
    if (relResNorms.back() < opt.relTol) {
        status.finalResNorm = absResNorms.back();
        status.status       = SB_STATUS_CONVERGED;
        return status;
    }
    int iter;
    for (iter = 1; iter <= opt.maxIters; ++iter) {
        if (fmgMode) {
            fmg_residualEq(c, r, 0);  // allFMG = true (MGSolverI.H:444, 517)
        } else {
            vCycle_residualEq(c, r, 0, /*corIsPreCond: op.preCond(c, r, 0) fused into the first relaxation*/ true);
        }
        op.incr(phi, c, 1.0);
        op.residual(r, phi, rhs, homog);
        absResNorms.push_back(op.norm(r, opt.normType));
        relResNorms.push_back(absResNorms.back() / absResNorms[0]);
        lastIters = iter;
        if (absResNorms.back() < opt.absTol) { status.status = SB_STATUS_CONVERGED; break; }
        if (relResNorms.back() < opt.relTol) { status.status = SB_STATUS_CONVERGED; break; }
        if (relResNorms[iter] > relResNorms[iter - 1]) {
            op.incr(phi, c, -1.0);
            absResNorms.pop_back();
            relResNorms.pop_back();
            --iter;
            lastIters     = iter;
            status.status = SB_STATUS_DIVERGED;
            break;
        }
        if (relResNorms[iter] > (1.0 - opt.hang) * relResNorms[iter - 1]) { status.status = SB_STATUS_HANG; break; }
    }
    status.finalResNorm = absResNorms.back();
    if (op.relaxMethod == SB_RELAX_VERTLINE) checkPivotAll();
    return status;
}

// MGSolver<T>::vCycle_residualEq (MGSolverI.H:617-754)
void MGSolver::vCycle_residualEq(double* a_cor, const double* a_res, int depth, bool corIsPreCond)
{
    const int pre = corIsPreCond ? Op::RELAX_PRE_PRECOND : Op::RELAX_PRE_NONE;
    cudaEvent_t e0;
    Context*    pc = ops[depth]->ctx;
    if (depth == aggDepth) {  // the rest of the hierarchy runs on rank 0
        Op& dist = *ops[depth];
        pc->phaseBegin(&e0);
        aggGather(a_res, aggRes, SB_CELL, dist);
        if (!corIsPreCond) aggGather(a_cor, aggCor, SB_CELL, dist);
        if (agg) agg->vCycle_residualEq(aggCor, aggRes, 0, corIsPreCond);
        aggScatter(a_cor, aggCor, dist);
        pc->phaseEnd("agglomerated", depth, e0);
        return;
    }
    Op& op = *ops[depth];
    pc->phaseBegin(&e0);
    if (tinyTail(depth, a_cor, a_res, corIsPreCond)) {
        pc->phaseEnd("tail", depth, e0);
        return;
    }
    if (depth == opt.maxDepth) {
        op.relax(a_cor, a_res, opt.numSmoothBottom, false, pre);
        if (bottom) bottom->solve(a_cor, a_res, true, false);
        pc->phaseEnd("bottom", depth, e0);
        return;
    }
    Op&     crseOp  = *ops[depth + 1];
    double* crseCor = cor[depth + 1];
    double* crseRes = res[depth + 1];
    double* tmp     = tmpRes[depth];
    op.relax(a_cor, a_res, opt.numSmoothDown, false, pre);   // (phase opened above)
    pc->phaseEnd("relax_down", depth, e0);
    pc->phaseBegin(&e0);
    op.residual(tmp, a_cor, a_res, true);
    op.MGRestrict(crseOp, crseRes, tmp);
    pc->phaseEnd("residual_restrict", depth, e0);
    const int numCycles = std::abs(opt.numCycles);
    if (numCycles == 0) crseOp.preCond(crseCor, crseRes, 0);
    for (int i = 0; i < numCycles; ++i) vCycle_residualEq(crseCor, crseRes, depth + 1, /*crseOp.preCond(crseCor, crseRes, 0)*/ i == 0);
    pc->phaseBegin(&e0);
    const bool shiftPending = op.MGProlong(crseOp, a_cor, crseCor, opt.prolongOrder, /*deferKernel*/ true);
    pc->phaseEnd("prolong", depth, e0);
    pc->phaseBegin(&e0);
    op.relax(a_cor, a_res, opt.numSmoothUp, /*resUnchanged since the down-relax*/ opt.numSmoothDown >= 2,
             shiftPending ? Op::RELAX_PRE_SHIFT : Op::RELAX_PRE_NONE);
    pc->phaseEnd("relax_up", depth, e0);
}

// The deepest depths of the V-cycle -- as many as fit a shared-memory arena together, wholly on this rank -- as ONE single-CTA
// kernel (sb_tiny.cu: tiny_tail_k): relaxations, residual, restriction, prolongation, null-space removal, bottom smooths and the
// whole BiCGStab solve, with the solver's decisions made on the device.  The host-driven path costs ~50 launches per
// relaxation on such a depth and six host round trips per BiCGStab iteration.  SB_TINY_TAIL=0 keeps the host-driven path.
int MGSolver::tinyTailStart()
{
    if (tailStart != -2) return tailStart;
    tailStart = -1;
    static const bool allowed = [] { const char* e = getenv("SB_TINY_TAIL"); return !(e && std::string(e) == "0"); }();
    if (!allowed || !bottom || aggDepth >= 0 || std::abs(opt.numCycles) != 1 || opt.prolongOrder > 1) return tailStart;
    auto ok = [&](const Op& op) {
        if (op.ctx->nranks != 1) return false;
        if (op.relaxMethod != SB_RELAX_VERTLINE && op.relaxMethod != SB_RELAX_GSRB) return false;
        if (!k::tiny_level_fits(op.lay, op.nlocal())) return false;
        for (int d = 0; d < 3; ++d)
            for (int s = 0; s < 2; ++s) {
                const int kind = op.side[d][s].kind;
                if (!(kind < 0 || sideIsBC(kind) || kind == SIDE_PERIODIC_SELF)) return false;
            }
        return true;
    };
    static const long long maxCells = [] { const char* e = getenv("SB_TINY_CELLS"); return e ? atoll(e) : (1LL << 40); }();
    auto cells = [](const Op& op) { return (long long)op.lay.nx * op.lay.ny * op.lay.nz; };
    int d = opt.maxDepth;
    if (d < 0 || d >= (int)ops.size() || !ok(*ops[d]) || cells(*ops[d]) > maxCells) return tailStart;
    size_t bytes = k::tiny_level_bytes(ops[d]->lay, ops[d]->nlocal(), false, true);
    if (bytes > k::tiny_arena_limit()) return tailStart;
    // grow upwards while the staged copies of all levels fit the arena together
    while (d > 0 && opt.maxDepth - (d - 1) + 1 <= k::TINY_MAXLEV && ok(*ops[d - 1]) && cells(*ops[d - 1]) <= maxCells) {
        const size_t more = k::tiny_level_bytes(ops[d - 1]->lay, ops[d - 1]->nlocal(), true, false);
        if (bytes + more > k::tiny_arena_limit()) break;
        bytes += more;
        --d;
    }
    tailStart = d;
    return tailStart;
}
bool MGSolver::tinyTail(int depth, double* a_cor, const double* a_res, bool corIsPreCond)
{
    const int start = tinyTailStart();
    if (start < 0 || depth < start || ops[depth]->ctx->isProfiling()) return false;
    k::TinyTailArgs a;
    a.nlev = opt.maxDepth - depth + 1;
    for (int l = 0; l < a.nlev; ++l) {
        Op&           op = *ops[depth + l];
        k::TinyLevel& V  = a.lev[l];
        V.L = op.lay; V.c = op.coef();
        for (int d = 0; d < 3; ++d)
            for (int s = 0; s < 2; ++s) V.side[d][s] = op.side[d][s];
        V.dim = op.dim; V.relaxMethod = op.relaxMethod;
        const BoxList bl = op.boxlist();
        V.boxLo = bl.lo; V.boxHi = bl.hi; V.nboxes = bl.n;
        V.hasNullSpace = op.hasNullSpace ? 1 : 0;
        V.dv = op.dim == 2 ? op.dXi[0] * op.dXi[2] : op.dXi[0] * op.dXi[1] * op.dXi[2];
        for (int d = 0; d < 3; ++d) V.ref[d] = l + 1 < a.nlev ? op.domain.size(d) / ops[depth + l + 1]->domain.size(d) : 1;
        V.cor = l == 0 ? a_cor : cor[depth + l];
        V.res = l == 0 ? const_cast<double*>(a_res) : res[depth + l];
        V.tmp = l + 1 < a.nlev ? tmpRes[depth + l] : nullptr;
        V.pivotFlag = op.pivotFlag;
        V.lineTab   = op.relaxMethod == SB_RELAX_VERTLINE && op.lineFast ? op.lineTab : nullptr;
    }
    a.opt = bottom->opt;
    a.numSmoothDown = opt.numSmoothDown; a.numSmoothUp = opt.numSmoothUp; a.numSmoothBottom = opt.numSmoothBottom;
    a.prolongOrder = opt.prolongOrder; a.corIsPreCond = corIsPreCond ? 1 : 0;
    BiCGStabSolver& b = *bottom;
    double* w[8] = {b.r, b.r_tilde, b.e, b.p, b.p_tilde, b.s_tilde, b.t, b.v};
    for (int i = 0; i < 8; ++i) a.w[i] = w[i];
    if (!tailOut) SB_CUDA(cudaMalloc((void**)&tailOut, 16 * sizeof(double)));
    a.out = tailOut;
    size_t bytes = 0;
    for (int l = 0; l < a.nlev; ++l) bytes += k::tiny_level_bytes(ops[depth + l]->lay, ops[depth + l]->nlocal(), l + 1 < a.nlev, l + 1 == a.nlev);
    k::tiny_tail(ops[depth]->st(), a, bytes);
    static const bool dbg = [] { const char* e = getenv("SB_TINY_DEBUG"); return e && std::string(e) == "1"; }();
    if (dbg) {
        double h[10];
        SB_CUDA(cudaMemcpyAsync(h, tailOut, sizeof(h), cudaMemcpyDeviceToHost, ops[depth]->st()));
        ops[depth]->ctx->sync();
        fprintf(stderr, "tiny_tail depth %d levels %d: bottom status %d iters %d restarts %d res %.3e -> %.3e | us: stage %.1f down %.1f smooth %.1f bicgstab %.1f up %.1f\n",
                depth, a.nlev, (int)h[0], (int)h[3], (int)h[4], h[1], h[2], h[5] * 1e-3, h[6] * 1e-3, h[7] * 1e-3, h[8] * 1e-3, h[9] * 1e-3);
    }
    return true;
}

// MGSolver<T>::fmg_residualEq (MGSolverI.H:758-820)
void MGSolver::fmg_residualEq(double* a_cor, const double* a_res, int depth)
{
    if (depth == aggDepth) {
        Op& dist = *ops[depth];
        aggGather(a_res, aggRes, SB_CELL, dist);
        if (agg) agg->fmg_residualEq(aggCor, aggRes, 0);
        aggScatter(a_cor, aggCor, dist);
        return;
    }
    Op& op = *ops[depth];
    op.setToZero(a_cor);
    if (depth < opt.maxDepth) {
        Op&     crseOp  = *ops[depth + 1];
        double* crseCor = cor[depth + 1];
        double* crseRes = res[depth + 1];
        op.MGRestrict(crseOp, crseRes, a_res);
        fmg_residualEq(crseCor, crseRes, depth + 1);
        const bool shiftPending = op.MGProlong(crseOp, a_cor, crseCor, opt.prolongOrderFMG, /*deferKernel*/ true);
        op.relax(a_cor, a_res, opt.numSmoothUpFMG, false, shiftPending ? Op::RELAX_PRE_SHIFT : Op::RELAX_PRE_NONE);
    }
    const int numCycles = std::abs(opt.numCycles);
    for (int i = 0; i < numCycles; ++i) vCycle_residualEq(a_cor, a_res, depth);
}

// ---------------------------------------------------------------------------------------------
// LevelHybridSolver (LevelHybridSolver.cpp).
int HybridSolver::computeSolveMode(const Op& op)
{
    // :457-498: lepticity = min(dXi_x, dXi_y) / L_z
    const double Lz   = (double)op.domain.size(2) * op.dXi[2];
    const double dh   = op.dim == 2 ? std::min(op.dXi[0], op.dXi[0]) : std::min(op.dXi[0], op.dXi[1]);
    const double lept = dh / Lz;
    if (lept > 1.0) return SB_MODE_LEPTIC;
    if (lept > 0.2) return SB_MODE_LEPTIC_MG;
    return SB_MODE_MG;
}
// LevelHybridSolver::define (LevelHybridSolver.cpp:127-186)
void HybridSolver::define(Op& top, const sb_mg_options& o)
{
    op   = &top;
    opt  = o;
    optDefine = o;
    mode = isHybrid ? computeSolveMode(top) : SB_MODE_MG;
    const bool useLeptic = mode == SB_MODE_LEPTIC || mode == SB_MODE_LEPTIC_MG;
    const bool useMG     = mode == SB_MODE_MG || mode == SB_MODE_LEPTIC_MG;
    if (useLeptic) {
        leptic.reset(new LepticSolver);
        leptic->define(top, o);
    }
    if (useMG) {
        mg.define(top, o, {}, true);
        opt.maxDepth = mg.opt.maxDepth;
    } else {
        opt.maxDepth = -1;  // the MG options keep proj.maxDepth: no MGSolver is defined in pure leptic mode
        mg.opt       = opt;
    }
    cor = top.alloc();
    res = top.alloc();
    if (mode == SB_MODE_LEPTIC_MG) localRes = top.alloc();
}
HybridSolver::~HybridSolver()
{
    if (cor) cudaFree(cor);
    if (res) cudaFree(res);
    if (localRes) cudaFree(localRes);
    if (pDiv) cudaFree(pDiv);
    if (pPhi) cudaFree(pPhi);
    for (double* q : pGrad) if (q) cudaFree(q);
}
void HybridSolver::setQuickAndDirty(bool on)
{
    if (on == quickAndDirty) return;
    if (on) {
        savedOpt = opt; savedMgOpt = mg.opt; savedSwaps = maxSolverSwaps;
        sb_mg_options q = optDefine;  // getDefaultOptions(): the proj.* values this solver was defined with
        q.absTol = 1.0e-300; q.relTol = 1.0e-300; q.numCycles = -1; q.maxIters = 1; q.verbosity = 0;
        q.bottom.absTol = 1.0e-300; q.bottom.relTol = 1.0e-300; q.bottom.verbosity = 0;
        if (!mg.ops.empty()) mg.modifyOptionsExceptMaxDepth(q);
        const int md = opt.maxDepth;
        opt = optDefine; opt.maxDepth = md;
        opt.absTol = 1.0e-300; opt.relTol = 1.0e-2;
        maxSolverSwaps = 1;
    } else {
        if (!mg.ops.empty()) mg.modifyOptionsExceptMaxDepth(savedMgOpt);
        opt = savedOpt;
        maxSolverSwaps = savedSwaps;
    }
    quickAndDirty = on;
}
void HybridSolver::projectTemps()
{
    if (pDiv) return;
    pDiv = op->alloc(); pPhi = op->alloc();
    for (int d = 0; d < 3; ++d)
        if (!(op->dim == 2 && d == 1)) pGrad[d] = op->alloc();
}
SolverStatus HybridSolver::projectCorrect(double* const vel[3], double* p, double projDt, int velGhost, double* phiOut, double* initDivNorm,
                                          double* finalDivNorm)
{
    Op& o = *op;
    projectTemps();
    if (velGhost >= 0) o.scaleVelocity(vel, velGhost, true);                  // sendToAdvectingVelocity      :290
    o.levelDivergence(pDiv, vel);                                             //                              :293
    const double n0 = o.norm(pDiv, opt.normType);                             //                              :297
    if (initDivNorm) *initDivNorm = n0;
    SolverStatus st = solve(pPhi, pDiv, true, true, -1.0);                    //                              :316
    o.levelGradient(pGrad, pPhi, true);                                       //                              :328
    for (int d = 0; d < 3; ++d)
        if (pGrad[d]) k::incr_valid(o.st(), o.lay, vel[d], pGrad[d], -1.0, d);  // vel.plus(gradPhi, -1.0)     :331-336
    const double smallReal = 1.0e4 * std::numeric_limits<double>::epsilon();
    if (p && !(std::abs(projDt) < smallReal)) o.incr(p, pPhi, 1.0 / projDt);  // p.plus(phi, 1 / projDt)      :340-347
    if (finalDivNorm) {                                                       //                              :354-357
        o.levelDivergence(pDiv, vel);
        *finalDivNorm = o.norm(pDiv, opt.normType);
    }
    if (phiOut) o.assignLocal(phiOut, pPhi);
    if (velGhost >= 0) o.scaleVelocity(vel, velGhost, false);                 // sendToCartesianVelocity      :372
    return st;
}
SolverStatus HybridSolver::projectPredict(double* const vel[3], double* p, double projDt, int velGhost, double norms[3], bool* usedFallback)
{
    Op& o = *op;
    projectTemps();
    SolverStatus st;
    if (velGhost >= 0) o.scaleVelocity(vel, velGhost, true);                  //                              :126
    o.levelDivergence(pDiv, vel);                                             //                              :137
    norms[0] = o.norm(pDiv, opt.normType);
    o.levelGradient(pGrad, p, false);                                         // levelGradient(gradP, p, crseP, t, false, false) :147
    for (int d = 0; d < 3; ++d)
        if (pGrad[d]) k::incr_valid(o.st(), o.lay, vel[d], pGrad[d], -projDt, d);  // vel.plus(gradP, -projDt)  :155-160
    o.levelDivergence(pDiv, vel);                                             //                              :170
    norms[1] = o.norm(pDiv, opt.normType);
    norms[2] = -1.0;
    if (usedFallback) *usedFallback = false;
    if (norms[1] > norms[0]) {                                                // the lagged pressure made it worse :176-186
        setQuickAndDirty(true);
        st = projectCorrect(vel, p, projDt, -1, nullptr, nullptr, nullptr);
        setQuickAndDirty(false);
        o.levelDivergence(pDiv, vel);
        norms[2] = o.norm(pDiv, opt.normType);
        if (usedFallback) *usedFallback = true;
    }
    if (velGhost >= 0) o.scaleVelocity(vel, velGhost, false);                 //                              :241
    return st;
}
SolverStatus HybridSolver::solve(double* phi, const double* rhs, bool homog, bool setPhiToZero, double metric)
{
    if (!isHybrid) return mg.solve(phi, rhs, homog, setPhiToZero, metric);
    // LevelHybridSolver::solve (:266-292) + solveResidualEq (:296-452)
    Op& o = *op;
    if (setPhiToZero) o.setToZero(phi);
    o.residual(res, phi, rhs, homog);
    o.setToZero(cor);
    resNorms.clear();
    resNorms.push_back(o.norm(res, opt.normType));
    if (metric > 0.0) resNorms[0] = metric;
    SolverStatus st;
    st.initResNorm = resNorms.back();
    if (mode == SB_MODE_LEPTIC) {
        st = leptic->solve(cor, res, true, false);  // replaces the status wholesale, as the reference's assignment does
        resNorms.insert(resNorms.end(), leptic->resNorms.begin() + 1, leptic->resNorms.end());
    } else if (mode == SB_MODE_LEPTIC_MG) {
        for (int swaps = 0; swaps < maxSolverSwaps; ++swaps) {
            const double preLepticResNorm = resNorms.back();
            leptic->solve(cor, res, true, false);
            resNorms.insert(resNorms.end(), leptic->resNorms.begin() + 1, leptic->resNorms.end());
            if (resNorms.back() <= opt.absTol) { st.status = SB_STATUS_CONVERGED; break; }
            if (resNorms.back() <= opt.relTol * resNorms[0]) { st.status = SB_STATUS_CONVERGED; break; }
            mg.vCycle_residualEq(cor, res, 0);
            o.residual(localRes, cor, res, true);
            resNorms.push_back(o.norm(localRes, opt.normType));
            if (resNorms.back() <= opt.absTol) { st.status = SB_STATUS_CONVERGED; break; }
            if (resNorms.back() <= opt.relTol * resNorms[0]) { st.status = SB_STATUS_CONVERGED; break; }
            if (resNorms.back() > preLepticResNorm) { st.status = SB_STATUS_DIVERGED; break; }
            if (swaps == maxSolverSwaps - 1) { st.status = SB_STATUS_MAXITERS; break; }
        }
        if (o.relaxMethod == SB_RELAX_VERTLINE) mg.checkPivotAll();
    } else {
        st = mg.solve(cor, res, true, false, -1.0);
        resNorms.push_back(st.finalResNorm);
        st.initResNorm = resNorms[0];
    }
    st.finalResNorm = resNorms.back();
    o.incr(phi, cor, 1.0);
    return st;
}

}  // namespace sb
