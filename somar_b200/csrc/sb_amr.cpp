// sb_amr.cpp -- the AMR half of the projection's operator and the composite solver on top of it:
//   * PoissonOp::{applyBCs with a coarser level, AMROperator, AMROperatorNF, AMROperatorNC, AMRNormLevel, getFlux,
//     reflux (both forms), compDivergence}                        Grade3_Calculus/Elliptic/PoissonOp.cpp:726-763, 1156-1478, 1618-1631
//   * CFInterp::interpAtCFI -> MappedQuadCFInterp::coarseFineInterp    Grade3_Calculus/CFInterp.cpp:387-416,
//                                                                  Grade2_AnisotropicChombo/QuadCFInterp/MappedQuadCFInterp.cpp:578-622
//   * AnisotropicFluxRegister (define, incrementCoarse, incrementFine, reflux)   Grade2_AnisotropicChombo/AnisotropicFluxRegister.cpp
//   * AMRHybridSolver (define, solve, amrVCycle_residualEq, computeAMRResidual[Level])  Elliptic/AMRHybridSolver.cpp:52-760
//
// Layout.  A refined level is a rectangular patch of boxes; each rank owns a rectangle of it (its tile) as one fused
// array, exactly like a base level (sb_core.h: Lay).  Tile sides that border the coarser level are SIDE_CF.  Per
// refined level and rank, a CFLink holds what couples it to the coarser level:
//   crseBuf  coarse data under and around the tile (coarsened tile grown by 2; MappedQuadCFInterp::m_coarBuffer)
//   cfA      an array over the coarsened tile (AMRHybridSolver::m_vResC: restricted residual / coarse correction)
//   cfReg    the fine side of the flux register: sums of fine-face fluxes in the ghost faces of the coarsened tile
//   records  per coarse cell under a ghost face, the weights of its tangential derivative stencils (sb_amr_plan.cpp)
//   copiers  coarse level -> crseBuf, cfA <-> coarse level, cfReg -> coarse level (LevelCopier: local device copies,
//            NCCL send/recv for regions owned by another rank)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "sb_amr.h"
#include "sb_comm.h"
#include "sb_host.h"

namespace sb {

// ---------------------------------------------------------------------------------------------
namespace {
Box3 intersect(const Box3& a, const Box3& b)
{
    Box3 c;
    for (int d = 0; d < 3; ++d) { c.lo[d] = std::max(a.lo[d], b.lo[d]); c.hi[d] = std::min(a.hi[d], b.hi[d]); }
    return c;
}
bool empty(const Box3& b) { return b.hi[0] < b.lo[0] || b.hi[1] < b.lo[1] || b.hi[2] < b.lo[2]; }
// tiles of every rank from a box list (bounding boxes, as planDecomposition builds them)
std::vector<Box3> tilesOf(const std::vector<Box3>& boxes, const std::vector<int>& rank, int nr)
{
    std::vector<Box3> t(nr, Box3{{0, 0, 0}, {-1, -1, -1}});
    std::vector<int>  cnt(nr, 0);
    for (size_t b = 0; b < boxes.size(); ++b) {
        Box3& q = t[rank[b]];
        if (cnt[rank[b]]++ == 0) q = boxes[b];
        else
            for (int i = 0; i < 3; ++i) { q.lo[i] = std::min(q.lo[i], boxes[b].lo[i]); q.hi[i] = std::max(q.hi[i], boxes[b].hi[i]); }
    }
    return t;
}
}  // namespace

void LevelCopier::define(int rank, const std::vector<std::vector<Box3>>& src, const std::vector<std::vector<Box3>>& dst)
{
    items.clear();
    sendLen = recvLen = 0;
    const int nr = (int)src.size();
    for (int rs = 0; rs < nr; ++rs)
        for (int rd = 0; rd < nr; ++rd)
            for (const Box3& sb : src[rs])
                for (const Box3& db : dst[rd]) {
                    const Box3 x = intersect(sb, db);
                    if (empty(x)) continue;
                    Item it{rs, rd, x, 0};
                    if (rs != rd) {
                        if (rs == rank) { it.off = sendLen; sendLen += (size_t)x.numPts(); }
                        if (rd == rank) { it.off = recvLen; recvLen += (size_t)x.numPts(); }
                    }
                    if (rs == rank || rd == rank) items.push_back(it);
                }
    if (sendBuf) { cudaFree(sendBuf); sendBuf = nullptr; }
    if (recvBuf) { cudaFree(recvBuf); recvBuf = nullptr; }
    if (sendLen) SB_CUDA(cudaMalloc((void**)&sendBuf, sendLen * sizeof(double)));
    if (recvLen) SB_CUDA(cudaMalloc((void**)&recvBuf, recvLen * sizeof(double)));
}
LevelCopier::~LevelCopier()
{
    if (sendBuf) cudaFree(sendBuf);
    if (recvBuf) cudaFree(recvBuf);
}
void LevelCopier::exec(Context* ctx, const Lay& srcLay, const double* src, const Lay& dstLay, double* dst, int mode, double scale)
{
    const int me = ctx->rank;
    std::vector<Comm::Msg> msgs;
    for (const Item& it : items) {
        if (it.srcRank == me && it.dstRank == me) {
            k::copy_region(ctx->st, srcLay, src, dstLay, dst, it.box, mode, scale);
        } else if (it.srcRank == me) {
            k::stage_region(ctx->st, srcLay, const_cast<double*>(src), sendBuf + it.off, it.box, -1, 1.0);
            msgs.push_back({sendBuf + it.off, (size_t)it.box.numPts(), it.dstRank, true});
        } else {
            msgs.push_back({recvBuf + it.off, (size_t)it.box.numPts(), it.srcRank, false});
        }
    }
    if (msgs.empty()) return;
    if (!ctx->comm) SB_FAIL("inter-level copy between ranks needs sb_comm_init");
    ctx->comm->sendRecv(msgs);
    for (const Item& it : items)
        if (it.dstRank == me && it.srcRank != me) k::stage_region(ctx->st, dstLay, dst, recvBuf + it.off, it.box, mode, scale);
}

// ---------------------------------------------------------------------------------------------
struct CFLink {
    int               ref[3];
    Box3              ctile, bufBox;
    Lay               bufLay, cfLay;
    std::vector<Box3> crseTiles;   // tile of every rank on the coarser level
    Lay               crseLay;     // this rank's layout on the coarser level (what a coarse field must have)
    double *          crseBuf = nullptr, *cfA = nullptr, *cfReg = nullptr;
    double*           wdev = nullptr;  // all records of this rank, back to back
    bool              hasSide[3][2] = {};
    CFSideParams      sp[3][2];
    LevelCopier       toBuf, cfToCrse, crseToCf, regToCrse;
    ~CFLink()
    {
        cudaFree(crseBuf); cudaFree(cfA); cudaFree(cfReg); cudaFree(wdev);
    }
};

// m_cfInterp.define(m_grids, m_dXi, m_crseAMRGrids) (PoissonOp.cpp:118-120) -> MappedQuadCFInterp::define
// (MappedQuadCFInterp.cpp:68-217), plus what AMRHybridSolver::initializeSolve builds per solve (the coarsened holder
// and its two copiers, AMRHybridSolver.cpp:456-468) and the fine half of AnisotropicFluxRegister::define.
void Op::defineCF()
{
    if (!refined || depth != 0 || flatZ) return;
    cf.reset(new CFLink);
    CFLink&   L  = *cf;
    const int nr = ctx->nranks, me = ctx->rank;
    for (int d = 0; d < 3; ++d) L.ref[d] = crseRef[d];
    for (const Box3& b : boxes)
        if (!coarsenable(b, L.ref)) SB_FAIL("the boxes of a refined level must be coarsenable by the AMR refinement ratio");
    L.crseTiles = tilesOf(crseGrids.boxes, crseGrids.rank, nr);
    for (int r = 0; r < nr; ++r)
        if (empty(L.crseTiles[r])) SB_FAIL("every rank must own a tile of the coarser AMR level");
    L.crseLay = makeLay(L.crseTiles[me]);
    auto growBuf = [&](const Box3& t) {
        Box3 b = coarsen(t, L.ref);
        for (int d = 0; d < 3; ++d) { if (dim == 2 && d == 1) continue; b.lo[d] -= 2; b.hi[d] += 2; }
        return b;
    };
    L.ctile  = coarsen(tile, L.ref);
    L.bufBox = growBuf(tile);
    L.bufLay = makeLay(L.bufBox);
    L.cfLay  = makeLay(L.ctile);
    SB_CUDA(cudaMalloc((void**)&L.crseBuf, L.bufLay.n * sizeof(double)));
    SB_CUDA(cudaMemset(L.crseBuf, 0, L.bufLay.n * sizeof(double)));
    SB_CUDA(cudaMalloc((void**)&L.cfA, L.cfLay.n * sizeof(double)));
    SB_CUDA(cudaMemset(L.cfA, 0, L.cfLay.n * sizeof(double)));
    SB_CUDA(cudaMalloc((void**)&L.cfReg, L.cfLay.n * sizeof(double)));
    SB_CUDA(cudaMemset(L.cfReg, 0, L.cfLay.n * sizeof(double)));

    // side kinds of every rank's tile on this level (for the register slabs they send)
    std::vector<std::vector<Box3>> crseSrc(nr), bufDst(nr), cfBoxes(nr), regSrc(nr);
    for (int r = 0; r < nr; ++r) {
        crseSrc[r].push_back(L.crseTiles[r]);
        bufDst[r].push_back(intersect(growBuf(tiles[r]), crseGrids.domain));
        const Box3 ct = coarsen(tiles[r], L.ref);
        cfBoxes[r].push_back(ct);
        std::vector<Box3> tl;
        std::vector<int>  loc;
        SideBC            sd[3][2];
        planDecomposition(boxes, boxRank, domain, periodic, r, nr, tl, loc, sd, true);
        for (int d = 0; d < 3; ++d)
            for (int s = 0; s < 2; ++s) {
                if (sd[d][s].kind != SIDE_CF || (dim == 2 && d == 1)) continue;
                Box3 slab = ct;
                slab.lo[d] = slab.hi[d] = s ? ct.hi[d] + 1 : ct.lo[d] - 1;
                slab = intersect(slab, crseGrids.domain);
                if (!empty(slab)) regSrc[r].push_back(slab);
            }
    }
    L.toBuf.define(me, crseSrc, bufDst);
    L.cfToCrse.define(me, cfBoxes, crseSrc);
    L.crseToCf.define(me, crseSrc, cfBoxes);
    L.regToCrse.define(me, regSrc, crseSrc);

    // stencil records of this rank's coarse-fine sides
    std::vector<double> w;
    size_t              off[3][2][3] = {};
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s) {
            if (side[d][s].kind != SIDE_CF || (dim == 2 && d == 1)) continue;
            L.hasSide[d][s] = true;
            CFSideParams& P = L.sp[d][s];
            std::memset(&P, 0, sizeof(P));
            P.dir = d; P.side = s;
            int tr[2] = {-1, -1}, nt = 0;
            for (int t = 0; t < 3; ++t)
                if (t != d && !(dim == 2 && t == 1)) tr[nt++] = t;
            P.t0 = tr[0]; P.t1 = nt == 2 ? tr[1] : -1;
            const int nf[3] = {lay.nx, lay.ny, lay.nz};
            const int nc[3] = {L.ctile.size(0), L.ctile.size(1), L.ctile.size(2)};
            P.nf0 = nf[P.t0]; P.nf1 = P.t1 >= 0 ? nf[P.t1] : 1; P.nfn = nf[d];
            P.nc0 = nc[P.t0]; P.nc1 = P.t1 >= 0 ? nc[P.t1] : 1; P.ncn = nc[d];
            P.r0 = L.ref[P.t0]; P.r1 = P.t1 >= 0 ? L.ref[P.t1] : 1; P.rn = L.ref[d];
            for (int i = 0; i < 3; ++i) {
                P.flo[i] = tile.lo[i]; P.clo[i] = L.ctile.lo[i]; P.blo[i] = L.bufBox.lo[i];
                P.dxf[i] = dXi[i]; P.dxc[i] = dXi[i] * (double)L.ref[i];  // m_dxFine * RealVect(m_refRatio)
            }
            P.cn = s ? L.ctile.hi[d] + 1 : L.ctile.lo[d] - 1;
            P.B  = L.bufLay;
            if (P.nfn < 2) SB_FAIL("a refined tile needs at least 2 cells normal to a coarse-fine side");
            const size_t ncell = (size_t)P.nc0 * P.nc1;
            off[d][s][0] = w.size(); w.resize(w.size() + 10 * ncell, 0.0);
            off[d][s][1] = w.size(); w.resize(w.size() + 10 * ncell, 0.0);
            off[d][s][2] = w.size(); w.resize(w.size() + 9 * ncell, 0.0);
            std::vector<char> seen(ncell, 0);
            for (int lb : local) {
                const Box3& b = boxes[lb];
                if ((s ? b.hi[d] != tile.hi[d] : b.lo[d] != tile.lo[d])) continue;
                std::vector<int>    cells;
                std::vector<double> w1, w2, wm;
                planCFStencils(crseGrids.domain, periodic, L.ref, boxes, lb, d, s, cells, w1, w2, wm, dim, &crseGrids.boxes);
                for (size_t n = 0; n < cells.size() / 3; ++n) {
                    const int c[3] = {cells[3 * n], cells[3 * n + 1], cells[3 * n + 2]};
                    if (c[d] != P.cn) SB_FAIL("coarse-fine stencil record off the ghost slab");
                    const int    ia = c[P.t0] - L.ctile.lo[P.t0], ib = P.t1 >= 0 ? c[P.t1] - L.ctile.lo[P.t1] : 0;
                    if (ia < 0 || ia >= P.nc0 || ib < 0 || ib >= P.nc1) SB_FAIL("coarse-fine stencil record outside the tile side");
                    const size_t r = (size_t)ia + (size_t)P.nc0 * ib;
                    seen[r] = 1;
                    std::copy(w1.begin() + 10 * n, w1.begin() + 10 * n + 10, w.begin() + off[d][s][0] + 10 * r);
                    std::copy(w2.begin() + 10 * n, w2.begin() + 10 * n + 10, w.begin() + off[d][s][1] + 10 * r);
                    std::copy(wm.begin() + 9 * n, wm.begin() + 9 * n + 9, w.begin() + off[d][s][2] + 9 * r);
                }
            }
            for (char c : seen)
                if (!c) SB_FAIL("coarse-fine side with ghost cells that have no coarse cell under them (patch not rectangular?)");
        }
    if (!w.empty()) {
        SB_CUDA(cudaMalloc((void**)&L.wdev, w.size() * sizeof(double)));
        SB_CUDA(cudaMemcpy(L.wdev, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice));
        for (int d = 0; d < 3; ++d)
            for (int s = 0; s < 2; ++s)
                if (L.hasSide[d][s]) {
                    L.sp[d][s].w1 = L.wdev + off[d][s][0];
                    L.sp[d][s].w2 = L.wdev + off[d][s][1];
                    L.sp[d][s].wm = L.wdev + off[d][s][2];
                }
    }
}

namespace {
void checkCrse(const Op& fine, const Op& crse)
{
    if (!fine.cf) SB_FAIL("this operator has no coarser AMR level (define it with crse_box_* in sb_level_desc)");
    if (!(crse.tile == fine.cf->crseTiles[fine.ctx->rank]) || !(crse.domain == fine.crseGrids.domain))
        SB_FAIL("the coarse field does not live on the coarser AMR level this operator was defined with");
}
}  // namespace

// CFInterp::interpAtCFI(fine, crse) (CFInterp.cpp:387-416) -> MappedQuadCFInterp::coarseFineInterp
// (MappedQuadCFInterp.cpp:578-622): copy the coarse data into the buffer, then every coarse-fine side.
void Op::interpAtCFI(double* phi, const Op& crseOp, const double* crsePhi)
{
    checkCrse(*this, crseOp);
    CFLink& L = *cf;
    L.toBuf.exec(ctx, crseOp.lay, crsePhi, L.bufLay, L.crseBuf);
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s)
            if (L.hasSide[d][s]) k::cf_interp(st(), L.sp[d][s], lay, phi, L.crseBuf);
}

// PoissonOp::applyBCs (PoissonOp.cpp:726-763) with a coarser level: exchange, coarse-fine ghosts (homogeneous
// formula or interpolation from the coarse data), physical BCs.  The three fills touch disjoint ghost cells.
void Op::applyBCsAMR(double* phi, const Op* crseOp, const double* crsePhi, bool homogCFI)
{
    applyBCs(phi, true);  // exchange + homogeneous CF formula on SIDE_CF + physical
    if (!homogCFI && refined) {
        if (!crseOp || !crsePhi) SB_FAIL("inhomogeneous coarse-fine BCs need the coarse level's data");
        interpAtCFI(phi, *crseOp, crsePhi);
    }
}

// PoissonOp::AMROperatorNF (PoissonOp.cpp:1178-1187)
void Op::AMROperatorNF(double* lhs, double* phi, const Op& crseOp, const double* crsePhi)
{
    applyBCsAMR(phi, &crseOp, crsePhi, false);
    k::apply_op(st(), lay, coef(), lhs, phi);
}
// PoissonOp::AMROperatorNC (PoissonOp.cpp:1195-1207)
void Op::AMROperatorNC(double* lhs, Op& fineOp, double* finePhi, double* phi)
{
    applyBCs(phi, true);
    k::apply_op(st(), lay, coef(), lhs, phi);
    reflux(lhs, fineOp, finePhi, phi);
}
// PoissonOp::AMROperator (PoissonOp.cpp:1156-1169)
void Op::AMROperator(double* lhs, Op& fineOp, double* finePhi, double* phi, const Op& crseOp, const double* crsePhi)
{
    applyBCsAMR(phi, &crseOp, crsePhi, false);
    k::apply_op(st(), lay, coef(), lhs, phi);
    reflux(lhs, fineOp, finePhi, phi);
}
// AMRMGOperator::AMRResidual / NF / NC (AMRMGOperator.H:107-175): L, then scale(-1), then incr(rhs, 1)
void Op::AMRResidual(double* res, Op* fineOp, double* finePhi, double* phi, const Op* crseOp, const double* crsePhi, const double* rhs)
{
    if (refined && !crseOp) SB_FAIL("AMRResidual on a refined level needs the coarser level");
    applyBCsAMR(phi, crseOp, crsePhi, !refined);
    if (!fineOp) {  // (-L) + rhs is rhs - L to the bit
        k::residual(st(), lay, coef(), res, phi, rhs);
        return;
    }
    k::apply_op(st(), lay, coef(), res, phi);
    reflux(res, *fineOp, finePhi, phi);
    k::axby_valid(st(), lay, res, res, rhs, -1.0, 1.0);
}

// PoissonOp::getFlux (PoissonOp.cpp:1295-1326) on every face of the level (the reference calls it per box and face box)
void Op::getFlux(double* const flux[3], const double* phi)
{
    for (int d = 0; d < 3; ++d) {
        if (dim == 2 && d == 1) continue;
        k::gradient(st(), lay, flux[d], phi, Jgup[d], d, 1.0 / dXi[d], beta, 1);
    }
}

namespace {
// the cells of the coarse tile just outside side (d, s) of the finer level's patch, in tile-local indices
bool coarseSlab(const Op& crse, const Op& fine, int d, int s, int lo[3], int n[3])
{
    const int* ref = fine.crseRef;
    const Box3 cp  = coarsen(fine.patch, ref);
    Box3       slab = cp;
    slab.lo[d] = slab.hi[d] = s ? cp.hi[d] + 1 : cp.lo[d] - 1;
    const Box3 x = intersect(intersect(slab, crse.domain), crse.tile);
    if (empty(x)) return false;
    for (int i = 0; i < 3; ++i) { lo[i] = x.lo[i] - crse.tile.lo[i]; n[i] = x.size(i); }
    return true;
}
bool patchSideIsCF(const Op& fine, int d, int s)
{
    // a side of the whole patch borders the coarser level unless it lies on the domain boundary (a patch that spans a
    // periodic direction has no side there at all)
    if (fine.dim == 2 && d == 1) return false;
    const bool atDom = s ? fine.patch.hi[d] == fine.domain.hi[d] : fine.patch.lo[d] == fine.domain.lo[d];
    if (!atDom) return true;
    if (!fine.periodic[d]) return false;
    return !(fine.patch.lo[d] == fine.domain.lo[d] && fine.patch.hi[d] == fine.domain.hi[d]);
}
void refluxImpl(Op& crse, double* res, Op& fine, double* finePhi, const double* phi, double* const flux[3], double* const fineFlux[3])
{
    if (!fine.cf) SB_FAIL("the finer operator has no coarser AMR level");
    checkCrse(fine, crse);
    CFLink&      L  = *fine.cf;
    cudaStream_t st = crse.st();
    // 2. coarse side (incrementCoarse) and the first half of reflux: res += -1 * coarseRegister
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s) {
            if (!patchSideIsCF(fine, d, s)) continue;
            int lo[3], n[3];
            if (!coarseSlab(crse, fine, d, s, lo, n)) continue;
            k::reflux_coarse(st, crse.lay, res, phi, crse.Jgup[d], flux ? flux[d] : nullptr, d, s, lo, n, 1.0 / crse.dXi[d], crse.beta);
        }
    // 3. fine side: ghosts at the coarse-fine interface, fine fluxes, sums per coarse cell
    if (finePhi) fine.interpAtCFI(finePhi, crse, phi);
    bool any = false;
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s) {
            if (!L.hasSide[d][s]) continue;
            const double sign   = s ? 1.0 : -1.0;
            const double aScale = 1.0 / crse.dXi[d];  // "Yes, the coarse dXi!" (PoissonOp.cpp:1405-1406)
            const double denom  = (double)((L.ref[0] * L.ref[1] * L.ref[2]) / L.ref[d]);
            const double scale  = sign * aScale / denom;
            k::fine_register(st, L.sp[d][s], fine.lay, L.cfLay, L.cfReg, finePhi, fine.Jgup[d], fineFlux ? fineFlux[d] : nullptr,
                             1.0 / fine.dXi[d], fine.beta, scale);
            any = true;
        }
    (void)any;
    // 4. second half of reflux: res += fineRegister * (-1) through the reverse copier
    L.regToCrse.exec(crse.ctx, L.cfLay, L.cfReg, crse.lay, res, 1, -1.0);
}
}  // namespace

// PoissonOp::reflux(res, finePhi, phi, fineRefRatio, finerOp) (PoissonOp.cpp:1333-1418)
void Op::reflux(double* res, Op& fineOp, double* finePhi, const double* phi)
{
    refluxImpl(*this, res, fineOp, finePhi, phi, nullptr, nullptr);
}
// PoissonOp::reflux(div, flux, fineFlux) (PoissonOp.cpp:1425-1477)
void Op::refluxFlux(double* div, double* const flux[3], Op& fineOp, double* const fineFlux[3])
{
    refluxImpl(*this, div, fineOp, nullptr, nullptr, flux, fineFlux);
}
// PoissonOp::compDivergence (PoissonOp.cpp:1618-1631)
void Op::compDivergence(double* div, double* const flux[3], Op* fineOp, double* const fineFlux[3])
{
    levelDivergence(div, flux);
    if (fineOp) refluxFlux(div, flux, *fineOp, fineFlux);
}

void Op::averageDownTo(Op& crseOp, double* crse, const double* fine)
{
    checkCrse(*this, crseOp);
    CFLink& L = *cf;
    k::restrict_avg(st(), lay, L.cfLay, L.ref, L.cfA, fine);        // localCoarsen -> CFINTERP_UNMAPPEDAVERAGE
    L.cfToCrse.exec(ctx, L.cfLay, L.cfA, crseOp.lay, crse);         // localCrse.copyTo(a_crse, m_crseToUserCopier)
}

// PoissonOp::AMRNormLevel (PoissonOp.cpp:1225-1286): per box, the cells under the finer level count as zero and
// numPts stays the box's (FArrayBox::norm(validBox, p)); powScale = prod(dXi).
double Op::AMRNormLevel(const double* x, const Op* fineOp, int p)
{
    double powScale = 1.0;  // RealVect::product (D_TERM order)
    powScale        = dim == 2 ? dXi[0] * dXi[2] : dXi[0] * dXi[1] * dXi[2];
    if (!fineOp) return norm(x, p, powScale);
    if (p < 0 || p > 2) SB_FAIL("norm type must be 0, 1 or 2");
    const Box3 cp = coarsen(fineOp->patch, fineOp->crseRef);
    Box3       m;
    for (int d = 0; d < 3; ++d) { m.lo[d] = cp.lo[d] - tile.lo[d]; m.hi[d] = cp.hi[d] - tile.lo[d]; }
    const int nl = nlocal();
    k::reduce_boxes(st(), lay, boxlist(), p, x, nullptr, 0.0, redPartial, redOut, &m);
    SB_CUDA(cudaMemcpyAsync(ctx->hpin, redOut, nl * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    ctx->sync();
    double ret = 0.0;
    for (int b = 0; b < nl; ++b) {
        const double numPts = (double)boxes[local[b]].numPts();
        double       boxVal;
        if (p == 0) boxVal = ctx->hpin[b];
        else if (p == 1) boxVal = ctx->hpin[b] / numPts;
        else boxVal = std::sqrt(ctx->hpin[b] / numPts);
        if (p == 0) ret = std::max(ret, boxVal);
        else ret += std::pow(boxVal, p);
    }
    if (p == 0) ctx->allreduceMax(&ret, 1);
    else {
        ctx->allreduceSum(&ret, 1);
        ret = std::pow(ret * powScale, 1.0 / (double)p);
    }
    return ret;
}

// ---------------------------------------------------------------------------------------------
// Elliptic::AMRHybridSolver (Elliptic/AMRHybridSolver.cpp).
void AMRSolver::define(const std::vector<Op*>& a_ops, int a_lmin, int a_lmax, const sb_mg_options& o)
{
    if (a_lmax < a_lmin || a_lmin < 0 || a_lmax >= (int)a_ops.size()) SB_FAIL("AMRHybridSolver::define: bad lmin / lmax");
    opt   = o;
    lmin  = a_lmin; lmax = a_lmax;
    lbase = lmin > 0 ? lmin - 1 : lmin;
    ops   = a_ops;
    for (int l = lbase; l <= lmax; ++l) {
        if (!ops[l]) SB_FAIL("AMRHybridSolver::define: missing operator");
        if (l > lbase) {
            if (!ops[l]->refined || !ops[l]->cf) SB_FAIL("AMRHybridSolver::define: level without a coarser AMR level");
            checkCrse(*ops[l], *ops[l - 1]);
        }
    }
    // one LevelHybridSolver per level, LevelHybridSolver::getDefaultOptions() (the proj.* parameters) (:87-100)
    hybrid.clear();
    hybrid.resize(lmax + 1);
    for (int l = lmin; l <= lmax; ++l) {
        hybrid[l].reset(new HybridSolver);
        hybrid[l]->isHybrid = true;
        hybrid[l]->define(*ops[l], o);
    }
    ve.assign(lmax + 1, nullptr); vr.assign(lmax + 1, nullptr); scratch.assign(lmax + 1, nullptr);
    for (int l = lbase; l <= lmax; ++l) ve[l] = ops[l]->alloc();
    for (int l = lmin; l <= lmax; ++l) { vr[l] = ops[l]->alloc(); scratch[l] = ops[l]->alloc(); }
}
AMRSolver::~AMRSolver()
{
    for (auto* v : {&ve, &vr, &scratch})
        for (double* q : *v)
            if (q) cudaFree(q);
}

// AMRHybridSolver::computeAMRResidualLevel (:567-650)
void AMRSolver::computeAMRResidualLevel(double* res, double* finePhi, double* phi, const double* crsePhi, const double* rhs, int lev)
{
    Op*       fineOp = nullptr;
    const Op* crseOp = nullptr;
    if (lmax != lmin) {
        if (lev == lmax) crseOp = ops[lev - 1];                             // nothing above, something below
        else if (lev == lmin) {
            fineOp = ops[lev + 1];
            if (lev != 0) crseOp = ops[lev - 1];
        } else { fineOp = ops[lev + 1]; crseOp = ops[lev - 1]; }
    } else if (lev != 0) crseOp = ops[lev - 1];
    ops[lev]->AMRResidual(res, fineOp, fineOp ? finePhi : nullptr, phi, crseOp, crseOp ? crsePhi : nullptr, rhs);
}

// AMRHybridSolver::computeAMRResidual (:507-558)
double AMRSolver::computeAMRResidual(std::vector<double*>& res, const std::vector<double*>& phi, const std::vector<const double*>& rhs,
                                     bool computeNorm)
{
    double    rnorm = 0.0;
    const int p     = opt.normType;
    for (int l = lmin; l <= lmax; ++l) {
        double*       finePhi = l < lmax ? phi[l + 1] : nullptr;
        const double* crsePhi = l > 0 ? phi[l - 1] : nullptr;
        computeAMRResidualLevel(res[l], finePhi, phi[l], crsePhi, rhs[l], l);
    }
    // (the norms need every level's residual only through the covered-cell mask, so they can follow the loop)
    if (computeNorm)
        for (int l = lmin; l <= lmax; ++l) {
            const double localNorm = ops[l]->AMRNormLevel(res[l], l == lmax ? nullptr : ops[l + 1], p);
            if (p == 0) rnorm = std::max(rnorm, localNorm);
            else rnorm = rnorm + std::pow(localNorm, p);
        }
    if (p != 0) rnorm = std::pow(rnorm, 1.0 / (double)p);
    return rnorm;
}

// AMRHybridSolver::amrVCycle_residualEq (:327-417)
void AMRSolver::amrVCycle_residualEq(std::vector<double*>& phi, std::vector<double*>& rhs, int lev)
{
    Op& op = *ops[lev];
    if (lev > lmin) {
        Op&     crseOp      = *ops[lev - 1];
        double* crsePhi     = phi[lev - 1];
        double* crseRhs     = rhs[lev - 1];
        double* scr         = scratch[lev];
        double* crseScratch = scratch[lev - 1];
        CFLink& L           = *op.cf;
        const double* reallyCrsePhi = lev >= 2 ? phi[lev - 2] : nullptr;

        // --- smooth down: a whole level solve (:660-685) ---
        hybrid[lev]->solve(phi[lev], rhs[lev], true, false, -1.0);

        // --- restrict residual ---
        crseOp.assignLocal(crseScratch, crseRhs);
        computeAMRResidualLevel(crseRhs, phi[lev], crsePhi, reallyCrsePhi, crseScratch, lev - 1);
        op.AMRResidual(scr, nullptr, nullptr, phi[lev], &crseOp, crsePhi, rhs[lev]);     // AMRResidualNF
        k::restrict_avg(op.st(), op.lay, L.cfLay, L.ref, L.cfA, scr);                   // MGRestrict onto the coarsened grids
        L.cfToCrse.exec(op.ctx, L.cfLay, L.cfA, crseOp.lay, crseRhs);                   // crseOp->assign(crseRhs, resC, copier)

        // --- coarse solve ---
        const int numCycles = std::abs(opt.numCycles);
        for (int c = 0; c < numCycles; ++c) amrVCycle_residualEq(phi, rhs, lev - 1);

        // --- prolong increment (order 0) and update residual ---
        L.crseToCf.exec(op.ctx, crseOp.lay, crsePhi, L.cfLay, L.cfA);
        k::prolong_const(op.st(), op.lay, L.cfLay, L.ref, phi[lev], L.cfA);
        op.removeKernel(phi[lev]);
        op.assignLocal(scr, rhs[lev]);                                                   // AMRUpdateResidualNF
        op.AMRResidual(rhs[lev], nullptr, nullptr, phi[lev], &crseOp, crsePhi, scr);

        // --- smooth up, in incremental form ---
        op.setToZero(scr);
        hybrid[lev]->solve(scr, rhs[lev], true, false, -1.0);
        op.incr(phi[lev], scr, 1.0);
    } else {
        hybrid[lev]->solve(phi[lev], rhs[lev], true, false, -1.0);  // bottom solve
    }
}

// AMRHybridSolver::solve (:128-322)
SolverStatus AMRSolver::solve(std::vector<double*>& vphi, const std::vector<const double*>& vrhs, bool homog, bool setPhiToZero,
                              double a_metric)
{
    (void)homog;  // the physical BCs of this ABI carry no data (see Op::applyBCs)
    status.clear();
    absResNorms.clear();
    std::vector<double> relResNorms;
    if (!vphi[lbase]) SB_FAIL("AMRHybridSolver::solve: phi on level lbase is missing");
    for (int l = lmin; l <= lmax; ++l)
        if (!vphi[l] || !vrhs[l]) SB_FAIL("AMRHybridSolver::solve: missing level data");
    if (setPhiToZero)
        for (int l = lbase; l <= lmax; ++l) ops[l]->setToZero(vphi[l]);

    double resNorm     = computeAMRResidual(vr, vphi, vrhs, true);
    double initResNorm = resNorm;
    if (a_metric > 0.0) initResNorm = a_metric;
    else if (opt.convergenceMetric > 0.0) initResNorm = opt.convergenceMetric;
    absResNorms.push_back(resNorm);
    relResNorms.push_back(1.0);
    status.initResNorm = absResNorms[0];
    lastIters          = 0;
    if (absResNorms.back() < opt.absTol) {
        status.finalResNorm = absResNorms.back();
        status.status       = SB_STATUS_CONVERGED;
        return status;
    }
    int iter;
    for (iter = 1; iter <= opt.maxIters; ++iter) {
        for (int l = lbase; l <= lmax; ++l) ops[l]->setToZero(ve[l]);
        amrVCycle_residualEq(ve, vr, lmax);
        for (int l = lmin; l <= lmax; ++l) ops[l]->incr(vphi[l], ve[l], 1.0);
        resNorm = computeAMRResidual(vr, vphi, vrhs, true);
        absResNorms.push_back(resNorm);
        relResNorms.push_back(absResNorms.back() / initResNorm);
        lastIters = iter;
        if (absResNorms.back() < opt.absTol) { status.status = SB_STATUS_CONVERGED; break; }
        if (relResNorms.back() < opt.relTol) { status.status = SB_STATUS_CONVERGED; break; }
        if (relResNorms[iter] > relResNorms[iter - 1]) {
            for (int l = lmin; l <= lmax; ++l) ops[l]->incr(vphi[l], ve[l], -1.0);
            absResNorms.pop_back();
            relResNorms.pop_back();
            --iter;
            lastIters     = iter;
            status.status = SB_STATUS_DIVERGED;
            break;
        }
        if (relResNorms[iter] > (1.0 - opt.hang) * relResNorms[iter - 1]) { status.status = SB_STATUS_HANG; break; }
    }
    if (iter == opt.maxIters) status.status = SB_STATUS_MAXITERS;  // as the reference tests it (:291-296)
    // (the reference never records a final norm after the loop; report the last one)
    status.finalResNorm = absResNorms.back();
    for (int l = lmin; l <= lmax; ++l)
        if (ops[l]->relaxMethod == SB_RELAX_VERTLINE && !hybrid[l]->mg.ops.empty()) hybrid[l]->mg.checkPivotAll();
    return status;
}

}  // namespace sb
