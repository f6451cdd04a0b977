// B200PoissonOp.cpp -- see B200PoissonOp.H.
#include "B200PoissonOp.H"

#include <cstdlib>

#include "AnisotropicRefinementTools.H"
#include "CartesianMap.H"
#include "SPMD.H"

namespace somar_b200 {

void check(int a_rc, const char* a_where)
{
    if (a_rc != 0) {
        std::string msg = std::string(a_where) + ": " + sb_last_error();
        MayDay::Error(msg.c_str());
    }
}

sb_context* deviceContext()
{
    static sb_context* s_ctx = nullptr;
    if (!s_ctx) {
        const char* dev = std::getenv("SOMAR_B200_DEVICE");
        check(sb_context_create(&s_ctx, dev ? std::atoi(dev) : 0, procID(), numProc()), "sb_context_create");
    }
    return s_ctx;
}

B200PoissonOp::Handle::~Handle()
{
    for (sb_field*& p : f) if (p) sb_field_destroy(p);
    for (auto& set : fl) for (sb_field*& p : set) if (p) sb_field_destroy(p);
    if (op) sb_op_destroy(op);
}

namespace {
// GeoSourceInterface::interp behind the C ABI's map callback.  mu is the library's slot (2-D builds: x, z in slots 0, 2).
void mapCallback(double* a_x, const double* a_xi, int a_n, int a_mu, void* a_user)
{
    const GeoSourceInterface* geo = static_cast<const GeoSourceInterface*>(a_user);
    Vector<Real> xi(a_n), x(a_n);
    for (int i = 0; i < a_n; ++i) xi[i] = a_xi[i];
    geo->interp(x, xi, (SpaceDim == 2 && a_mu == 2) ? 1 : a_mu);
    for (int i = 0; i < a_n; ++i) a_x[i] = x[i];
}
void to3(int a_out[3], const IntVect& a_iv, const int a_fill)
{
    a_out[0] = a_iv[0];
    a_out[1] = SpaceDim == 2 ? a_fill : a_iv[1];
    a_out[2] = a_iv[SpaceDim - 1];
}
}  // namespace

B200PoissonOp::B200PoissonOp(const LevelGeometry& a_levGeo, const DisjointBoxLayout& a_fineGrids, const DisjointBoxLayout& a_crseGrids,
                             const int a_numComps, const std::shared_ptr<BCTools::BCFunction> a_bcFuncPtr, const Real a_alpha,
                             const Real a_beta, const LevelData<FluxBox>* a_JgupPtr)
: Elliptic::PoissonOp(a_levGeo, a_fineGrids, a_crseGrids, a_numComps, a_bcFuncPtr, a_alpha, a_beta, a_JgupPtr)
, m_h(new Handle)
{
    if (a_numComps != 1) MayDay::Error("B200PoissonOp: one component (the projector's pressure)");
    sb_level_desc d;
    std::memset(&d, 0, sizeof(d));
    d.dim = SpaceDim;
    const Box& dom = m_domain.domainBox();
    to3(d.domain_lo, dom.smallEnd(), 0);
    to3(d.domain_hi, dom.bigEnd(), 0);
    for (int dir = 0; dir < SpaceDim; ++dir) {
        d.periodic[slotOfDir(dir)] = m_domain.isPeriodic(dir) ? 1 : 0;
        d.dXi[slotOfDir(dir)]      = m_dXi[dir];
    }
    if (SpaceDim == 2) d.dXi[1] = 1.0;
    // grids, in layout order, with their owners
    std::vector<int> lo, hi, rk;
    for (LayoutIterator lit = m_grids.layoutIterator(); lit.ok(); ++lit) {
        int l[3], h[3];
        to3(l, m_grids[lit].smallEnd(), 0);
        to3(h, m_grids[lit].bigEnd(), 0);
        lo.insert(lo.end(), l, l + 3);
        hi.insert(hi.end(), h, h + 3);
        rk.push_back((int)m_grids.procID(lit()));
    }
    d.num_boxes = (int)rk.size();
    d.box_lo = lo.data(); d.box_hi = hi.data(); d.box_rank = rk.data();
    // coarser AMR level (PoissonOp.cpp:83-93, 118-120)
    std::vector<int> clo, chi, crk;
    if (m_crseAMRGrids.isClosed()) {
        for (LayoutIterator lit = m_crseAMRGrids.layoutIterator(); lit.ok(); ++lit) {
            int l[3], h[3];
            to3(l, m_crseAMRGrids[lit].smallEnd(), 0);
            to3(h, m_crseAMRGrids[lit].bigEnd(), 0);
            clo.insert(clo.end(), l, l + 3);
            chi.insert(chi.end(), h, h + 3);
            crk.push_back((int)m_crseAMRGrids.procID(lit()));
        }
        d.num_crse_boxes = (int)crk.size();
        d.crse_box_lo = clo.data(); d.crse_box_hi = chi.data(); d.crse_box_rank = crk.data();
        const Box& cdom = m_crseAMRGrids.physDomain().domainBox();
        to3(d.crse_domain_lo, cdom.smallEnd(), 0);
        to3(d.crse_domain_hi, cdom.bigEnd(), 0);
    }
    // geometry: the map itself (the 1-D matrix-element tables are rebuilt from it at every MG depth, PoissonOp.cpp:510-548)
    if (dynamic_cast<const CartesianMap*>(&m_geoSrc)) {
        d.map_kind = SB_MAP_CARTESIAN;
    } else {
        d.map_kind = SB_MAP_CALLBACK;
        d.map_fn   = &mapCallback;
        d.map_user = const_cast<GeoSourceInterface*>(&m_geoSrc);
    }
    // BCs: probe the functor once per side; this ABI carries Robin pairs without boundary data
    for (int dir = 0; dir < SpaceDim; ++dir)
        for (SideIterator sit; sit.ok(); ++sit) {
            const Box       b(IntVect::Zero, IntVect::Zero);
            FArrayBox       al(b, 1), be(b, 1), bc(b, 1), st(b, 1), x(b, SpaceDim);
            st.setVal(0.0); x.setVal(0.0); bc.setVal(0.0);
            DataIterator dit = m_grids.dataIterator();
            (*a_bcFuncPtr)(al, be, bc, st, x, dit.ok() ? dit() : DataIndex(), dir, sit(), 0.0, false);
            if (bc(IntVect::Zero, 0) != 0.0) MayDay::Error("B200PoissonOp: boundary conditions with data are not supported (HomogNeumBC expected)");
            d.bc_alpha[slotOfDir(dir)][sit() == Side::Lo ? 0 : 1] = al(IntVect::Zero, 0);
            d.bc_beta[slotOfDir(dir)][sit() == Side::Lo ? 0 : 1]  = be(IntVect::Zero, 0);
        }
    if (SpaceDim == 2) { d.bc_alpha[1][0] = d.bc_alpha[1][1] = 0.0; d.bc_beta[1][0] = d.bc_beta[1][1] = 1.0; }
    d.alpha = a_alpha; d.beta = a_beta; d.relax_method = m_relaxMethod;
    check(sb_op_create(deviceContext(), &d, &m_h->op), "sb_op_create");
    // metric: LevelGeometry's caches (or the caller's Jgup), box by box
    for (DataIterator dit(m_grids); dit.ok(); ++dit) {
        int box_id = 0, n = 0;
        for (LayoutIterator lit = m_grids.layoutIterator(); lit.ok(); ++lit, ++n)
            if (lit() == dit()) box_id = n;
        int l[3], h[3];
        const FArrayBox& J = m_J[dit];
        to3(l, J.box().smallEnd(), 0); to3(h, J.box().bigEnd(), 0);
        check(sb_op_set_metric(m_h->op, SB_CELL, box_id, J.dataPtr(0), l, h), "sb_op_set_metric(J)");
        for (int dir = 0; dir < SpaceDim; ++dir) {
            const FArrayBox& G = m_Jgup[dit][dir];
            to3(l, G.box().smallEnd(), 0); to3(h, G.box().bigEnd(), 0);
            check(sb_op_set_metric(m_h->op, slotOfDir(dir), box_id, G.dataPtr(0), l, h), "sb_op_set_metric(Jgup)");
        }
    }
    check(sb_op_finalize(m_h->op), "sb_op_finalize");
    int ns = 0;
    check(sb_op_has_null_space(m_h->op, &ns), "sb_op_has_null_space");
    if ((ns != 0) != m_hasNullSpace) MayDay::Error("B200PoissonOp: the device operator disagrees with PoissonOp::checkForNullSpace");
}

B200PoissonOp::B200PoissonOp(const B200PoissonOp& a_src) : Elliptic::PoissonOp(a_src), m_h(a_src.m_h) {}

B200PoissonOp::B200PoissonOp(const B200PoissonOp& a_fine, const DisjointBoxLayout& a_crseGrids, const IntVect& a_refRatio)
: Elliptic::PoissonOp(a_fine, a_crseGrids, a_refRatio), m_h(new Handle)
{
    int ref[3];
    to3(ref, a_refRatio, 1);
    check(sb_op_new_mg_operator(a_fine.m_h->op, ref, &m_h->op), "sb_op_new_mg_operator");
}

B200PoissonOp::~B200PoissonOp() {}

B200PoissonOp::MGOpType* B200PoissonOp::newMGOperator(const IntVect& a_refRatio) const
{
    if (a_refRatio == IntVect::Unit) return new B200PoissonOp(*this);
    DisjointBoxLayout crseGrids;
    ::coarsen(crseGrids, m_grids, a_refRatio);
    return new B200PoissonOp(*this, crseGrids, a_refRatio);
}

// ---------------------------------------------------------------------------------------------
sb_field* B200PoissonOp::field(const int a_slot) const
{
    if (!m_h->f[a_slot]) check(sb_field_create(m_h->op, SB_CELL, &m_h->f[a_slot]), "sb_field_create");
    return m_h->f[a_slot];
}
sb_field* B200PoissonOp::flux(const int a_set, const int a_dir) const
{
    sb_field*& p = m_h->fl[a_set][a_dir];
    if (!p) check(sb_field_create(m_h->op, a_dir, &p), "sb_field_create(face)");
    return p;
}
void B200PoissonOp::upload(sb_field* a_dst, const StateType& a_src) const
{
    // Valid cells only.  On the device the boxes of a rank are fused into one array, where the ghost cells of an
    // interior box side ARE the neighbouring box's valid cells: a box's (stale) host ghosts must not land there.
    // Every ghost the operator reads is refilled on the device by applyBCs.
    const DisjointBoxLayout& grids = a_src.getBoxes();
    for (DataIterator dit = a_src.dataIterator(); dit.ok(); ++dit) {
        FArrayBox tmp(grids[dit], 1);
        tmp.copy(a_src[dit], grids[dit], 0, grids[dit], 0, 1);
        int l[3], h[3];
        to3(l, tmp.box().smallEnd(), 0); to3(h, tmp.box().bigEnd(), 0);
        check(sb_field_upload(a_dst, tmp.dataPtr(0), l, h), "sb_field_upload");
    }
}
void B200PoissonOp::download(StateType& a_dst, sb_field* a_src) const
{
    const DisjointBoxLayout& grids = a_dst.getBoxes();
    for (DataIterator dit = a_dst.dataIterator(); dit.ok(); ++dit) {
        FArrayBox& fab = a_dst[dit];
        FArrayBox  tmp(grids[dit], 1);   // valid cells only: the caller's ghosts are the caller's
        int l[3], h[3];
        to3(l, tmp.box().smallEnd(), 0); to3(h, tmp.box().bigEnd(), 0);
        check(sb_field_download(a_src, tmp.dataPtr(0), l, h), "sb_field_download");
        fab.copy(tmp, grids[dit], 0, grids[dit], 0, 1);
    }
}
void B200PoissonOp::uploadFlux(const int a_set, const LevelData<FluxBox>& a_src) const
{
    const DisjointBoxLayout& grids = a_src.getBoxes();
    for (DataIterator dit = a_src.dataIterator(); dit.ok(); ++dit)
        for (int dir = 0; dir < SpaceDim; ++dir) {
            // the faces of the valid box only (a FluxBox may carry ghost faces)
            Box fc = surroundingNodes(grids[dit], dir);
            fc &= a_src[dit][dir].box();
            FArrayBox tmp(fc, 1);
            tmp.copy(a_src[dit][dir], fc, 0, fc, 0, 1);
            int l[3], h[3];
            to3(l, fc.smallEnd(), 0); to3(h, fc.bigEnd(), 0);
            check(sb_field_upload(flux(a_set, slotOfDir(dir)), tmp.dataPtr(0), l, h), "sb_field_upload(face)");
        }
}
void B200PoissonOp::downloadFlux(LevelData<FluxBox>& a_dst, const int a_set) const
{
    const DisjointBoxLayout& grids = a_dst.getBoxes();
    for (DataIterator dit = a_dst.dataIterator(); dit.ok(); ++dit)
        for (int dir = 0; dir < SpaceDim; ++dir) {
            Box fc = surroundingNodes(grids[dit], dir);
            fc &= a_dst[dit][dir].box();
            FArrayBox tmp(fc, 1);
            int l[3], h[3];
            to3(l, fc.smallEnd(), 0); to3(h, fc.bigEnd(), 0);
            check(sb_field_download(flux(a_set, slotOfDir(dir)), tmp.dataPtr(0), l, h), "sb_field_download(face)");
            a_dst[dit][dir].copy(tmp, fc, 0, fc, 0, 1);
        }
}

// ---------------------------------------------------------------------------------------------
void B200PoissonOp::applyOp(StateType& a_lhs, StateType& a_phi, const StateType* a_crsePhiPtr, const Real, const bool a_homogPhysBCs,
                            const bool a_homogCFIBCs) const
{
    if (a_crsePhiPtr && !a_homogCFIBCs) {   // inhomogeneous coarse-fine BCs: the AMR entry points (AMROperatorNF, ...) own that path
        Elliptic::PoissonOp::applyOp(a_lhs, a_phi, a_crsePhiPtr, 0.0, a_homogPhysBCs, a_homogCFIBCs);
        return;
    }
    upload(field(S_PHI), a_phi);
    check(sb_op_apply_op(handle(), field(S_LHS), field(S_PHI), a_homogPhysBCs ? 1 : 0), "sb_op_apply_op");
    download(a_lhs, field(S_LHS));
}
void B200PoissonOp::preCond(StateType& a_phi, const StateType& a_rhs, const Real, const int a_relaxIters) const
{
    upload(field(S_RHS), a_rhs);
    check(sb_op_precond(handle(), field(S_PHI), field(S_RHS), a_relaxIters), "sb_op_precond");
    download(a_phi, field(S_PHI));
}
void B200PoissonOp::relax(StateType& a_cor, const StateType& a_res, const Real, const int a_iters) const
{
    upload(field(S_PHI), a_cor);
    upload(field(S_RHS), a_res);
    check(sb_op_relax(handle(), field(S_PHI), field(S_RHS), a_iters), "sb_op_relax");
    download(a_cor, field(S_PHI));
}
void B200PoissonOp::removeKernel(StateType& a_phi) const
{
    if (!m_hasNullSpace) return;
    upload(field(S_PHI), a_phi);
    check(sb_op_remove_kernel(handle(), field(S_PHI)), "sb_op_remove_kernel");
    download(a_phi, field(S_PHI));
}
void B200PoissonOp::MGRestrict(StateType& a_crseRes, const StateType& a_fineRes, const Real, const IntVect&, const MGOpType& a_crseOp) const
{
    const B200PoissonOp& crse = dynamic_cast<const B200PoissonOp&>(a_crseOp);
    upload(field(S_RHS), a_fineRes);
    check(sb_op_mg_restrict(handle(), crse.handle(), crse.field(S_RHS), field(S_RHS)), "sb_op_mg_restrict");
    crse.download(a_crseRes, crse.field(S_RHS));
}
void B200PoissonOp::MGProlong(StateType& a_finePhi, StateType& a_crseCor, const Real, const IntVect&, const MGOpType& a_crseOp,
                              const int a_interpOrder) const
{
    const B200PoissonOp& crse = dynamic_cast<const B200PoissonOp&>(a_crseOp);
    upload(field(S_PHI), a_finePhi);
    crse.upload(crse.field(S_PHI), a_crseCor);
    check(sb_op_mg_prolong(handle(), crse.handle(), field(S_PHI), crse.field(S_PHI), a_interpOrder), "sb_op_mg_prolong");
    download(a_finePhi, field(S_PHI));
}
Real B200PoissonOp::norm(const StateType& a_x, const int a_p, const Real a_powScale) const
{
    upload(field(S_LHS), a_x);
    double v = 0.0;
    check(sb_op_norm(handle(), field(S_LHS), a_p, a_powScale, &v), "sb_op_norm");
    return v;
}
Real B200PoissonOp::dotProduct(const StateType& a_x, const StateType& a_y) const
{
    upload(field(S_LHS), a_x);
    upload(field(S_RHS), a_y);
    double v = 0.0;
    check(sb_op_dot(handle(), field(S_LHS), field(S_RHS), &v), "sb_op_dot");
    return v;
}
void B200PoissonOp::levelGradient(LevelData<FluxBox>& a_gradPhi, StateType& a_phi, const StateType* a_crsePhiPtr, const Real a_time,
                                  const bool a_homogPhysBCs, const bool a_homogCFIBCs) const
{
    if (a_crsePhiPtr && !a_homogCFIBCs) {
        Elliptic::PoissonOp::levelGradient(a_gradPhi, a_phi, a_crsePhiPtr, a_time, a_homogPhysBCs, a_homogCFIBCs);
        return;
    }
    upload(field(S_PHI), a_phi);
    sb_field* g[3] = {flux(0, 0), SpaceDim == 3 ? flux(0, 1) : nullptr, flux(0, 2)};
    check(sb_op_level_gradient(handle(), g, field(S_PHI), a_homogPhysBCs ? 1 : 0), "sb_op_level_gradient");
    downloadFlux(a_gradPhi, 0);
}
void B200PoissonOp::compGradient(LevelData<FluxBox>& a_gradPhi, StateType& a_phi, const StateType* a_crsePhiPtr, const Real a_time,
                                 const bool a_homogPhysBCs, const bool a_homogCFIBCs) const
{
    this->levelGradient(a_gradPhi, a_phi, a_crsePhiPtr, a_time, a_homogPhysBCs, a_homogCFIBCs);
}
void B200PoissonOp::levelDivergence(StateType& a_div, const LevelData<FluxBox>& a_flux) const
{
    uploadFlux(1, a_flux);
    sb_field* v[3] = {flux(1, 0), SpaceDim == 3 ? flux(1, 1) : nullptr, flux(1, 2)};
    check(sb_op_level_divergence(handle(), field(S_LHS), v), "sb_op_level_divergence");
    download(a_div, field(S_LHS));
}

};  // namespace somar_b200
