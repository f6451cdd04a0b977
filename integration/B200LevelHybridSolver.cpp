// B200LevelHybridSolver.cpp -- see B200LevelHybridSolver.H.
#include "B200LevelHybridSolver.H"

namespace somar_b200 {

// MGSolver<T>::Options + BiCGStabSolver<T>::Options -> sb_mg_options.  The hybrid solver's own tolerances are the
// proj.* values the MG options were built from (LevelHybridSolver.cpp:15-43).
void B200LevelHybridSolver::toC(sb_mg_options& a_out, const Options& a_in)
{
    sb_mg_default_options(&a_out);
    const auto& m = a_in.mgOptions;
    a_out.absTol = a_in.absTol; a_out.relTol = a_in.relTol; a_out.hang = a_in.hang; a_out.normType = a_in.normType;
    a_out.convergenceMetric = m.convergenceMetric;
    a_out.numSmoothDown = m.numSmoothDown; a_out.numSmoothUp = m.numSmoothUp; a_out.numSmoothBottom = m.numSmoothBottom;
    a_out.numSmoothPrecond = m.numSmoothPrecond; a_out.prolongOrder = m.prolongOrder; a_out.prolongOrderFMG = m.prolongOrderFMG;
    a_out.numSmoothUpFMG = m.numSmoothUpFMG; a_out.maxDepth = m.maxDepth; a_out.numCycles = m.numCycles; a_out.maxIters = m.maxIters;
    a_out.verbosity = 0;
    const auto& b = m.bottomOptions;
    a_out.bottom.absTol = b.absTol; a_out.bottom.relTol = b.relTol; a_out.bottom.small = b.small; a_out.bottom.hang = b.hang;
    a_out.bottom.convergenceMetric = b.convergenceMetric; a_out.bottom.maxIters = b.maxIters; a_out.bottom.maxRestarts = b.maxRestarts;
    a_out.bottom.normType = b.normType; a_out.bottom.verbosity = 0; a_out.bottom.numSmoothPrecond = b.numSmoothPrecond;
}

void B200LevelHybridSolver::define(std::shared_ptr<const Elliptic::MGOperator<StateType>> a_mgOpPtr, const Options& a_opts)
{
    this->clear();
    m_op = std::dynamic_pointer_cast<const B200PoissonOp>(a_mgOpPtr);
    if (!m_op) MayDay::Error("B200LevelHybridSolver::define needs a B200PoissonOp");
    m_options = a_opts;
    sb_mg_options o;
    toC(o, a_opts);
    check(sb_hybrid_solver_create(m_op->handle(), &o, &m_solver), "sb_hybrid_solver_create");
}

void B200LevelHybridSolver::modifyOptionsExceptMaxDepth(const Options& a_opt)
{
    if (!m_solver) MayDay::Error("This can only be called AFTER B200LevelHybridSolver is defined.");
    m_options = a_opt;
    sb_mg_options o;
    toC(o, a_opt);
    check(sb_solver_set_options(m_solver, &o), "sb_solver_set_options");
}

void B200LevelHybridSolver::clear()
{
    if (m_solver) sb_solver_destroy(m_solver);
    m_solver = nullptr;
    m_op.reset();
    m_solverStatus.clear();
    m_resNorms.clear();
}

Elliptic::SolverStatus B200LevelHybridSolver::solve(StateType& a_phi, const StateType* a_crsePhiPtr, const StateType& a_rhs, const Real,
                                                    const bool a_useHomogBCs, const bool a_setPhiToZero,
                                                    const Real a_convergenceMetric) const
{
    if (!m_solver) MayDay::Error("B200LevelHybridSolver::solve called before define");
    if (a_crsePhiPtr) MayDay::Error("B200LevelHybridSolver::solve: coarse-level data belongs to the AMR solver (sb_amr_solver_*)");
    sb_field* phi = m_op->field(B200PoissonOp::S_PHI);
    sb_field* rhs = m_op->field(B200PoissonOp::S_RHS);
    m_op->upload(rhs, a_rhs);
    if (!a_setPhiToZero) m_op->upload(phi, a_phi);
    sb_solver_status st;
    check(sb_solver_solve(m_solver, phi, rhs, a_useHomogBCs ? 1 : 0, a_setPhiToZero ? 1 : 0, a_convergenceMetric, &st), "sb_solver_solve");
    m_op->download(a_phi, phi);
    m_solverStatus.clear();
    m_solverStatus.setSolverStatus(st.status);
    m_solverStatus.setInitResNorm(st.init_res_norm);
    m_solverStatus.setFinalResNorm(st.final_res_norm);
    // m_resNorms of the reference: in MG mode the initial and the final norm (LevelHybridSolver.cpp:392-400), in the leptic
    // modes one entry per leptic order / V-cycle
    if (st.solve_mode == SB_MODE_MG) m_resNorms = {st.init_res_norm, st.final_res_norm};
    else m_resNorms.assign(st.res_norms, st.res_norms + st.num_norms);
    m_mode = st.solve_mode; m_maxDepth = st.max_depth; m_ms = st.device_ms;
    return m_solverStatus;
}

};  // namespace somar_b200
