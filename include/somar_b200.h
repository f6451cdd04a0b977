/* somar_b200.h -- C ABI of the B200-native pressure projection for SOMAR.
 *
 * This is the drop-in boundary: plain C, POD arguments, opaque handles, int return codes
 * (0 = ok; on failure sb_last_error() holds the message -- the C++ shim turns that into
 * MayDay::Error, the reference's error convention, Grade1_Basics/Debug.H:258).
 *
 * Every entry point names the reference interface it stands in for (path:line relative to
 * /root/reference/src).  T = LevelData<FArrayBox>.  Conventions taken over unchanged:
 *   - fp64 everywhere, Fortran order (x fastest), cell indices are global ints and may be
 *     negative (Grade0_Chombo/BoxTools/BaseFabImplem.H:237);
 *   - the vertical is the last direction and is never split between boxes when
 *     relaxMethod == VERTLINE (Grade3_Calculus/Elliptic/PoissonOp.cpp:550-575);
 *   - 2-D problems (CH_SPACEDIM == 2 builds of the reference) are passed with dim = 2; their
 *     directions (x, z) are stored in slots 0 and 2 of every int[3]/double[3] here, slot 1
 *     must describe a single layer (lo = hi = 0);
 *   - a "field" is one LevelData: all boxes of this rank, 1 component, 1 ghost layer.
 */
#ifndef SOMAR_B200_H
#define SOMAR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sb_context sb_context;
typedef struct sb_op      sb_op;     /* Elliptic::PoissonOp at one MG depth            */
typedef struct sb_field   sb_field;  /* LevelData<FArrayBox> or one FluxBox direction  */
typedef struct sb_solver  sb_solver; /* Elliptic::MGSolver<T> / LevelHybridSolver      */
typedef struct sb_amr_solver sb_amr_solver; /* Elliptic::AMRHybridSolver             */

/* ProjectorParameters::RelaxMethod (Grade5_SOMAR/ProjectorParameters.H) */
enum { SB_RELAX_NONE = 0, SB_RELAX_JACOBI = 2, SB_RELAX_JACOBIRB = 3, SB_RELAX_GS = 4, SB_RELAX_GSRB = 5, SB_RELAX_VERTLINE = 6 };
/* Elliptic::SolverStatus (Grade3_Calculus/Elliptic/SolverStatus.H:46-51) */
enum { SB_STATUS_UNDEFINED = -1, SB_STATUS_DIVERGED = 0, SB_STATUS_CONVERGED = 1, SB_STATUS_SINGULAR = 2, SB_STATUS_MAXITERS = 3, SB_STATUS_HANG = 4 };
/* GeoSourceInterface implementations (Grade3_Calculus/maps) */
enum { SB_MAP_CARTESIAN = 0, SB_MAP_STRETCHED = 1, SB_MAP_CALLBACK = 2 };
/* field centering */
enum { SB_CELL = -1, SB_FACE_X = 0, SB_FACE_Y = 1, SB_FACE_Z = 2 };
/* streams of a context: the compute stream every operator call runs on, and two copy streams */
enum { SB_STREAM_COMPUTE = 0, SB_STREAM_H2D = 1, SB_STREAM_D2H = 2 };
/* LevelHybridSolver::SolveMode (LevelHybridSolver.H) */
enum { SB_MODE_MG = 1, SB_MODE_LEPTIC = 2, SB_MODE_LEPTIC_MG = 3 };

/* GeoSourceInterface::interp (GeoSourceInterface.H): x[n] = map_mu(xi[n]). */
typedef void (*sb_map_fn)(double* x, const double* xi, int n, int mu, void* user);

/* Everything PoissonOp's main constructor reads from LevelGeometry, the grids, the BC functor
 * and ProblemContext (PoissonOp.cpp:33-128). */
typedef struct sb_level_desc {
    int    dim;              /* 2 or 3 (CH_SPACEDIM of the reference build)                        */
    int    domain_lo[3];     /* ProblemDomain::domainBox()                                        */
    int    domain_hi[3];
    int    periodic[3];      /* ProblemDomain::isPeriodic(d)                                      */
    double dXi[3];           /* LevelGeometry::getDXi()                                           */
    int    num_boxes;        /* DisjointBoxLayout, all ranks, in layout order                     */
    const int* box_lo;       /* [num_boxes][3]                                                    */
    const int* box_hi;       /* [num_boxes][3]                                                    */
    const int* box_rank;     /* [num_boxes] owner (DisjointBoxLayout::procID); NULL = all rank 0  */
    int    map_kind;         /* SB_MAP_*                                                          */
    double map_xmin[3];      /* StretchedMap(xmin, xmax, ampl) (maps/StretchedMap.cpp:9-38)       */
    double map_xmax[3];
    double map_ampl[3];
    sb_map_fn map_fn;        /* SB_MAP_CALLBACK                                                   */
    void*  map_user;
    double bc_alpha[3][2];   /* Robin BC alpha*phi + beta*dphi/dn = 0 per (dir, side);            */
    double bc_beta[3][2];    /* BCTools::HomogNeumBC is alpha = 0, beta = 1 (AMRNSLevelBC.cpp:48) */
    double alpha;            /* operator J(alpha + beta*Lap); the projector uses 0, 1             */
    double beta;
    int    relax_method;     /* proj.relaxMethod, read by PoissonOp.cpp:54 and MGSolverI.H:154    */
    /* AMR: a_crseGrids of PoissonOp's constructor (PoissonOp.cpp:33-43, 83-93, 118-120), the
     * DisjointBoxLayout of the next coarser AMR level.  num_crse_boxes = 0 on the base level.  A
     * level with a coarser level is a refined patch: its boxes must form one rectangle, which need
     * not cover the domain; the refinement ratio is domain size / crse_domain size.  (a_fineGrids
     * is not needed here: the finer operator is an argument of every call that uses it.) */
    int    num_crse_boxes;
    const int* crse_box_lo;  /* [num_crse_boxes][3]                                               */
    const int* crse_box_hi;
    const int* crse_box_rank;/* NULL = all rank 0                                                 */
    int    crse_domain_lo[3];
    int    crse_domain_hi[3];
} sb_level_desc;

/* BiCGStabSolver<T>::Options (LevelSolver.H) */
typedef struct sb_bottom_options {
    double absTol, relTol, small, hang, convergenceMetric;
    int    maxIters, maxRestarts, normType, verbosity, numSmoothPrecond;
} sb_bottom_options;

/* MGSolver<T>::Options (MGSolver.H:57-92) */
typedef struct sb_mg_options {
    double absTol, relTol, convergenceMetric, hang;
    int    numSmoothDown, numSmoothUp, numSmoothBottom, numSmoothPrecond;
    int    prolongOrder, prolongOrderFMG, numSmoothUpFMG;
    int    maxDepth, numCycles, maxIters, normType, verbosity;
    sb_bottom_options bottom;
} sb_mg_options;

#define SB_MAX_HISTORY 64
typedef struct sb_solver_status {
    int    status;                        /* SB_STATUS_*                                        */
    int    num_iters;                     /* outer MG iterations executed                       */
    double init_res_norm;                 /* SolverStatus::getInitResNorm                       */
    double final_res_norm;                /* SolverStatus::getFinalResNorm                      */
    int    num_norms;                     /* entries of res_norms                               */
    double res_norms[SB_MAX_HISTORY];     /* MG mode: absResNorms of MGSolverI.H:281,360 in call order; leptic modes:
                                             LevelHybridSolver's m_resNorms (LevelHybridSolver.cpp:312-400) */
    int    solve_mode;                    /* SB_MODE_* (hybrid solver only)                     */
    int    max_depth;                     /* MGSolver::Options::maxDepth after define           */
    double device_ms;                     /* CUDA-event time of the solve on this rank          */
} sb_solver_status;

/* ---- lifecycle --------------------------------------------------------------------------- */
const char* sb_last_error(void);
int sb_version(void);
/* One context per process / GPU.  rank, nranks describe the horizontal decomposition. */
int sb_context_create(sb_context** ctx, int device, int rank, int nranks);
int sb_context_destroy(sb_context* ctx);
int sb_context_sync(sb_context* ctx);
/* Stream ordering for callers that pipeline independent solves (SB_STREAM_*): work enqueued on
 * `waiter` after this call starts only when everything enqueued on `signaller` so far has
 * finished; sb_context_stream_sync blocks the host on one stream. */
int sb_context_stream_wait(sb_context* ctx, int waiter, int signaller);
int sb_context_stream_sync(sb_context* ctx, int which);
/* Number of kernels this library launched on the context's stream since creation. */
long long sb_context_launch_count(sb_context* ctx);
/* CUDA-event stopwatch on the context's stream (what bench.py times with), and optional
 * per-kernel accounting: keys are "<kernel>@<depth>" with kernel in {vertline, gsrb, residual}. */
int sb_context_timer_start(sb_context* ctx);
int sb_context_timer_stop(sb_context* ctx, double* ms);
int sb_context_profile(sb_context* ctx, int enable);
int sb_context_profile_get(sb_context* ctx, const char* key, double* total_ms, long long* count);
/* NCCL communicator for halo exchange and scalar reductions (stands in for Chombo's MPI layer,
 * BoxTools/BoxLayoutDataI.H:665-812, BaseTools/Comm.cpp:19,41).  id is ncclUniqueId (128 B). */
int sb_comm_get_unique_id(void* id128);
int sb_comm_init(sb_context* ctx, const void* id128);

/* ---- host-only planning (no CUDA device needed) ---------------------------------------------- */
/* The rectangle of boxes a rank owns and what each of its sides touches: side index 2*dir+side,
 * kind 0 physical boundary, 1 periodic wrap onto itself, 2 neighbour rank (Copier::exchangeDefine,
 * BoxTools/Copier.cpp:784, at tile granularity). */
int sb_plan_tile(const sb_level_desc* desc, int rank, int nranks, int tile_lo[3], int tile_hi[3], int side_kind[6],
                 int side_neighbor[6], int* num_local_boxes);
/* The MG refinement schedule MGSolver::define would create (MGCoarseningStrategy.cpp:62-310):
 * HorizCoarseningStrategy(doVertCoarsening) when relax_method is VERTLINE, else Semicoarsening. */
/* The order in which one colour pass of the line relaxation walks the tiles (32 columns of one colour x one grid row) of an
 * nx x ny tile when it delivers its face layers to neighbouring ranks itself (sb_line_tma.cu): the tiles on the exchanged
 * sides nb_mask (bit 2 * dir + side) first.  order[3 u] = (bx, j, touched sides) of slot u.  Host only. */
int sb_plan_line_tile_order(int nx, int ny, int nb_mask, int* order, int capacity, int* num_tiles);
int sb_plan_schedule(const sb_level_desc* desc, int max_depth, int* schedule, int capacity, int* num_sched);

/* Stencil records of the quadratic coarse-fine ghost interpolation (MappedQuadCFStencil::define / buildStencils,
 * Grade2_AnisotropicChombo/QuadCFInterp/MappedCFStencil.cpp:833-1237) for side (dir, side: 0 low, 1 high) of fine box
 * `box` of a refined patch: the coarse cells under the fine ghost layer (cells[3 n]) and, per cell, weights of the
 * tangential first / second derivative stencils (w_first / w_second: [2 tangential directions, ascending][offsets -2..2],
 * to be divided by dx and dx^2) and of the mixed derivative (w_mixed: [o1 + 1][o0 + 1], offsets -1..1, to be divided by
 * dx_t0 dx_t1).  Centred, one-sided, order-dropped and out-of-buffer cases are resolved into the weights.  Call with
 * NULL arrays to get num_cells.  Host-only groundwork for the two-level AMR solve. */
int sb_plan_cf_stencils(const int dom_lo[3], const int dom_hi[3], const int periodic[3], const int ref[3], int num_fine_boxes,
                        const int* fine_lo, const int* fine_hi, int box, int dir, int side, int capacity, int* num_cells, int* cells,
                        double* w_first, double* w_second, double* w_mixed);

/* ---- PoissonOp --------------------------------------------------------------------------- */
/* PoissonOp::PoissonOp(levGeo, fineGrids, crseGrids, 1, bcFunc, alpha, beta) PoissonOp.cpp:33.
 * The metric (J, Jgup) is built the way LevelGeometry::createMetricCache does
 * (LevelGeometry.cpp:238-277) from the map in desc, or can be overwritten with
 * sb_op_set_metric before sb_op_finalize. */
int sb_op_create(sb_context* ctx, const sb_level_desc* desc, sb_op** op);
/* Raw LevelGeometry caches for one box: J over [lo,hi] (cell box), Jgup_d over the face box. */
int sb_op_set_metric(sb_op* op, int centering, int box_id, const double* host, const int lo[3], const int hi[3]);
/* PoissonOp::setAlphaAndBeta -> cacheMatrixElements + checkForNullSpace (PoissonOp.cpp:510-718). */
int sb_op_finalize(sb_op* op);
int sb_op_destroy(sb_op* op);
int sb_op_has_null_space(sb_op* op, int* out);
/* How the relaxations of this operator exchange face ghosts with neighbouring tiles (LevelData::exchange between the
 * colours, PoissonOp.cpp:1957-1965): SB_HALO_NONE one rank; SB_HALO_NCCL grouped ncclSend / ncclRecv; SB_HALO_PEER stores
 * into the neighbour's arrays over NVLink (CUDA IPC).  Meaningful once a relaxation has run. */
enum { SB_HALO_NONE = 0, SB_HALO_NCCL = 1, SB_HALO_PEER = 2 };
int sb_op_halo_mode(sb_op* op, int* out);
/* MGOperator::newMGOperator(refRatio) (PoissonOp.cpp:970-985, coarsening ctor :334-405). */
int sb_op_new_mg_operator(sb_op* op, const int ref[3], sb_op** crse);
int sb_op_get_info(sb_op* op, int domain_lo[3], int domain_hi[3], double dXi[3], int* num_local_boxes);
/* Read back a cached coefficient over the domain box in global Fortran order (tests): which =
 * 0 J, 1 Dinv, 2/3/4 M_d (length 2*N_d: lower diagonals then upper), 5/6/7 Jgup_d (face box). */
int sb_op_get_coefficient(sb_op* op, int which, double* host, long long capacity);

/* ---- fields ------------------------------------------------------------------------------ */
int sb_field_create(sb_op* op, int centering, sb_field** f);
int sb_field_destroy(sb_field* f);
/* Copy host FAB data (box [lo,hi] in the field's centering, ghosts included, Fortran order)
 * into / out of the device field; only the part inside this rank's valid+ghost region moves. */
int sb_field_upload(sb_field* f, const double* host, const int lo[3], const int hi[3]);
int sb_field_download(sb_field* f, double* host, const int lo[3], const int hi[3]);
/* The same copies, asynchronous on the context's H2D / D2H stream (host memory must be pinned);
 * order them against the compute stream with sb_context_stream_wait. */
int sb_field_upload_async(sb_field* f, const double* pinned_host, const int lo[3], const int hi[3]);
int sb_field_download_async(sb_field* f, double* pinned_host, const int lo[3], const int hi[3]);

/* ---- LevelOperator / StateOps / MGOperator methods ------------------------------------------ */
int sb_op_apply_bcs(sb_op* op, sb_field* phi, int homog);                         /* PoissonOp.cpp:726-763 */
int sb_op_apply_op(sb_op* op, sb_field* lhs, sb_field* phi, int homog);           /* PoissonOp.H applyOp   */
int sb_op_residual(sb_op* op, sb_field* res, sb_field* phi, sb_field* rhs, int homog); /* LevelOperator.H:104 */
int sb_op_relax(sb_op* op, sb_field* cor, sb_field* res, int iters);              /* PoissonOp.cpp:917     */
int sb_op_precond(sb_op* op, sb_field* phi, sb_field* rhs, int relax_iters);      /* PoissonOp.cpp:893     */
int sb_op_remove_kernel(sb_op* op, sb_field* phi);                                /* PoissonOp.cpp:821     */
int sb_op_norm(sb_op* op, sb_field* x, int p, double pow_scale, double* out);     /* LDFABOps.cpp:134      */
int sb_op_dot(sb_op* op, sb_field* a, sb_field* b, double* out);                  /* LDFABOps.cpp:98       */
int sb_op_incr(sb_op* op, sb_field* lhs, sb_field* x, double scale);              /* LDFABOps.cpp:170      */
int sb_op_axby(sb_op* op, sb_field* lhs, sb_field* x, sb_field* y, double a, double b);
int sb_op_scale(sb_op* op, sb_field* lhs, double scale);
int sb_op_set_to_zero(sb_op* op, sb_field* lhs);
int sb_op_assign_local(sb_op* op, sb_field* dst, sb_field* src);
int sb_op_mg_restrict(sb_op* fine, sb_op* crse, sb_field* crse_res, sb_field* fine_res);          /* PoissonOp.cpp:993  */
int sb_op_mg_prolong(sb_op* fine, sb_op* crse, sb_field* fine_phi, sb_field* crse_cor, int order); /* PoissonOp.cpp:1032 */
/* PoissonOp::levelDivergence (PoissonOp.cpp:1568) and levelGradient (:1486).  vel/grad are the
 * D directions of a LevelData<FluxBox>. */
int sb_op_level_divergence(sb_op* op, sb_field* div, sb_field* const vel[3]);
int sb_op_level_gradient(sb_op* op, sb_field* const grad[3], sb_field* phi, int homog);
/* AMRNSLevel::sendToAdvectingVelocity / sendToCartesianVelocity (AMRNSLevelFill.cpp:194-280), in place on
 * device-resident face fields: vel[d] *= (or /=) dx^mu/dXi^mu for the other directions mu = (d+1)%D, (d+2)%D
 * in that order.  ghost = ghost width of the caller's LevelData<FluxBox> (the reference builds its 1-D
 * tables from the grown FAB's small end, which shows in the last bits for stretched maps).  With these the
 * velocity can stay on the device across toAdvecting -> levelDivergence -> solve -> levelGradient ->
 * flux_incr -> toCartesian. */
int sb_op_send_to_advecting_velocity(sb_op* op, sb_field* const vel[3], int ghost);
int sb_op_send_to_cartesian_velocity(sb_op* op, sb_field* const vel[3], int ghost);
/* vel[d] -= scale * grad[d]  (AMRNSLevelProject.cpp:331-336 with scale 1; :155-160 with projDt). */
int sb_op_flux_incr(sb_op* op, sb_field* const vel[3], sb_field* const grad[3], double scale);

/* ---- AMRMGOperator (Elliptic/AMRMGOperator.H:43-218) and PoissonOp's AMR extras (PoissonOp.H:350-405) ---------
 * An operator created with crse_box_* in its sb_level_desc is a refined level.  Coarse data (phi_crse, crse_phi) is a
 * field of the next coarser level's operator; finer_op / phi_fine belong to the next finer level.  homog_phys is
 * accepted for interface fidelity (the Robin BCs of this ABI carry no boundary data). */
/* PoissonOp::applyBCs(phi, crsePhiPtr, time, homogPhys, homogCFI) (PoissonOp.cpp:726-763): coarse-fine ghosts by the
 * homogeneous formula (CFInterp.cpp:429-515) or, homog_cfi = 0, interpolated from crse_phi (MappedQuadCFInterp). */
int sb_op_apply_bcs_amr(sb_op* op, sb_field* phi, sb_field* crse_phi, int homog_phys, int homog_cfi);
/* AMROperator / AMROperatorNF / AMROperatorNC (PoissonOp.cpp:1156-1207) */
int sb_op_amr_operator(sb_op* op, sb_field* lhs, sb_field* phi_fine, sb_field* phi, sb_field* phi_crse, int homog_phys, sb_op* finer_op);
int sb_op_amr_operator_nf(sb_op* op, sb_field* lhs, sb_field* phi, sb_field* phi_crse, int homog_phys);
int sb_op_amr_operator_nc(sb_op* op, sb_field* lhs, sb_field* phi_fine, sb_field* phi, int homog_phys, sb_op* finer_op);
/* AMRResidual (phi_fine, finer_op and phi_crse given), AMRResidualNF (phi_fine = finer_op = NULL), AMRResidualNC
 * (phi_crse = NULL) (AMRMGOperator.H:107-175): res = rhs - L[phi]. */
int sb_op_amr_residual(sb_op* op, sb_field* res, sb_field* phi_fine, sb_field* phi, sb_field* phi_crse, sb_field* rhs, int homog_phys,
                       sb_op* finer_op);
/* AMRNormLevel(res, fineResPtr, refRatio, p) (PoissonOp.cpp:1225-1286); finer_op = NULL: plain norm with powScale = prod(dXi). */
int sb_op_amr_norm_level(sb_op* op, sb_field* res, sb_op* finer_op, int p, double* out);
/* getFlux (PoissonOp.cpp:1295-1326) on every face of the level: beta Jg^{dd} dphi/dXi_d; phi's ghosts must be filled. */
int sb_op_get_flux(sb_op* op, sb_field* const flux[3], sb_field* phi);
/* reflux(res, finePhi, phi, fineRefRatio, finerOp) (PoissonOp.cpp:1333-1418) and reflux(div, flux, fineFlux) (:1425-1477) */
int sb_op_reflux(sb_op* op, sb_field* res, sb_field* fine_phi, sb_field* phi, sb_op* finer_op);
int sb_op_reflux_flux(sb_op* op, sb_field* div, sb_field* const flux[3], sb_field* const fine_flux[3], sb_op* finer_op);
/* compDivergence(div, flux, fineFluxPtr) (PoissonOp.cpp:1618-1631); fine_flux = finer_op = NULL: no finer level */
int sb_op_comp_divergence(sb_op* op, sb_field* div, sb_field* const flux[3], sb_field* const fine_flux[3], sb_op* finer_op);
/* levelGradient / compGradient(gradPhi, phi, crsePhiPtr, time, homogPhys, homogCFI) (PoissonOp.cpp:1486-1561) */
int sb_op_comp_gradient(sb_op* op, sb_field* const grad[3], sb_field* phi, sb_field* crse_phi, int homog_phys, int homog_cfi);
/* CFInterp::coarsen(crse, fine, false, NULL) (CFInterp.cpp:846-865), what AMRNSLevel::averageDown does to the composite
 * right-hand side before the sync projection's solve (AMRNSLevelUtil.cpp:748-781, AMRNSLevelProject.cpp:752): the cells
 * of the coarser level under the fine level become the block averages of the fine data. */
int sb_op_average_down(sb_op* fine_op, sb_field* crse, sb_field* fine);
/* Bounding box of the level's boxes (the refined patch; the domain on a base level) and this rank's tile. */
int sb_op_get_patch(sb_op* op, int patch_lo[3], int patch_hi[3], int tile_lo[3], int tile_hi[3]);

/* ---- AMRHybridSolver (Elliptic/AMRHybridSolver.H) ---------------------------------------------------------------- */
/* define(Vector<shared_ptr<const AMRMGOperator>>, lmin, lmax, opts) (AMRHybridSolver.cpp:52-104): ops[0..num_levels),
 * entries below lbase = max(lmin - 1, 0) may be NULL.  opt stands for the proj.* parameters: the solver's own
 * tolerances / iteration caps and the options of the per-level LevelHybridSolvers both come from them. */
int sb_amr_solver_create(sb_op* const* ops, int num_levels, int lmin, int lmax, const sb_mg_options* opt, sb_amr_solver** s);
int sb_amr_solver_destroy(sb_amr_solver* s);
/* solve(Vector<T*>& phi, Vector<const T*>& rhs, time, homogBCs, setPhiToZero, convergenceMetric) (AMRHybridSolver.cpp:128-322).
 * phi[lbase..lmax], rhs[lmin..lmax]; status->res_norms = the composite residual norm before and after every iteration. */
int sb_amr_solver_solve(sb_amr_solver* s, sb_field* const* phi, sb_field* const* rhs, int homog, int set_phi_to_zero,
                        double convergence_metric, sb_solver_status* status);

/* ---- solvers ----------------------------------------------------------------------------- */
void sb_mg_default_options(sb_mg_options* opt);        /* ProjectorParameters.cpp:124-222 defaults */
void sb_mg_quick_and_dirty_options(sb_mg_options* opt);/* MGSolverI.H:56-74                        */
/* MGSolver<T>::define(topOp, opt, refSchedule) (MGSolverI.H:140-204).  schedule = num_sched
 * ref ratios ending with (1,1,1), or NULL/0 to run the reference's coarsening strategy. */
int sb_mgsolver_create(sb_op* top, const sb_mg_options* opt, const int* schedule, int num_sched, sb_solver** s);
/* LevelHybridSolver::define (LevelHybridSolver.cpp:127-190): picks the solve mode by lepticity. */
int sb_hybrid_solver_create(sb_op* top, const sb_mg_options* opt, sb_solver** s);
int sb_solver_destroy(sb_solver* s);
int sb_solver_get_schedule(sb_solver* s, int* schedule, int capacity, int* num_sched);
int sb_solver_set_options(sb_solver* s, const sb_mg_options* opt);   /* modifyOptionsExceptMaxDepth */
/* MGSolver<T>::solve / LevelHybridSolver::solve(phi, NULL, rhs, time, homog, setPhiToZero, metric). */
int sb_solver_solve(sb_solver* s, sb_field* phi, sb_field* rhs, int homog, int set_phi_to_zero,
                    double convergence_metric, sb_solver_status* status);
/* One vCycle_residualEq(cor, res, depth 0) (MGSolverI.H:617) -- the unit of the headline metric. */
int sb_solver_vcycle(sb_solver* s, sb_field* cor, sb_field* res);
/* The body of one outer iteration of MGSolver::vCycle (MGSolverI.H:342-347): op.preCond(cor, res, 0)
 * followed by vCycle_residualEq(cor, res, depth 0); same result as sb_op_precond + sb_solver_vcycle,
 * with the preconditioning pass fused into the first relaxation. */
int sb_solver_precond_vcycle(sb_solver* s, sb_field* cor, sb_field* res);

/* ---- the projection bracket on device-resident fields (any number of ranks) ------------------ */
/* AMRNSLevel::projectCorrect (AMRNSLevelProject.cpp:247-373) between the level's own BC fills: [vel -> advecting],
 * div = levelDivergence(vel), solve L[phi] = div (homogeneous BCs, phi from zero), vel -= levelGradient(phi),
 * p += phi / proj_dt (p may be NULL; skipped for proj_dt = 0), diagnostic divergence, [vel -> Cartesian].
 * vel_ghost >= 0: vel is the Cartesian face velocity of a FluxBox with that ghost width and is transformed on the way
 * in and out as the reference does (sb_op_send_to_*_velocity); vel_ghost < 0: vel is the advecting velocity already.
 * phi_out (may be NULL) receives the correction.  Nothing leaves the device but the two norms and the status. */
int sb_project_correct(sb_solver* s, sb_field* const vel[3], sb_field* p, double proj_dt, int vel_ghost, sb_field* phi_out,
                       double* init_div_norm, double* final_div_norm, sb_solver_status* status);
/* AMRNSLevel::projectPredict (AMRNSLevelProject.cpp:63-243): vel -= proj_dt * levelGradient(p, ..., false, false) with the
 * lagged pressure; if that raised |Div[vel]|, one quick-and-dirty projectCorrect (LevelHybridSolver::
 * getQuickAndDirtyOptions switched on around it, :176-182).  norms = {initial, lagged, corrected (-1 without the
 * fallback)}; status describes the fallback's solve (UNDEFINED if there was none). */
int sb_project_predict(sb_solver* s, sb_field* const vel[3], sb_field* p, double proj_dt, int vel_ghost, double norms[3], int* used_fallback,
                       sb_solver_status* status);

/* ---- the projection bracket, host buffers in and out --------------------------------------- */
/* AMRNSLevel::projectCorrect, single level (AMRNSLevelProject.cpp:247-373): vel is the advecting
 * (J-scaled) face velocity, D host arrays over the face-centred *domain* box in global Fortran
 * order; p may be NULL.  On return vel is projected, phi (cell-centred domain box) holds the
 * correction, p += phi/projDt.  Copies H2D/D2H are part of the call. */
int sb_project_host(sb_solver* s, double* const vel[3], double* phi, double* p, double proj_dt,
                    double* init_div_norm, double* final_div_norm, sb_solver_status* status);

#ifdef __cplusplus
}
#endif
#endif
