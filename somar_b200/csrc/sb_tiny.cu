// sb_tiny.cu -- the bottom of a V-cycle in ONE kernel launch: the bottom smooths and the whole BiCGStab solve of
// MGSolver::vCycle_residualEq's deepest depth (MGSolverI.H:657-664 -> LevelSolverI.H:252-536), run by a single CTA.
//
// Why.  The bottom grid of the semicoarsened hierarchy is tiny (8 x 8 x 2 cells on the S5 grid), so the host-driven solver
// (BiCGStabSolver::solve, sb_mg.cpp) is pure latency: ~40 launches and six host round trips (dot products, norms) per
// iteration, 1.0 ms per V-cycle on one GPU -- and on 8 GPUs that millisecond is 8 % of the step (it runs on rank 0 while the
// others wait).  Here one CTA keeps every vector in (L1 / L2 resident) global memory, synchronises with __syncthreads, and
// makes the solver's decisions on the device: no launches, no round trips.
//
// Same arithmetic as the host-driven path, statement for statement in the solver logic:
//   * applyOp / residual: stencil7 expression of apply_op_k, ghosts as fill_ghosts_dir_k (Robin / periodic / homogeneous CF);
//   * relaxation: vertical line relaxation with dgtsv's no-interchange elimination order (the arithmetic of vertline_k,
//     PoissonOpF.ChF:851-1019) or point red-black Gauss-Seidel (gsrb_k, PoissonOpF.ChF:420-474), physical ghosts refreshed
//     before the first colour only (PoissonOp.cpp:1957-1965);
//   * norms per reference box and combined in box order (FArrayBox.cpp:138-141, LDFABOps.cpp:134-165), dot products summed
//     over boxes in box order; only the order of the additions inside a box differs (as in reduce1_k).
#include "sb_core.h"

namespace sb {
namespace k {

void note_launch();

namespace {
constexpr int TB      = 512;
constexpr int MAXNZ   = 32;
constexpr int MAXBOX  = 256;

struct Dev {
    const TinyBottomArgs& A;
    double*               boxval;  // shared [MAXBOX]
    double*               red;     // shared [32]
    double*               bc;      // shared [1] broadcast
};

template <class F>
__device__ __forceinline__ void for_valid(const Lay& L, F f)
{
    const int n = L.nx * L.ny * L.nz;
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
        const int i = m % L.nx, j = (m / L.nx) % L.ny, k = m / (L.nx * L.ny);
        f(L.idx(i, j, k), i, j, k);
    }
}

// fill_ghosts_dir_k for every direction; physToo = false refreshes only the periodic images (Op::exchange)
__device__ void t_fill_ghosts(const TinyBottomArgs& A, double* phi, bool physToo)
{
    const Lay& L = A.L;
    for (int dir = 0; dir < 3; ++dir) {
        if (A.dim == 2 && dir == 1) continue;
        int       na, nb, nn;
        long long sa, sb, sn;
        if (dir == 0) { na = L.ny; nb = L.nz; sa = L.sy; sb = L.sz; sn = 1; nn = L.nx; }
        else if (dir == 1) { na = L.nx; nb = L.nz; sa = 1; sb = L.sz; sn = L.sy; nn = L.ny; }
        else { na = L.nx; nb = L.ny; sa = 1; sb = L.sy; sn = L.sz; nn = L.nz; }
        for (int m = threadIdx.x; m < na * nb; m += blockDim.x) {
            const int       a = m % na, b = m / na;
            const long long base = L.idx(0, 0, 0) + sa * a + sb * b;
            for (int side = 0; side < 2; ++side) {
                const SideBC&   bc = A.side[dir][side];
                const long long g  = side ? base + sn * nn : base - sn;
                const long long p0 = side ? base + sn * (nn - 1) : base;
                const long long p1 = side ? base + sn * (nn - 2) : base + sn;
                if (bc.kind >= 0 && sideIsBC(bc.kind)) {
                    if (physToo) phi[g] = sideGhost(bc, phi[p0], phi[p1]);
                } else if (bc.kind == SIDE_PERIODIC_SELF) {
                    phi[g] = side ? phi[base] : phi[base + sn * (nn - 1)];
                }
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ double t_stencil7(const Lay& L, const Coef& c, const double* phi, long long q, int i, int j, int k)
{
    double s = c.mxl[i] * phi[q - 1] + c.mxr[i] * phi[q + 1] + c.myl[j] * phi[q - L.sy] + c.myr[j] * phi[q + L.sy];
    s        = s + c.mzl[k] * phi[q - L.sz] + c.mzr[k] * phi[q + L.sz];
    return s;
}
// Op::applyOp (rhs == null) / Op::residual: applyBCs, then apply_op_k
__device__ void t_apply(const TinyBottomArgs& A, double* out, double* phi, const double* rhs)
{
    t_fill_ghosts(A, phi, true);
    const Coef& c = A.c;
    for_valid(A.L, [&](long long q, int i, int j, int k) {
        const double lap = t_stencil7(A.L, c, phi, q, i, j, k);
        const double lhs = c.beta * c.J[q] * lap + phi[q] / c.Dinv[q];
        out[q]           = rhs ? rhs[q] - lhs : lhs;
    });
    __syncthreads();
}

// one colour of vertical line relaxation, one thread per column: vertline_k's statements with the modified diagonal and
// right-hand side in thread-local arrays
__device__ void t_line_pass(const TinyBottomArgs& A, double* phi, const double* rhs, int pass)
{
    const Lay&  L  = A.L;
    const Coef& c  = A.c;
    const int   hx = (L.nx + 1) / 2, N = L.nz;
    for (int m = threadIdx.x; m < hx * L.ny; m += blockDim.x) {
        const int ic = m % hx, j = m / hx;
        const int i  = ((L.lo0 + L.lo1 + j + pass) & 1) + 2 * ic;
        if (i >= L.nx) continue;
        const double    mxl = c.mxl[i], mxr = c.mxr[i], myl = c.myl[j], myr = c.myr[j];
        const long long slab = (long long)(OX + i) + L.sy * (long long)(1 + j);
        long long       q    = L.idx(i, j, 0);
        double          wd[MAXNZ], wb[MAXNZ];
        double lphi = mxl * phi[q - 1] + mxr * phi[q + 1] + myl * phi[q - L.sy] + myr * phi[q + L.sy];
        double Jb   = c.J[q] * c.beta;
        double b    = rhs[q] - Jb * lphi;
        double d    = 1.0 / c.Dinv[q] + c.loBC[slab];
        if (N == 1) {
            d = d + c.hiBC[slab];
            if (d == 0.0) { atomicOr(A.pivotFlag, 2); continue; }  // dgtsv INFO = N: the reference leaves B unsolved
            phi[q] = b / d;
            continue;
        }
        double du  = c.beta * c.J[q] * c.mzr[0];
        int    bad = 0;
        for (int k = 0; k < N - 1; ++k) {
            const long long qn = q + L.sz;
            const double    Jn = c.J[qn];
            lphi               = mxl * phi[qn - 1] + mxr * phi[qn + 1] + myl * phi[qn - L.sy] + myr * phi[qn + L.sy];
            double bn          = rhs[qn] - Jn * c.beta * lphi;
            double dn          = 1.0 / c.Dinv[qn];
            if (k + 1 == N - 1) dn = dn + c.hiBC[slab];
            const double dl = c.beta * Jn * c.mzl[k + 1];
            if (!(fabs(d) >= fabs(dl)) || d == 0.0) bad = 1;
            const double fact = dl / d;
            wd[k]             = d;
            wb[k]             = b;
            dn                = dn - fact * du;
            bn                = bn - fact * b;
            d                 = dn;
            b                 = bn;
            du                = c.beta * Jn * c.mzr[k + 1];
            q                 = qn;
        }
        if (d == 0.0) bad = 1;
        if (bad) atomicOr(A.pivotFlag, 1);
        double x = b / d;
        phi[q]   = x;
        for (int k = N - 2; k >= 0; --k) {
            q -= L.sz;
            const double duk = c.beta * c.J[q] * c.mzr[k];
            x                = (wb[k] - duk * x) / wd[k];
            phi[q]           = x;
        }
    }
    __syncthreads();
}
// one colour of point red-black Gauss-Seidel (gsrb_k)
__device__ void t_gsrb_pass(const TinyBottomArgs& A, double* phi, const double* rhs, int pass)
{
    const Lay&  L = A.L;
    const Coef& c = A.c;
    for_valid(L, [&](long long q, int i, int j, int k) {
        if ((L.lo0 + i + L.lo1 + j + L.lo2 + k + pass) & 1) return;
        const double s = t_stencil7(L, c, phi, q, i, j, k);
        phi[q]         = (rhs[q] - c.beta * c.J[q] * s) * c.Dinv[q];
    });
    __syncthreads();
}
// Op::relax for the two red-black relaxers (PoissonOp.cpp:1833-1870, 1927-2010)
__device__ void t_relax(const TinyBottomArgs& A, double* cor, const double* res, int iters)
{
    for (int it = 0; it < iters; ++it)
        for (int pass = 0; pass < 2; ++pass) {
            t_fill_ghosts(A, cor, pass == 0);
            if (A.relaxMethod == SB_RELAX_VERTLINE) t_line_pass(A, cor, res, pass);
            else t_gsrb_pass(A, cor, res, pass);
        }
}
// Op::preCond (PoissonOp.cpp:893-911)
__device__ void t_precond(const TinyBottomArgs& A, double* phi, const double* rhs, int iters)
{
    for_valid(A.L, [&](long long q, int, int, int) { phi[q] = rhs[q] * A.c.Dinv[q]; });
    __syncthreads();
    t_relax(A, phi, rhs, iters);
}

// Sum over each reference box of |x| (op 1), x^2 (2), x * y (3) or max |x| (0): a warp per box when there are many boxes,
// the whole CTA per box when there are few.  boxval[b] is valid for every thread on return.
__device__ void t_box_reduce(const Dev& D, int op, const double* x, const double* y)
{
    const TinyBottomArgs& A = D.A;
    const Lay&            L = A.L;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    auto term = [&](long long q) -> double {
        const double v = x[q];
        return op == 0 ? fabs(v) : op == 1 ? fabs(v) : op == 2 ? v * v : v * y[q];
    };
    auto comb = [&](double a, double b) -> double { return op == 0 ? fmax(a, b) : a + b; };
    if (A.nboxes >= nw) {
        for (int bx = w; bx < A.nboxes; bx += nw) {
            const int lo0 = A.boxLo[3 * bx], lo1 = A.boxLo[3 * bx + 1], lo2 = A.boxLo[3 * bx + 2];
            const int n0 = A.boxHi[3 * bx] - lo0 + 1, n1 = A.boxHi[3 * bx + 1] - lo1 + 1, n2 = A.boxHi[3 * bx + 2] - lo2 + 1;
            double    a = 0.0;
            for (int m = lane; m < n0 * n1 * n2; m += 32) a = comb(a, term(L.idx(lo0 + m % n0, lo1 + (m / n0) % n1, lo2 + m / (n0 * n1))));
            for (int o = 16; o > 0; o >>= 1) a = comb(a, __shfl_down_sync(0xffffffffu, a, o));
            if (lane == 0) D.boxval[bx] = a;
        }
        __syncthreads();
    } else {
        for (int bx = 0; bx < A.nboxes; ++bx) {
            const int lo0 = A.boxLo[3 * bx], lo1 = A.boxLo[3 * bx + 1], lo2 = A.boxLo[3 * bx + 2];
            const int n0 = A.boxHi[3 * bx] - lo0 + 1, n1 = A.boxHi[3 * bx + 1] - lo1 + 1, n2 = A.boxHi[3 * bx + 2] - lo2 + 1;
            double    a = 0.0;
            for (int m = threadIdx.x; m < n0 * n1 * n2; m += blockDim.x)
                a = comb(a, term(L.idx(lo0 + m % n0, lo1 + (m / n0) % n1, lo2 + m / (n0 * n1))));
            for (int o = 16; o > 0; o >>= 1) a = comb(a, __shfl_down_sync(0xffffffffu, a, o));
            if (lane == 0) D.red[w] = a;
            __syncthreads();
            if (w == 0) {
                a = lane < nw ? D.red[lane] : 0.0;
                for (int o = 16; o > 0; o >>= 1) a = comb(a, __shfl_down_sync(0xffffffffu, a, o));
                if (lane == 0) D.boxval[bx] = a;
            }
            __syncthreads();
        }
    }
}
// Op::norm (LDFABOps.cpp:134-165, FArrayBox.cpp:117-160): box norms combined in box order
__device__ double t_norm(const Dev& D, const double* x, int p)
{
    t_box_reduce(D, p, x, nullptr);
    const TinyBottomArgs& A = D.A;
    if (threadIdx.x == 0) {
        double ret = 0.0;
        for (int b = 0; b < A.nboxes; ++b) {
            const double numPts = (double)(A.boxHi[3 * b] - A.boxLo[3 * b] + 1) * (double)(A.boxHi[3 * b + 1] - A.boxLo[3 * b + 1] + 1) *
                                  (double)(A.boxHi[3 * b + 2] - A.boxLo[3 * b + 2] + 1);
            if (p == 0) ret = fmax(ret, D.boxval[b]);
            else if (p == 1) ret += D.boxval[b] / numPts;
            else { const double bv = sqrt(D.boxval[b] / numPts); ret += bv * bv; }  // pow(boxVal, 2), correctly rounded
        }
        if (p == 2) ret = sqrt(ret);  // pow(ret, 1 / 2)
        D.bc[0] = ret;
    }
    __syncthreads();
    const double r = D.bc[0];
    __syncthreads();
    return r;
}
__device__ double t_dot(const Dev& D, const double* a, const double* b)
{
    t_box_reduce(D, 3, a, b);
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int i = 0; i < D.A.nboxes; ++i) v += D.boxval[i];
        D.bc[0] = v;
    }
    __syncthreads();
    const double r = D.bc[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ void t_incr(const Lay& L, double* y, const double* x, double s)
{
    for_valid(L, [&](long long q, int, int, int) { y[q] = y[q] + s * x[q]; });
    __syncthreads();
}
__device__ __forceinline__ void t_copy(const Lay& L, double* y, const double* x)
{
    for_valid(L, [&](long long q, int, int, int) { y[q] = x[q]; });
    __syncthreads();
}
__device__ __forceinline__ void t_zero(const Lay& L, double* y)
{
    for (long long m = threadIdx.x; m < L.n; m += blockDim.x) y[m] = 0.0;  // setToZero clears the whole array (k::fill)
    __syncthreads();
}
__device__ __forceinline__ void t_scale(const Lay& L, double* y, double s)
{
    for_valid(L, [&](long long q, int, int, int) { y[q] = y[q] * s; });
    __syncthreads();
}

// MGSolver::vCycle_residualEq at the deepest depth: relax(numSmoothBottom), then BiCGStabSolver::solve(cor, res, homog = true,
// setPhiToZero = false) -- the statements of sb_mg.cpp's BiCGStabSolver::solve, i.e. of LevelSolverI.H:252-536.
__global__ void __launch_bounds__(TB, 1) tiny_bottom_k(TinyBottomArgs A)
{
    __shared__ double s_boxval[MAXBOX];
    __shared__ double s_red[32];
    __shared__ double s_bc[2];
    const Dev  D{A, s_boxval, s_red, s_bc};
    const Lay& L = A.L;
    double* const phi = A.phi;
    const double* rhs = A.rhs;
    double *const r = A.w[0], *const r_tilde = A.w[1], *const e = A.w[2], *const p = A.w[3], *const p_tilde = A.w[4],
                  *const s_tilde = A.w[5], *const t = A.w[6], *const v = A.w[7];
    const sb_bottom_options& opt = A.opt;

    if (A.corIsPreCond) {  // preCond(cor, res, 0) of the caller, deferred to here (Op::RELAX_PRE_PRECOND)
        for_valid(L, [&](long long q, int, int, int) { phi[q] = rhs[q] * A.c.Dinv[q]; });
        __syncthreads();
    }
    t_relax(A, phi, rhs, A.numSmoothBottom);
    if (!A.useBottomSolver) return;

    int    status = SB_STATUS_UNDEFINED;
    int    recount = 0;
    t_apply(A, r, phi, rhs);
    t_copy(L, r_tilde, r);
    t_zero(L, e);
    t_zero(L, p_tilde);
    t_zero(L, s_tilde);
    int    i      = 0;
    double rho[4] = {0, 0, 0, 0};
    double norm[2];
    norm[0]              = t_norm(D, r, opt.normType);
    double initial_norm  = norm[0];
    double initial_rnorm = norm[0];
    norm[1]              = norm[0];
    const double initResNorm = initial_norm;
    double finalResNorm = -1.0;
    double alpha[2] = {0, 0}, beta[2] = {0, 0}, omega[2] = {0, 0};
    bool   init     = true;
    int    restarts = 0;
    bool   done     = false;
    if (opt.convergenceMetric > 0.0) initial_norm = opt.convergenceMetric;
    const double smallReal = 1.0e4 * 2.220446049250313e-16;

    while ((i < opt.maxIters && norm[0] > opt.absTol * norm[1]) && (norm[1] > 0)) {
        i++;
        norm[1] = norm[0]; alpha[1] = alpha[0]; beta[1] = beta[0]; omega[1] = omega[0];
        rho[3] = rho[2]; rho[2] = rho[1];
        rho[1] = t_dot(D, r_tilde, r);
        if (fabs(rho[1]) < smallReal) {
            t_incr(L, phi, e, 1.0);
            finalResNorm = initial_norm;
            status       = SB_STATUS_SINGULAR;
            done         = true;
            break;
        }
        if (init) {
            t_copy(L, p, r);
            init = false;
        } else {
            beta[1] = (rho[1] / rho[2]) * (alpha[1] / omega[1]);
            t_scale(L, p, beta[1]);
            t_incr(L, p, v, -beta[1] * omega[1]);
            t_incr(L, p, r, 1.0);
        }
        t_precond(A, p_tilde, p, opt.numSmoothPrecond);
        t_apply(A, v, p_tilde, nullptr);
        const double m = t_dot(D, r_tilde, v);
        alpha[0]       = rho[1] / m;
        if (fabs(m) > opt.small * fabs(rho[1])) {
            t_incr(L, r, v, -alpha[0]);
            norm[0] = t_norm(D, r, opt.normType);
            t_incr(L, e, p_tilde, alpha[0]);
        } else {
            t_zero(L, r);
            norm[0] = 0.0;
        }
        if (norm[0] > opt.absTol * initial_norm && norm[0] > opt.relTol * initial_rnorm) {
            t_precond(A, s_tilde, r, opt.numSmoothPrecond);
            t_apply(A, t, s_tilde, nullptr);
            const double tr = t_dot(D, t, r);
            const double tt = t_dot(D, t, t);
            omega[0]        = tr / tt;
            t_incr(L, e, s_tilde, omega[0]);
            t_incr(L, r, t, -omega[0]);
            norm[0] = t_norm(D, r, opt.normType);
        }
        if (norm[0] <= opt.absTol * initial_norm || norm[0] <= opt.relTol * initial_rnorm) {
            finalResNorm = norm[0];
            status       = SB_STATUS_CONVERGED;
            break;
        }
        if (omega[0] == 0.0 || norm[0] > (1.0 - opt.hang) * norm[1]) {
            if (recount == 0) {
                recount = 1;
            } else {
                recount = 0;
                t_incr(L, phi, e, 1.0);
                if (restarts == opt.maxRestarts) {
                    finalResNorm = norm[0];
                    status       = SB_STATUS_MAXITERS;
                    done         = true;
                    break;
                }
                t_apply(A, r, phi, rhs);
                norm[0] = t_norm(D, r, opt.normType);
                rho[1] = 0.0; rho[2] = 0.0; rho[3] = 0.0;
                alpha[0] = 0; beta[0] = 0; omega[0] = 0;
                t_copy(L, r_tilde, r);
                t_zero(L, e);
                restarts++;
                init = true;
            }
        }
    }
    if (!done) {
        t_incr(L, phi, e, 1.0);
        finalResNorm = norm[0];
    }
    if (threadIdx.x == 0 && A.out) {
        A.out[0] = (double)status; A.out[1] = initResNorm; A.out[2] = finalResNorm; A.out[3] = (double)i; A.out[4] = (double)restarts;
    }
}
}  // namespace

bool tiny_bottom_fits(const Lay& L, int nboxes) { return L.nz <= MAXNZ && nboxes <= MAXBOX && (long long)L.nx * L.ny * L.nz <= 4096; }
void tiny_bottom(cudaStream_t st, const TinyBottomArgs& args)
{
    tiny_bottom_k<<<1, TB, 0, st>>>(args);
    note_launch();
}

}  // namespace k
}  // namespace sb
