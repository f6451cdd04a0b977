// sb_tiny.cu -- the tail of a V-cycle in ONE kernel launch: MGSolver::vCycle_residualEq (MGSolverI.H:617-754) over the deepest
// depths -- as many as fit a shared-memory arena together -- including the bottom smooths and the whole BiCGStab solve of the
// deepest one (LevelSolverI.H:252-536), run by a single CTA.
//
// Why.  The bottom grid of the semicoarsened hierarchy is tiny (8 x 8 x 2 cells on the S5 grid), so the host-driven solver
// (BiCGStabSolver::solve, sb_mg.cpp) is pure latency: ~40 launches and six host round trips (dot products, norms) per
// iteration, 1.0 ms per V-cycle on one GPU -- and on 8 GPUs that millisecond is 8 % of the step (it runs on rank 0 while the
// others wait).  Here one CTA keeps every array of those depths in shared memory, synchronises with __syncthreads, and makes
// the solver's decisions on the device: no launches, no round trips.
//
// Same arithmetic as the host-driven path, statement for statement in the solver logic:
//   * applyOp / residual: stencil7 expression of apply_op_k, ghosts as fill_ghosts_dir_k (Robin / periodic / homogeneous CF);
//   * relaxation: vertical line relaxation with dgtsv's no-interchange elimination order (the arithmetic of vertline_k,
//     PoissonOpF.ChF:851-1019) or point red-black Gauss-Seidel (gsrb_k, PoissonOpF.ChF:420-474), physical ghosts refreshed
//     before the first colour only (PoissonOp.cpp:1957-1965);
//   * restriction, linear prolongation and the null-space removal as restrict_k, prolong_k, reduce1_k (op 4) + sum_boxes_k;
//   * norms per reference box and combined in box order (FArrayBox.cpp:138-141, LDFABOps.cpp:134-165), dot products summed
//     over boxes in box order; only the order of the additions inside a box differs (as in reduce1_k).
#include <cstdlib>

#include "sb_core.h"

namespace sb {
namespace k {

void note_launch();

namespace {
// The helpers below are __noinline__ on purpose: inlined, the kernel was 468 KB of straight-line code that never fit the
// instruction cache, and every phase paid ~3 us of instruction fetch (measured: 3.5 us per colour pass on a 128-column grid).
constexpr int TB      = 512;
constexpr int MAXNZ   = 32;
constexpr int MAXBOX  = 256;

struct Dev {
    const TinyLevel&      A;
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

// fill_ghosts_dir_k for every direction; physToo = false refreshes only the periodic images (Op::exchange).  All ghost cells
// of the six sides form ONE index space, a cell per thread: dependent fp64 operations cost ~40 cycles each on this part, so
// a thread that walked the directions one after the other spent 4000 cycles here (measured).
__device__ __noinline__ void t_fill_ghosts(const TinyLevel& A, double* phi, bool physToo)
{
    const Lay& L  = A.L;
    const int  n0 = L.ny * L.nz, n1 = A.dim == 2 ? 0 : L.nx * L.nz, n2 = L.nx * L.ny;
    for (int m = threadIdx.x; m < 2 * (n0 + n1 + n2); m += blockDim.x) {
        const int side = m & 1;
        int       e = m >> 1, dir = 0;
        if (e >= n0) { e -= n0; dir = 1; if (e >= n1) { e -= n1; dir = 2; } }
        const SideBC& bc = A.side[dir][side];
        const bool    isBC = bc.kind >= 0 && sideIsBC(bc.kind);
        if (!(isBC && physToo) && bc.kind != SIDE_PERIODIC_SELF) continue;
        int       na, nn;
        long long sa, sb, sn;
        if (dir == 0) { na = L.ny; sa = L.sy; sb = L.sz; sn = 1; nn = L.nx; }
        else if (dir == 1) { na = L.nx; sa = 1; sb = L.sz; sn = L.sy; nn = L.ny; }
        else { na = L.nx; sa = 1; sb = L.sy; sn = L.sz; nn = L.nz; }
        const int       a = e % na, b = e / na;
        const long long base = L.idx(0, 0, 0) + sa * a + sb * b;
        const long long g  = side ? base + sn * nn : base - sn;
        const long long p0 = side ? base + sn * (nn - 1) : base;
        const long long p1 = side ? base + sn * (nn - 2) : base + sn;
        if (isBC) phi[g] = sideGhost(bc, phi[p0], phi[p1]);
        else phi[g] = side ? phi[base] : phi[base + sn * (nn - 1)];
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
__device__ __noinline__ void t_apply(const TinyLevel& A, double* out, double* phi, const double* rhs)
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
__device__ __noinline__ void t_line_pass(const TinyLevel& A, double* phi, const double* rhs, int pass)
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
        if (A.lineTab) {
            // every column of this depth shares one tridiagonal matrix: the recurrences of vertline_smem_k on the factorisation
            // tables of Op::buildLineTables (s, -f, g, -(MzR g)) -- no division on the dependency chain
            const double *ts = A.lineTab, *sf = ts + N, *sg = ts + 2 * N, *sc = ts + 3 * N;
            // z_k is parked in the column's own cells (read by nobody else during this colour), as vertline_smem_k does in
            // shared memory: thread-local arrays would live in local memory, which the arena leaves almost no L1 for
            // right-hand sides first (iterations independent of each other: unrolled, their ~40-cycle fp64 latencies overlap),
            // then the forward chain -- one dependent fma per level
#pragma unroll 4
            for (int k = 0; k < N; ++k) {
                const long long qk   = q + (long long)k * L.sz;
                const double    lphi = mxl * phi[qk - 1] + mxr * phi[qk + 1] + myl * phi[qk - L.sy] + myr * phi[qk + L.sy];
                phi[qk]              = rhs[qk] * ts[k] - lphi;
            }
            double y = 0.0;
            for (int k = 0; k < N; ++k) {
                const double b = phi[q];
                y              = k == 0 ? b : fma(sf[k - 1], y, b);
                phi[q]         = y * sg[k];
                q += L.sz;
            }
            q -= L.sz;
            double x = phi[q];
            for (int k = N - 2; k >= 0; --k) {
                q -= L.sz;
                x      = fma(sc[k], x, phi[q]);
                phi[q] = x;
            }
            continue;
        }
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
__device__ __noinline__ void t_gsrb_pass(const TinyLevel& A, double* phi, const double* rhs, int pass)
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
__device__ __noinline__ void t_relax(const TinyLevel& A, double* cor, const double* res, int iters)
{
    for (int it = 0; it < iters; ++it)
        for (int pass = 0; pass < 2; ++pass) {
            t_fill_ghosts(A, cor, pass == 0);
            if (A.relaxMethod == SB_RELAX_VERTLINE) t_line_pass(A, cor, res, pass);
            else t_gsrb_pass(A, cor, res, pass);
        }
}
// Op::preCond (PoissonOp.cpp:893-911)
__device__ __noinline__ void t_precond(const TinyLevel& A, double* phi, const double* rhs, int iters)
{
    for_valid(A.L, [&](long long q, int, int, int) { phi[q] = rhs[q] * A.c.Dinv[q]; });
    __syncthreads();
    t_relax(A, phi, rhs, iters);
}

// Sum over each reference box of |x| (op 1), x^2 (2), x * y (3) or max |x| (0): a warp per box when there are many boxes,
// the whole CTA per box when there are few.  boxval[b] is valid for every thread on return.
// fin: what Op::norm does to a box's sum before the boxes are combined -- 1: / numPts; 2: pow(sqrt(. / numPts), 2) -- done here,
// by the lane that owns the box, so that the division and the square root of 64 boxes do not queue up on one thread
__device__ __forceinline__ double t_box_final(double a, int fin, int n0, int n1, int n2)
{
    if (fin == 0) return a;
    const double numPts = (double)n0 * (double)n1 * (double)n2;
    if (fin == 1) return a / numPts;
    const double bv = sqrt(a / numPts);
    return bv * bv;
}
__device__ __noinline__ void t_box_reduce(const Dev& D, int op, const double* x, const double* y, int fin)
{
    const TinyLevel& A = D.A;
    const Lay&       L = A.L;
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
            if (lane == 0) D.boxval[bx] = t_box_final(a, fin, n0, n1, n2);
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
                if (lane == 0) D.boxval[bx] = t_box_final(a, fin, n0, n1, n2);
            }
            __syncthreads();
        }
    }
}
// Op::norm (LDFABOps.cpp:134-165, FArrayBox.cpp:117-160): box norms combined in box order
__device__ __noinline__ double t_norm(const Dev& D, const double* x, int p)
{
    t_box_reduce(D, p, x, nullptr, p);
    const TinyLevel& A = D.A;
    if (threadIdx.x == 0) {
        double ret = 0.0;
        for (int b = 0; b < A.nboxes; ++b) {
            if (p == 0) ret = fmax(ret, D.boxval[b]);
            else ret += D.boxval[b];   // ret += pow(boxVal, p): boxval holds pow(boxVal, p), correctly rounded for p = 1, 2
        }
        if (p == 2) ret = sqrt(ret);  // pow(ret, 1 / 2)
        D.bc[0] = ret;
    }
    __syncthreads();
    const double r = D.bc[0];
    __syncthreads();
    return r;
}
__device__ __noinline__ double t_dot(const Dev& D, const double* a, const double* b)
{
    t_box_reduce(D, 3, a, b, 0);
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
__device__ __noinline__ void t_incr(const Lay& L, double* y, const double* x, double s)
{
    for_valid(L, [&](long long q, int, int, int) { y[q] = y[q] + s * x[q]; });
    __syncthreads();
}
__device__ __noinline__ void t_copy(const Lay& L, double* y, const double* x)
{
    for_valid(L, [&](long long q, int, int, int) { y[q] = x[q]; });
    __syncthreads();
}
__device__ __noinline__ void t_zero(const Lay& L, double* y)
{
    for (long long m = threadIdx.x; m < L.n; m += blockDim.x) y[m] = 0.0;  // setToZero clears the whole array (k::fill)
    __syncthreads();
}
__device__ __noinline__ void t_scale(const Lay& L, double* y, double s)
{
    for_valid(L, [&](long long q, int, int, int) { y[q] = y[q] * s; });
    __syncthreads();
}

// restrict_k: block average in Fortran loop order (CFInterpF.ChF:1085-1120)
__device__ __noinline__ void t_restrict(const TinyLevel& F, const TinyLevel& C, double* crse, const double* fine)
{
    const int    r0 = F.ref[0], r1 = F.ref[1], r2 = F.ref[2];
    const double refScale = 1.0 / (double)(r0 * r1 * r2);
    for_valid(C.L, [&](long long qc, int i, int j, int k) {
        double s = 0.0;
        for (int c = 0; c < r2; ++c)
            for (int b = 0; b < r1; ++b)
                for (int a = 0; a < r0; ++a) s = s + fine[F.L.idx(i * r0 + a, j * r1 + b, k * r2 + c)];
        crse[qc] = s * refScale;
    });
    __syncthreads();
}
// Op::MGProlong, order 0 or 1 (PoissonOp.cpp:1032-1090): crse.applyBCs, then prolong_k's expression per fine cell
__device__ __noinline__ void t_prolong(const TinyLevel& F, const TinyLevel& C, double* fine, double* crse, int order)
{
    if (order >= 1) t_fill_ghosts(C, crse, true);
    const int r0 = F.ref[0], r1 = F.ref[1], r2 = F.ref[2];
    const Lay& Lc = C.L;
    for_valid(F.L, [&](long long qf, int i, int j, int k) {
        const int       ic = i / r0, jc = j / r1, kc = k / r2;
        const long long qc = Lc.idx(ic, jc, kc);
        double          f  = fine[qf];
        f                  = f + crse[qc];
        if (order == 1) {
            const long long s0 = r0 == 1 ? 0 : 1, s1 = r1 == 1 ? 0 : Lc.sy, s2 = r2 == 1 ? 0 : Lc.sz;
            const double    sc0 = r0 == 1 ? 0.0 : 1.0 / r0, sc1 = r1 == 1 ? 0.0 : 1.0 / r1, sc2 = r2 == 1 ? 0.0 : 1.0 / r2;
            const double    m0 = 0.5 * (crse[qc + s0] - crse[qc - s0]);
            const double    m1 = 0.5 * (crse[qc + s1] - crse[qc - s1]);
            const double    m2 = 0.5 * (crse[qc + s2] - crse[qc - s2]);
            const double    dxf0 = -0.5 + (((i - ic * r0) + 0.5) * sc0);
            const double    dxf1 = -0.5 + (((j - jc * r1) + 0.5) * sc1);
            const double    dxf2 = -0.5 + (((k - kc * r2) + 0.5) * sc2);
            f                    = f + dxf0 * m0 + dxf1 * m1 + dxf2 * m2;
        }
        fine[qf] = f;
    });
    __syncthreads();
}
// Op::removeKernel (PoissonOp.cpp:821-846, Integral.cpp:249-266, IntegralF.ChF:37-58): phi -= sum(J dv phi) / sum(J dv), the two
// sums per box, then over the boxes in box order
__device__ __noinline__ void t_remove_kernel(const Dev& D, double* phi)
{
    const TinyLevel& A = D.A;
    if (!A.hasNullSpace) return;
    const Lay& L = A.L;
    const int  lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // (sum, vol) per box: boxval[b], boxval[MAXBOX + b] -- a warp per box
    for (int bx = w; bx < A.nboxes; bx += nw) {
        const int lo0 = A.boxLo[3 * bx], lo1 = A.boxLo[3 * bx + 1], lo2 = A.boxLo[3 * bx + 2];
        const int n0 = A.boxHi[3 * bx] - lo0 + 1, n1 = A.boxHi[3 * bx + 1] - lo1 + 1, n2 = A.boxHi[3 * bx + 2] - lo2 + 1;
        double    a = 0.0, b = 0.0;
        for (int m = lane; m < n0 * n1 * n2; m += 32) {
            const long long q = L.idx(lo0 + m % n0, lo1 + (m / n0) % n1, lo2 + m / (n0 * n1));
            const double    s = A.c.J[q] * A.dv;
            a = a + s * phi[q];
            b = b + s;
        }
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
        if (lane == 0) { D.boxval[bx] = a; D.boxval[MAXBOX + bx] = b; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sum = 0.0, vol = 0.0;
        for (int b = 0; b < A.nboxes; ++b) { sum += D.boxval[b]; vol += D.boxval[MAXBOX + b]; }   // sum_boxes_k
        D.bc[0] = sum / vol;
    }
    __syncthreads();
    const double avg = D.bc[0];
    for_valid(L, [&](long long q, int, int, int) { phi[q] = phi[q] - avg; });
    __syncthreads();
}

// BiCGStabSolver::solve(phi, rhs, homog = true, setPhiToZero = false) -- the statements of sb_mg.cpp's BiCGStabSolver::solve,
// i.e. of LevelSolverI.H:252-536
__device__ __noinline__ void t_bicgstab(const Dev& D, const TinyTailArgs& T, double* const* W, double* phi, const double* rhs)
{
    const TinyLevel& A = D.A;
    const Lay&       L = A.L;
    double *const r = W[0], *const r_tilde = W[1], *const e = W[2], *const p = W[3], *const p_tilde = W[4],
                  *const s_tilde = W[5], *const t = W[6], *const v = W[7];
    const sb_bottom_options& opt = T.opt;
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
    if (threadIdx.x == 0 && T.out) {
        T.out[0] = (double)status; T.out[1] = initResNorm; T.out[2] = finalResNorm; T.out[3] = (double)i; T.out[4] = (double)restarts;
    }
}

// Staging.  A single CTA is latency-bound on every global-memory round trip (about a microsecond each on this part), and a
// V-cycle tail makes thousands of dependent ones.  So the fields, coefficient arrays and tables of the tail's levels are copied
// into a shared-memory arena first -- with the SAME layout (strides, ghosts, padding), so every index expression above works
// on the staged copy unchanged -- and only the correction of the tail's first level goes back at the end.
struct Stager {
    double* arena;
    long long top = 0;
    __device__ double* take(const double* g, long long n, bool copy)
    {
        double* p = arena + top;
        top += (n + 1) & ~1LL;
        for (long long m = threadIdx.x; m < n; m += blockDim.x) p[m] = (copy && g) ? g[m] : 0.0;
        return p;
    }
    __device__ int* takeInts(const int* g, int n)
    {
        int* p = reinterpret_cast<int*>(arena + top);
        top += ((long long)n + 1) / 2 + 1;
        for (int m = threadIdx.x; m < n; m += blockDim.x) p[m] = g[m];
        return p;
    }
};

// MGSolver::vCycle_residualEq (MGSolverI.H:617-754) from tail level 0 down to the bottom and back, numCycles = 1
__global__ void __launch_bounds__(TB, 1) tiny_tail_k(const __grid_constant__ TinyTailArgs T)
{
    extern __shared__ __align__(16) double arena[];
    __shared__ TinyLevel s_lev[TINY_MAXLEV];
    __shared__ double*   s_w[8];
    __shared__ double s_boxval[2 * MAXBOX];
    __shared__ double s_red[32];
    __shared__ double s_bc[2];
    const int nb = T.nlev - 1;
    unsigned long long tm[6];
    auto stamp = [&](int i) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm[i])); };
    stamp(0);

    // ---- stage ----
    {
        Stager S{arena};
        for (int l = 0; l < T.nlev; ++l) {
            TinyLevel V = T.lev[l];
            const Lay& L = V.L;
            const long long n = L.n;
            const int ntab = 2 * (L.nx + L.ny + L.nz);
            // level 0 brings its right-hand side (and, unless it is still to be preconditioned, its correction) along
            double* cor = S.take(T.lev[l].cor, n, l == 0 && !T.corIsPreCond);
            double* res = S.take(T.lev[l].res, n, l == 0);
            double* tmp = l < nb ? S.take(nullptr, n, false) : nullptr;
            double* J   = S.take(T.lev[l].c.J, n, true);
            double* Di  = S.take(T.lev[l].c.Dinv, n, true);
            double* mt  = S.take(T.lev[l].c.mxl, ntab, true);   // mxl | mxr | myl | myr | mzl | mzr are contiguous (Op::mtab)
            double* lt  = T.lev[l].lineTab ? S.take(T.lev[l].lineTab, 4 * L.nz, true) : nullptr;
            double* lo  = S.take(T.lev[l].c.loBC, L.sz, true);
            double* hi  = S.take(T.lev[l].c.hiBC, L.sz, true);
            int*    blo = S.takeInts(T.lev[l].boxLo, 3 * V.nboxes);
            int*    bhi = S.takeInts(T.lev[l].boxHi, 3 * V.nboxes);
            if (threadIdx.x == 0) {
                V.cor = cor; V.res = res; V.tmp = tmp;
                V.c.J = J; V.c.Dinv = Di;
                V.c.mxl = mt; V.c.mxr = mt + L.nx; V.c.myl = mt + 2 * L.nx; V.c.myr = V.c.myl + L.ny;
                V.c.mzl = mt + 2 * (L.nx + L.ny); V.c.mzr = V.c.mzl + L.nz;
                V.c.loBC = lo; V.c.hiBC = hi; V.c.tabJ = nullptr; V.c.tabD = nullptr;
                V.boxLo = blo; V.boxHi = bhi;
                V.lineTab = lt;
                s_lev[l] = V;
            }
        }
        for (int i = 0; i < 8; ++i) {
            double* p = S.take(nullptr, T.lev[nb].L.n, false);
            if (threadIdx.x == 0) s_w[i] = p;
        }
        __syncthreads();
    }
    stamp(1);

    for (int d = 0; d < nb; ++d) {
        const TinyLevel& V = s_lev[d];
        const TinyLevel& C = s_lev[d + 1];
        if (d > 0 || T.corIsPreCond) {  // crseOp.preCond(crseCor, crseRes, 0), deferred to the visit (Op::RELAX_PRE_PRECOND)
            for_valid(V.L, [&](long long q, int, int, int) { V.cor[q] = V.res[q] * V.c.Dinv[q]; });
            __syncthreads();
        }
        t_relax(V, V.cor, V.res, T.numSmoothDown);
        t_apply(V, V.tmp, V.cor, V.res);
        t_restrict(V, C, C.res, V.tmp);
    }
    {
        const TinyLevel& B = s_lev[nb];
        const Dev        D{B, s_boxval, s_red, s_bc};
        if (nb > 0 || T.corIsPreCond) {
            for_valid(B.L, [&](long long q, int, int, int) { B.cor[q] = B.res[q] * B.c.Dinv[q]; });
            __syncthreads();
        }
        stamp(2);
        t_relax(B, B.cor, B.res, T.numSmoothBottom);
        stamp(3);
        t_bicgstab(D, T, s_w, B.cor, B.res);
        stamp(4);
    }
    for (int d = nb - 1; d >= 0; --d) {
        const TinyLevel& V = s_lev[d];
        const TinyLevel& C = s_lev[d + 1];
        const Dev        D{V, s_boxval, s_red, s_bc};
        t_prolong(V, C, V.cor, C.cor, T.prolongOrder);
        t_remove_kernel(D, V.cor);
        t_relax(V, V.cor, V.res, T.numSmoothUp);
    }
    // ---- the result: the correction of the tail's first level (valid cells) ----
    {
        const TinyLevel& V = s_lev[0];
        double* const    g = T.lev[0].cor;
        for_valid(V.L, [&](long long q, int, int, int) { g[q] = V.cor[q]; });
    }
    stamp(5);
    if (threadIdx.x == 0 && T.out)   // out[0..4]: the bottom solve's record; out[5..9]: ns spent staging, going down, in the bottom smooths, in BiCGStab, going up
        for (int i = 0; i < 5; ++i) T.out[5 + i] = (double)(tm[i + 1] - tm[i]);
}
}  // namespace

// shared memory (bytes) the staged copy of a level takes; bottom: the BiCGStab work vectors as well
size_t tiny_level_bytes(const Lay& L, int nboxes, bool nonBottom, bool bottom)
{
    auto r2 = [](long long n) { return (n + 1) & ~1LL; };
    long long d = (4 + (nonBottom ? 1 : 0)) * r2(L.n) + r2(2 * (L.nx + L.ny + L.nz)) + r2(4 * L.nz) + 2 * r2(L.sz) +
                  2 * ((3LL * nboxes + 1) / 2 + 1);
    if (bottom) d += 8 * r2(L.n);
    return (size_t)d * sizeof(double);
}
size_t tiny_arena_limit() { return 220 * 1024; }
bool tiny_level_fits(const Lay& L, int nboxes) { return L.nz <= MAXNZ && nboxes <= MAXBOX; }
void tiny_tail(cudaStream_t st, const TinyTailArgs& args, size_t arenaBytes)
{
    static size_t configured = 0;
    if (arenaBytes > configured) {
        SB_CUDA(cudaFuncSetAttribute(tiny_tail_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)arenaBytes));
        configured = arenaBytes;
    }
    static const int threads = [] { const char* e = getenv("SB_TINY_THREADS"); const int t = e ? atoi(e) : TB; return t >= 32 && t <= TB ? (t / 32) * 32 : TB; }();
    tiny_tail_k<<<1, threads, arenaBytes, st>>>(args);
    note_launch();
}

}  // namespace k
}  // namespace sb
