// sb_kernels.cu -- hand-written fp64 CUDA kernels (sm_100a) for SOMAR's pressure projection.
//
// Compiled with -fmad=false: every kernel evaluates its expression in the association order of
// the reference's Chombo-Fortran leaf (cited per kernel, path:line relative to
// /root/reference/src/Grade3_Calculus), so element-wise results agree with the CPU code to the
// last bit and only reductions differ (summation order).
//
// Layout: see sb_core.h (Lay).  x is contiguous; threadIdx.x always runs along x.
#include "sb_core.h"

namespace sb {
namespace k {

static long long g_launches = 0;
long long launch_count() { return g_launches; }
#define LAUNCHED() (++g_launches)
void note_launch() { ++g_launches; }
void note_launches(long long n) { g_launches += n; }

static inline dim3 grid3(int nx, int ny, int nz, dim3 b)
{
    return dim3((nx + b.x - 1) / b.x, (ny + b.y - 1) / b.y, (nz + b.z - 1) / b.z);
}
static const dim3 B3(64, 4, 1);

// ------------------------------------------------------------------------------------------
// BLAS-1 over valid cells (Elliptic/StateOps/LDFABOps.cpp:170-250, BoxTools/FArrayBox.cpp).
// faceDir >= 0 extends the range by one in that direction (face-centred data).
// ------------------------------------------------------------------------------------------
__global__ void fill_k(double* a, long long n, double v)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long s = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += s) a[i] = v;
}
void fill(cudaStream_t st, double* a, long long n, double v)
{
    if (v == 0.0) {
        cudaMemsetAsync(a, 0, n * sizeof(double), st);
    } else {
        fill_k<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 32), 256, 0, st>>>(a, n, v);
    }
    LAUNCHED();
}

template <class F>
__global__ void valid_k(Lay L, int ex, int ey, int ez, F f)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= L.nx + ex || j >= L.ny + ey || k >= L.nz + ez) return;
    f(L.idx(i, j, k), i, j, k);
}
template <class F>
static void launch_valid(cudaStream_t st, const Lay& L, int faceDir, F f)
{
    const int ex = faceDir == 0, ey = faceDir == 1, ez = faceDir == 2;
    valid_k<<<grid3(L.nx + ex, L.ny + ey, L.nz + ez, B3), B3, 0, st>>>(L, ex, ey, ez, f);
    LAUNCHED();
}

void copy_valid(cudaStream_t st, const Lay& L, double* dst, const double* src)
{
    launch_valid(st, L, -1, [=] __device__(long long c, int, int, int) { dst[c] = src[c]; });
}
void scale_valid(cudaStream_t st, const Lay& L, double* a, double s, int faceDir)
{
    launch_valid(st, L, faceDir, [=] __device__(long long c, int, int, int) { a[c] = a[c] * s; });
}
// FArrayBox::plus(x, scale): this += scale * x
void incr_valid(cudaStream_t st, const Lay& L, double* y, const double* x, double s, int faceDir)
{
    launch_valid(st, L, faceDir, [=] __device__(long long c, int, int, int) { y[c] = y[c] + s * x[c]; });
}
// Grade1_Basics/FABAlgebraF.ChF:145-165
void axby_valid(cudaStream_t st, const Lay& L, double* z, const double* x, const double* y, double a, double b)
{
    launch_valid(st, L, -1, [=] __device__(long long c, int, int, int) { z[c] = a * x[c] + b * y[c]; });
}
// PoissonOp::removeKernel (Elliptic/PoissonOp.cpp:838-842): phi -= sum/vol
void add_scalar_valid(cudaStream_t st, const Lay& L, double* a, const double* sumvol)
{
    launch_valid(st, L, -1, [=] __device__(long long c, int, int, int) {
        const double avg = sumvol[0] / sumvol[1];
        a[c]             = a[c] - avg;
    });
}
void mult_valid(cudaStream_t st, const Lay& L, double* dst, const double* src, const double* m)
{
    launch_valid(st, L, -1, [=] __device__(long long c, int, int, int) { dst[c] = src[c] * m[c]; });
}

// ------------------------------------------------------------------------------------------
// Ghost fill of one direction: physical Robin BC (BCToolsF.ChF:222-337, as driven by
// BCTools.cpp:334-415) or periodic wrap inside the tile.  ext = 1 extends the tangential
// range over the ghosts of the lower-numbered directions (used before the quadratic prolong,
// where edge ghosts matter: PoissonOp.cpp:1115-1125).
// ------------------------------------------------------------------------------------------
__global__ void fill_ghosts_dir_k(Lay L, double* phi, int dir, SideBC lo, SideBC hi, int ext0, int ext1)
{
    // (a, b) run over the two tangential directions in ascending order.
    int       na, nb;
    long long sa, sb, sn;
    int       nn;
    if (dir == 0) { na = L.ny; nb = L.nz; sa = L.sy; sb = L.sz; sn = 1; nn = L.nx; }
    else if (dir == 1) { na = L.nx; nb = L.nz; sa = 1; sb = L.sz; sn = L.sy; nn = L.ny; }
    else { na = L.nx; nb = L.ny; sa = 1; sb = L.sy; sn = L.sz; nn = L.nz; }
    const int a = blockIdx.x * blockDim.x + threadIdx.x - ext0;
    const int b = blockIdx.y * blockDim.y + threadIdx.y - ext1;
    if (a >= na + ext0 || b >= nb + ext1) return;
    const long long base = L.idx(0, 0, 0) + sa * a + sb * b;  // cell 0 along dir
    for (int side = 0; side < 2; ++side) {
        const SideBC&   bc = side ? hi : lo;
        const long long g  = side ? base + sn * nn : base - sn;             // ghost
        const long long p0 = side ? base + sn * (nn - 1) : base;            // first interior
        const long long p1 = side ? base + sn * (nn - 2) : base + sn;       // second interior
        if (sideIsBC(bc.kind)) {
            phi[g] = sideGhost(bc, phi[p0], phi[p1]);
        } else if (bc.kind == SIDE_PERIODIC_SELF) {
            phi[g] = side ? phi[base] : phi[base + sn * (nn - 1)];
        }
    }
}
void fill_ghosts_dir(cudaStream_t st, const Lay& L, double* phi, int dir, const SideBC& lo, const SideBC& hi, int ext0,
                     int ext1)
{
    if (lo.kind == SIDE_NEIGHBOR && hi.kind == SIDE_NEIGHBOR) return;
    int na, nb;
    if (dir == 0) { na = L.ny; nb = L.nz; }
    else if (dir == 1) { na = L.nx; nb = L.nz; }
    else { na = L.nx; nb = L.ny; }
    dim3 b(dir == 0 ? 8 : 64, dir == 0 ? 16 : 2, 1);
    fill_ghosts_dir_k<<<grid3(na + 2 * ext0, nb + 2 * ext1, 1, b), b, 0, st>>>(L, phi, dir, lo, hi, ext0, ext1);
    LAUNCHED();
}
void fill_ghosts(cudaStream_t st, const Lay& L, double* phi, const SideBC bc[3][2], int dim)
{
    for (int d = 0; d < 3; ++d) {
        if (dim == 2 && d == 1) continue;
        fill_ghosts_dir(st, L, phi, d, bc[d][0], bc[d][1], 0, 0);
    }
}

// Edge ghosts where two physical sides meet: BCTools::extrapCorners order 2
// (BCTools.cpp:66-150 -> BCToolsF.ChF:123-196), applied to the domain-corner box; everywhere
// else the reference's CornerCopier leaves the neighbouring box's face ghost there, which in
// the fused tile is what the direction-by-direction fill already produced.
__global__ void extrap_edge_k(Lay L, double* phi, int adir, int aside, int bdir, int bside)
{
    const int cdir = 3 - adir - bdir;
    const int nc   = cdir == 0 ? L.nx : (cdir == 1 ? L.ny : L.nz);
    const int t    = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nc) return;
    const long long sd[3] = {1, L.sy, L.sz};
    const int       nd[3] = {L.nx, L.ny, L.nz};
    int             ijk[3];
    ijk[cdir] = t;
    ijk[adir] = aside ? nd[adir] : -1;
    ijk[bdir] = bside ? nd[bdir] : -1;
    const long long e  = L.idx(ijk[0], ijk[1], ijk[2]);
    const long long ai = aside ? -sd[adir] : sd[adir];  // step inward along adir
    const long long bi = bside ? -sd[bdir] : sd[bdir];
    const double aval  = 3.0 * phi[e + ai] - 3.0 * phi[e + 2 * ai] + phi[e + 3 * ai];
    const double bval  = 3.0 * phi[e + bi] - 3.0 * phi[e + 2 * bi] + phi[e + 3 * bi];
    phi[e]             = 0.5 * (aval + bval);
}
// Fill every ghost a 9-point-per-plane stencil can touch (faces + edges).
void fill_ghosts_with_edges(cudaStream_t st, const Lay& L, double* phi, const SideBC bc[3][2], int dim)
{
    fill_ghosts_dir(st, L, phi, 0, bc[0][0], bc[0][1], 0, 0);
    if (dim == 3) fill_ghosts_dir(st, L, phi, 1, bc[1][0], bc[1][1], 1, 0);
    fill_ghosts_dir(st, L, phi, 2, bc[2][0], bc[2][1], 1, dim == 3 ? 1 : 0);
    extrap_domain_edges(st, L, phi, bc, dim);
}
void extrap_domain_edges(cudaStream_t st, const Lay& L, double* phi, const SideBC bc[3][2], int dim)
{
    for (int a = 0; a < 3; ++a)
        for (int b = a + 1; b < 3; ++b) {
            if (dim == 2 && (a == 1 || b == 1)) continue;
            for (int as = 0; as < 2; ++as)
                for (int bs = 0; bs < 2; ++bs) {
                    if (!sideIsBC(bc[a][as].kind) || !sideIsBC(bc[b][bs].kind)) continue;
                    const int c  = 3 - a - b;
                    const int nc = c == 0 ? L.nx : (c == 1 ? L.ny : L.nz);
                    extrap_edge_k<<<(nc + 127) / 128, 128, 0, st>>>(L, phi, a, as, b, bs);
                    LAUNCHED();
                }
        }
}

// Pack / unpack one face layer (valid cells adjacent to the side -> buffer; buffer -> ghosts).
// Stands in for Copier motion items of LevelData::exchange (BoxTools/BoxLayoutDataI.H:665-812).
// ext0 / ext1 extend the tangential ranges by one ghost on each end (edge exchange before the
// quadratic prolongation); the buffer is (na + 2 ext0) x (nb + 2 ext1).
__global__ void pack_face_k(Lay L, const double* phi, int dir, int layer, double* buf, int unpack, double* phiw, int ext0,
                            int ext1)
{
    int na, nb;
    long long sa, sb, sn;
    if (dir == 0) { na = L.ny; nb = L.nz; sa = L.sy; sb = L.sz; sn = 1; }
    else if (dir == 1) { na = L.nx; nb = L.nz; sa = 1; sb = L.sz; sn = L.sy; }
    else { na = L.nx; nb = L.ny; sa = 1; sb = L.sy; sn = L.sz; }
    const int a = blockIdx.x * blockDim.x + threadIdx.x - ext0;
    const int b = blockIdx.y * blockDim.y + threadIdx.y - ext1;
    if (a >= na + ext0 || b >= nb + ext1) return;
    const long long c = L.idx(0, 0, 0) + sa * a + sb * b + sn * layer;
    const long long m = (a + ext0) + (long long)(na + 2 * ext0) * (b + ext1);
    if (unpack) phiw[c] = buf[m];
    else buf[m] = phi[c];
}
size_t face_count(const Lay& L, int dir, int ext0, int ext1)
{
    const int na = dir == 0 ? L.ny : L.nx, nb = dir == 2 ? L.ny : L.nz;
    return (size_t)(na + 2 * ext0) * (nb + 2 * ext1);
}
void pack_face(cudaStream_t st, const Lay& L, const double* phi, int dir, int side, double* buf, int ext0, int ext1)
{
    int na, nb, nn;
    if (dir == 0) { na = L.ny; nb = L.nz; nn = L.nx; }
    else if (dir == 1) { na = L.nx; nb = L.nz; nn = L.ny; }
    else { na = L.nx; nb = L.ny; nn = L.nz; }
    dim3 b(dir == 0 ? 8 : 64, dir == 0 ? 16 : 2, 1);
    pack_face_k<<<grid3(na + 2 * ext0, nb + 2 * ext1, 1, b), b, 0, st>>>(L, phi, dir, side ? nn - 1 : 0, buf, 0, nullptr, ext0, ext1);
    LAUNCHED();
}
void unpack_face(cudaStream_t st, const Lay& L, double* phi, int dir, int side, const double* buf, int ext0, int ext1)
{
    int na, nb, nn;
    if (dir == 0) { na = L.ny; nb = L.nz; nn = L.nx; }
    else if (dir == 1) { na = L.nx; nb = L.nz; nn = L.ny; }
    else { na = L.nx; nb = L.ny; nn = L.nz; }
    dim3 b(dir == 0 ? 8 : 64, dir == 0 ? 16 : 2, 1);
    pack_face_k<<<grid3(na + 2 * ext0, nb + 2 * ext1, 1, b), b, 0, st>>>(L, nullptr, dir, side ? nn : -1, const_cast<double*>(buf), 1,
                                                                          phi, ext0, ext1);
    LAUNCHED();
}

// ------------------------------------------------------------------------------------------
// Operator: lhs = beta*J*Sum_d(M_dL phi_- + M_dR phi_+) + phi/Dinv
// (Elliptic/PoissonOpF.ChF:151-199), and the residual rhs - lhs (LevelOperator.H:104-117:
// applyOp, scale(-1), incr(rhs, 1) -- (-lhs) + rhs is bit-identical to rhs - lhs).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double stencil7(const Lay& L, const Coef& c, const double* __restrict__ phi, long long q, int i,
                                           int j, int k)
{
    double s = c.mxl[i] * phi[q - 1] + c.mxr[i] * phi[q + 1] + c.myl[j] * phi[q - L.sy] + c.myr[j] * phi[q + L.sy];
    s        = s + c.mzl[k] * phi[q - L.sz] + c.mzr[k] * phi[q + L.sz];
    return s;
}
__global__ void apply_op_k(Lay L, Coef c, double* __restrict__ out, const double* __restrict__ phi,
                           const double* __restrict__ rhs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= L.nx || j >= L.ny) return;
    const long long q   = L.idx(i, j, k);
    const double    lap = stencil7(L, c, phi, q, i, j, k);
    const double    Jq  = c.tabJ ? c.tabJ[k] : c.J[q];        // the same bits either way (Op::detectColumnCoefficients)
    const double    Dq  = c.tabD ? c.tabD[k] : c.Dinv[q];
    const double    lhs = c.beta * Jq * lap + phi[q] / Dq;
    out[q]              = rhs ? rhs[q] - lhs : lhs;
}
void apply_op(cudaStream_t st, const Lay& L, const Coef& c, double* lhs, const double* phi)
{
    apply_op_k<<<grid3(L.nx, L.ny, L.nz, B3), B3, 0, st>>>(L, c, lhs, phi, nullptr);
    LAUNCHED();
}
void residual(cudaStream_t st, const Lay& L, const Coef& c, double* res, const double* phi, const double* rhs)
{
    apply_op_k<<<grid3(L.nx, L.ny, L.nz, B3), B3, 0, st>>>(L, c, res, phi, rhs);
    LAUNCHED();
}

// Point red-black Gauss-Seidel, one colour: phi = (rhs - beta*J*S[phi]) * Dinv on cells with
// (i+j+k+pass) even in global indices (PoissonOpF.ChF:420-474, DO_RBPASS :27-46).
__global__ void gsrb_k(Lay L, Coef c, double* __restrict__ phi, const double* __restrict__ rhs, int pass)
{
    const int j  = blockIdx.y * blockDim.y + threadIdx.y;
    const int k  = blockIdx.z;
    const int i0 = (L.lo0 + L.lo1 + j + L.lo2 + k + pass) & 1;
    const int i  = i0 + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= L.nx || j >= L.ny) return;
    const long long q = L.idx(i, j, k);
    const double    s = stencil7(L, c, phi, q, i, j, k);
    phi[q]            = (rhs[q] - c.beta * c.J[q] * s) * c.Dinv[q];
}
void gsrb_pass(cudaStream_t st, const Lay& L, const Coef& c, double* phi, const double* rhs, int pass)
{
    gsrb_k<<<grid3((L.nx + 1) / 2, L.ny, L.nz, B3), B3, 0, st>>>(L, c, phi, rhs, pass);
    LAUNCHED();
}

// Jacobi / red-black Jacobi update phi += Dinv*res (PoissonOpF.ChF:264-310).
__global__ void jacobi_k(Lay L, Coef c, double* __restrict__ phi, const double* __restrict__ res, int pass)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= L.nx || j >= L.ny) return;
    if (pass >= 0 && ((L.lo0 + i + L.lo1 + j + L.lo2 + k + pass) & 1)) return;
    const long long q = L.idx(i, j, k);
    phi[q]            = phi[q] + c.Dinv[q] * res[q];
}
void jacobi(cudaStream_t st, const Lay& L, const Coef& c, double* phi, const double* res, int pass)
{
    jacobi_k<<<grid3(L.nx, L.ny, L.nz, B3), B3, 0, st>>>(L, c, phi, res, pass);
    LAUNCHED();
}

// ------------------------------------------------------------------------------------------
// Vertical line relaxation, one colour (PoissonOpF.ChF:851-1019 + LAPACK dgtsv, NRHS = 1).
// One thread owns one column with (i+j+pass) even.  The elimination follows dgtsv's
// no-interchange branch operation by operation; if dgtsv would have interchanged rows
// (|d| < |dl|) the flag is raised so the host can fail loudly instead of drifting.
// v1: modified diagonal d' and rhs b' are parked in compact scratch (wd, wb).
// ------------------------------------------------------------------------------------------
__global__ void vertline_k(Lay L, Coef c, double* __restrict__ phi, const double* __restrict__ rhs, int pass,
                           double* __restrict__ wd, double* __restrict__ wb, int* pivotFlag)
{
    const int hx = (L.nx + 1) / 2;
    const int ic = blockIdx.x * blockDim.x + threadIdx.x;
    const int j  = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= L.ny) return;
    const int i = ((L.lo0 + L.lo1 + j + pass) & 1) + 2 * ic;
    if (i >= L.nx) return;
    const int       N   = L.nz;
    const double    mxl = c.mxl[i], mxr = c.mxr[i], myl = c.myl[j], myr = c.myr[j];
    const long long q0  = L.idx(i, j, 0);
    const long long s2  = L.idx(i, j, -1) - L.sz + L.sz;  // slab index (k = -1 plane offset removed below)
    const long long slab = (long long)(OX + i) + L.sy * (long long)(1 + j);
    (void)s2;
    const long long ws = (long long)hx * L.ny;  // scratch plane stride
    long long       w  = ic + (long long)hx * j;

    // row k = 0
    long long q    = q0;
    double    Jb   = c.J[q] * c.beta;  // Jval * beta
    double    lphi = mxl * phi[q - 1] + mxr * phi[q + 1] + myl * phi[q - L.sy] + myr * phi[q + L.sy];
    double    b    = rhs[q] - Jb * lphi;
    double    d    = 1.0 / c.Dinv[q] + c.loBC[slab];
    if (N == 1) {
        d = d + c.hiBC[slab];
        if (d == 0.0) { atomicOr(pivotFlag, 2); return; }  // dgtsv INFO = N: reference leaves B unsolved
        phi[q] = b / d;
        return;
    }
    double du = c.beta * c.J[q] * c.mzr[0];
    int    bad = 0;
    for (int k = 0; k < N - 1; ++k) {
        // next row's raw entries
        const long long qn = q + L.sz;
        const double    Jn = c.J[qn];
        lphi               = mxl * phi[qn - 1] + mxr * phi[qn + 1] + myl * phi[qn - L.sy] + myr * phi[qn + L.sy];
        double bn          = rhs[qn] - Jn * c.beta * lphi;
        double dn          = 1.0 / c.Dinv[qn];
        if (k + 1 == N - 1) dn = dn + c.hiBC[slab];
        const double dl = c.beta * Jn * c.mzl[k + 1];
        // dgtsv: if |d(i)| >= |dl(i)| ... fact = dl/d; d(i+1) -= fact*du(i); b(i+1) -= fact*b(i)
        if (!(fabs(d) >= fabs(dl)) || d == 0.0) bad = 1;
        const double fact = dl / d;
        wd[w]             = d;
        wb[w]             = b;
        dn                = dn - fact * du;
        bn                = bn - fact * b;
        d                 = dn;
        b                 = bn;
        du                = c.beta * Jn * c.mzr[k + 1];
        q                 = qn;
        w += ws;
    }
    if (d == 0.0) bad = 1;
    if (bad) atomicOr(pivotFlag, 1);
    // back substitution: b(N) /= d(N); b(i) = (b(i) - du(i)*b(i+1)) / d(i)
    double x = b / d;
    phi[q]   = x;
    for (int k = N - 2; k >= 0; --k) {
        q -= L.sz;
        w -= ws;
        const double duk = c.beta * c.J[q] * c.mzr[k];
        x                = (wb[w] - duk * x) / wd[w];
        phi[q]           = x;
    }
}
void vertline_pass(cudaStream_t st, const Lay& L, const Coef& c, double* phi, const double* rhs, int pass, double* wd,
                   double* wb, int* pivotFlag)
{
    const int hx = (L.nx + 1) / 2;
    dim3      b(hx >= 64 ? 64 : 32, hx >= 64 ? 2 : 4, 1);
    vertline_k<<<grid3(hx, L.ny, 1, b), b, 0, st>>>(L, c, phi, rhs, pass, wd, wb, pivotFlag);
    LAUNCHED();
}

// ------------------------------------------------------------------------------------------
// Vertical line relaxation, fast path for columns that all share one tridiagonal matrix.
//
// Dividing row k of the reference's column system (PoissonOpF.ChF:905-955) by beta*J_k gives
//   MzL_k x_{k-1} + (alpha/beta - h - MzL_k - MzR_k + bc_k) x_k + MzR_k x_{k+1}
//        = rhs_k / (beta J_k) - (MxL phi_W + MxR phi_E + MyL phi_S + MyR phi_N),
// with h = MxL+MxR+MyL+MyR of the column.  When the horizontal metric is uniform (h and
// J/Jz(k) do not depend on i, j -- every Cartesian-horizontal grid, stretched or not in z) the
// matrix is the same for every column and its Thomas factorisation is a 1-D table built once
// per MG depth on the host (tab: s = 1/(beta J_k), f_k = MzL_{k+1}/d'_k, g_k = 1/d'_k, MzR_k).
// A CTA owns 32 columns of one colour in one grid row: all warps build the right-hand sides
// into shared memory (x-contiguous loads, every load independent), one warp runs the two
// Thomas sweeps out of shared memory, so no intermediate ever goes to HBM and J/Dinv are not
// read at all.  Same mathematics as dgtsv's no-pivot branch; rounding differs at 1e-16.
// ------------------------------------------------------------------------------------------
template <int NW, int U, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) vertline_smem_k(Lay L, Coef c, const double* __restrict__ tab,
                                                          double* __restrict__ phi, const double* __restrict__ rhs, int pass, int dbg)
{
    extern __shared__ double sm[];  // [N][32] right-hand sides, then tables -f, g, -(MzR g) (3*N)
    const int     N    = L.nz;
    double* const sf   = sm + (size_t)N * 32;
    double* const sg   = sf + N;
    double* const sc   = sg + N;
    const int     lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int     j    = blockIdx.y;
    const int     i    = ((L.lo0 + L.lo1 + j + pass) & 1) + 2 * (blockIdx.x * 32 + lane);
    const bool    act  = i < L.nx;
    const int     ii   = act ? i : 0;
    const double  mxl = c.mxl[ii], mxr = c.mxr[ii], myl = c.myl[j], myr = c.myr[j];
    const double* ts  = tab;
    for (int k = threadIdx.x; k < N; k += NW * 32) {
        sf[k] = tab[N + k];
        sg[k] = tab[2 * N + k];
        sc[k] = tab[3 * N + k];
    }
    const long long sz = L.sz, sy = L.sy;
    const double*   pq = phi + L.idx(ii, j, 0);
    const double*   rq = rhs + L.idx(ii, j, 0);

    // Phase 1: right-hand sides b_k = rhs_k s_k - (MxL phi_W + MxR phi_E + MyL phi_S + MyR phi_N).
    // Each warp takes U consecutive levels per trip; the 5*U loads of trip t+1 are issued before
    // trip t is consumed (register double buffer), so loads are in flight all the time.
    double a0[U][5], a1[U][5];
    auto   issue = [&](double(&a)[U][5], int k0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int       k = min(k0 + u, N - 1);
            const long long o = (long long)k * sz;
            if (dbg != 2) { a[u][0] = pq[o - 1]; a[u][1] = pq[o + 1]; a[u][2] = pq[o - sy]; a[u][3] = pq[o + sy]; a[u][4] = rq[o]; }
            else { a[u][0] = a[u][1] = a[u][2] = a[u][3] = a[u][4] = (double)k; }
        }
    };
    auto consume = [&](double(&a)[U][5], int k0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = k0 + u;
            if (k < N) {
                const double lphi = mxl * a[u][0] + mxr * a[u][1] + myl * a[u][2] + myr * a[u][3];
                sm[k * 32 + lane] = act ? a[u][4] * ts[k] - lphi : 0.0;
            }
        }
    };
    {
        int k0 = w * U;
        if (k0 < N) issue(a0, k0);
        while (k0 < N) {
            const int k1 = k0 + NW * U;
            if (k1 < N) issue(a1, k1);
            consume(a0, k0);
            k0 = k1;
            if (k0 >= N) break;
            const int k2 = k0 + NW * U;
            if (k2 < N) issue(a0, k2);
            consume(a1, k0);
            k0 = k2;
        }
    }
    __syncthreads();
    if (w != 0 || dbg == 1) return;

    // Phase 2 (one warp, one column per lane): y_k = b_k - f_{k-1} y_{k-1}, z_k = y_k g_k parked
    // in place; then x_{N-1} = z_{N-1}, x_k = z_k - (MzR_k g_k) x_{k+1}.  Batches of T levels:
    // all shared-memory operands of a batch are in registers before its FMA chain starts.
    constexpr int T = 8;
    double*       col = sm + lane;
    double        y   = col[0];
    col[0]            = y * sg[0];
    int k = 1;
    for (; k + T <= N; k += T) {
        double bb[T], ff[T], gg[T], zz[T];
#pragma unroll
        for (int u = 0; u < T; ++u) { bb[u] = col[(k + u) * 32]; ff[u] = sf[k + u - 1]; gg[u] = sg[k + u]; }
#pragma unroll
        for (int u = 0; u < T; ++u) { y = fma(ff[u], y, bb[u]); zz[u] = y * gg[u]; }
#pragma unroll
        for (int u = 0; u < T; ++u) col[(k + u) * 32] = zz[u];
    }
    for (; k < N; ++k) {
        y           = fma(sf[k - 1], y, col[k * 32]);
        col[k * 32] = y * sg[k];
    }
    __syncwarp();
    double  x = col[(N - 1) * 32];
    double* p = phi + L.idx(ii, j, N - 1);
    if (act) *p = x;
    k = N - 2;
    for (; k - (T - 1) >= 0; k -= T) {
        double zb[T], cg[T], xx[T];
#pragma unroll
        for (int u = 0; u < T; ++u) { zb[u] = col[(k - u) * 32]; cg[u] = sc[k - u]; }
#pragma unroll
        for (int u = 0; u < T; ++u) { x = fma(cg[u], x, zb[u]); xx[u] = x; }
        if (act) {
#pragma unroll
            for (int u = 0; u < T; ++u) p[-(long long)(u + 1) * sz] = xx[u];
        }
        p -= (long long)T * sz;
    }
    for (; k >= 0; --k) {
        x = fma(sc[k], x, col[k * 32]);
        p -= sz;
        if (act) *p = x;
    }
}

size_t vertline_smem_bytes(int nz) { return ((size_t)nz * 32 + 3 * (size_t)nz) * sizeof(double); }
void vertline_smem_pass(cudaStream_t st, const Lay& L, const Coef& c, const double* tab, double* phi, const double* rhs,
                        int pass)
{
    static int dbg = -1; if (dbg < 0) { const char* e = getenv("SB_LINE_DEBUG"); dbg = e ? atoi(e) : 0; }
    const int    hx = (L.nx + 1) / 2;
    const size_t sh = vertline_smem_bytes(L.nz);
    static int   variant = -1;
    if (variant < 0) {
        const char* e = getenv("SB_LINE_VARIANT");  // development knob: 0 = <8,4,2>, 1 = <4,8,2>, 2 = <8,2,3>
        variant       = e ? atoi(e) : 0;
    }
#define SB_LAUNCH_VL(NW, U, MB)                                                                                         \
    {                                                                                                                   \
        static size_t configured = 0;                                                                                   \
        if (sh > configured) {                                                                                          \
            cudaFuncSetAttribute(vertline_smem_k<NW, U, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);     \
            configured = sh;                                                                                            \
        }                                                                                                               \
        vertline_smem_k<NW, U, MB><<<dim3((hx + 31) / 32, L.ny), NW * 32, sh, st>>>(L, c, tab, phi, rhs, pass, dbg);        \
    }
    if (variant == 1) SB_LAUNCH_VL(4, 8, 2)
    else if (variant == 2) SB_LAUNCH_VL(8, 2, 3)
    else SB_LAUNCH_VL(8, 4, 2)
#undef SB_LAUNCH_VL
    LAUNCHED();
}
// max over the tile of |J(i,j,k) - Jcol[k]| / |Jcol[k]| (is the metric a function of z only?)
__global__ void jdev_k(Lay L, const double* __restrict__ J, const double* __restrict__ jcol, double* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    double    v = 0.0;
    if (i < L.nx && j < L.ny) v = fabs(J[L.idx(i, j, k)] - jcol[k]) / fabs(jcol[k]);
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v > 0.0) atomicMax((unsigned long long*)out, (unsigned long long)__double_as_longlong(v));
}
void j_deviation(cudaStream_t st, const Lay& L, const double* J, const double* jcol, double* out)
{
    cudaMemsetAsync(out, 0, sizeof(double), st);
    jdev_k<<<grid3(L.nx, L.ny, L.nz, B3), B3, 0, st>>>(L, J, jcol, out);
    LAUNCHED();
}

// ------------------------------------------------------------------------------------------
// Restriction: block average, sum in Fortran loop order (CFInterpF.ChF:1085-1120).
// ------------------------------------------------------------------------------------------
__global__ void restrict_k(Lay Lf, Lay Lc, int r0, int r1, int r2, double* __restrict__ crse, const double* __restrict__ fine)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= Lc.nx || j >= Lc.ny) return;
    const double refScale = 1.0 / (double)(r0 * r1 * r2);
    double       s        = 0.0;
    for (int c = 0; c < r2; ++c)
        for (int b = 0; b < r1; ++b)
            for (int a = 0; a < r0; ++a) s = s + fine[Lf.idx(i * r0 + a, j * r1 + b, k * r2 + c)];
    crse[Lc.idx(i, j, k)] = s * refScale;
}
void restrict_avg(cudaStream_t st, const Lay& Lf, const Lay& Lc, const int ref[3], double* crse, const double* fine)
{
    restrict_k<<<grid3(Lc.nx, Lc.ny, Lc.nz, B3), B3, 0, st>>>(Lf, Lc, ref[0], ref[1], ref[2], crse, fine);
    LAUNCHED();
}
// Face-centred average (CFInterpF.ChF:1221-1258): Jgup coarsening at MG setup.
__global__ void restrict_face_k(Lay Lf, Lay Lc, int r0, int r1, int r2, int dir, double* __restrict__ crse,
                                const double* __restrict__ fine)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= Lc.nx + (dir == 0) || j >= Lc.ny + (dir == 1) || k >= Lc.nz + (dir == 2)) return;
    const int    rd       = dir == 0 ? r0 : (dir == 1 ? r1 : r2);
    const double refScale = (double)rd / (double)(r0 * r1 * r2);
    const int    n0 = dir == 0 ? 1 : r0, n1 = dir == 1 ? 1 : r1, n2 = dir == 2 ? 1 : r2;
    double       s = 0.0;
    for (int c = 0; c < n2; ++c)
        for (int b = 0; b < n1; ++b)
            for (int a = 0; a < n0; ++a) s = s + fine[Lf.idx(i * r0 + a, j * r1 + b, k * r2 + c)];
    crse[Lc.idx(i, j, k)] = refScale * s;
}
void restrict_face(cudaStream_t st, const Lay& Lf, const Lay& Lc, const int ref[3], int dir, double* crse, const double* fine)
{
    restrict_face_k<<<grid3(Lc.nx + 1, Lc.ny + 1, Lc.nz + 1, B3), B3, 0, st>>>(Lf, Lc, ref[0], ref[1], ref[2], dir, crse, fine);
    LAUNCHED();
}

// ------------------------------------------------------------------------------------------
// Prolongation fine += I(crse) (PoissonOpF.ChF:1031-1320).  One thread per fine cell.
// order 0: constant; order 1 adds the slope terms in the same kernel (the running sum is
// evaluated in the reference's order: ((fine + crse) + dxf0*m0) + dxf1*m1) + dxf2*m2).
// ------------------------------------------------------------------------------------------
__global__ void prolong_k(Lay Lf, Lay Lc, int r0, int r1, int r2, double* __restrict__ fine, const double* __restrict__ crse,
                          int order)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= Lf.nx || j >= Lf.ny) return;
    const int       ic = i / r0, jc = j / r1, kc = k / r2;
    const long long qc = Lc.idx(ic, jc, kc);
    const long long qf = Lf.idx(i, j, k);
    double          f  = fine[qf];
    if (order == 0 || order == 1) f = f + crse[qc];
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
    if (order == 2) {  // QuadUpgrade1 :1203-1245
        const double mid  = -2.0 * crse[qc];
        const double mm0  = 0.25 * (crse[qc + 1] + mid + crse[qc - 1]);
        const double mm1  = 0.25 * (crse[qc + Lc.sy] + mid + crse[qc - Lc.sy]);
        const double mm2  = 0.25 * (crse[qc + Lc.sz] + mid + crse[qc - Lc.sz]);
        const double dxf0 = -0.5 + (((i - ic * r0) + 0.5) / r0);
        const double dxf1 = -0.5 + (((j - jc * r1) + 0.5) / r1);
        const double dxf2 = -0.5 + (((k - kc * r2) + 0.5) / r2);
        f                 = f + dxf0 * dxf0 * mm0 + dxf1 * dxf1 * mm1 + dxf2 * dxf2 * mm2;
    }
    fine[qf] = f;
}
// QuadUpgrade2 :1249-1322 (mixed second differences; note the reference's own sign pattern for
// the x-y pair, and that the 2-D build uses that same pattern for its only pair).
__global__ void prolong_quad2_k(Lay Lf, Lay Lc, int r0, int r1, int r2, double* __restrict__ fine,
                                const double* __restrict__ crse, int dim)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= Lf.nx || j >= Lf.ny) return;
    const int       ic = i / r0, jc = j / r1, kc = k / r2;
    const long long qc = Lc.idx(ic, jc, kc);
    const long long qf = Lf.idx(i, j, k);
    const long long sy = Lc.sy, sz = Lc.sz;
    const double    dxf0 = -0.5 + (((i - ic * r0) + 0.5) / r0);
    const double    dxf1 = -0.5 + (((j - jc * r1) + 0.5) / r1);
    const double    dxf2 = -0.5 + (((k - kc * r2) + 0.5) / r2);
    if (dim == 2) {
        // CH_SPACEDIM == 2 branch: directions (x, z) here.
        const double mm0 = 0.25 * (crse[qc + 1 + sz] + crse[qc + 1 - sz] - crse[qc - 1 + sz] - crse[qc - 1 - sz]);
        fine[qf]         = fine[qf] + dxf0 * dxf2 * mm0;
    } else {
        const double mm0 = 0.25 * (crse[qc + sy + sz] - crse[qc + sy - sz] - crse[qc - sy + sz] + crse[qc - sy - sz]);
        const double mm1 = 0.25 * (crse[qc + 1 + sz] - crse[qc + 1 - sz] - crse[qc - 1 + sz] + crse[qc - 1 - sz]);
        const double mm2 = 0.25 * (crse[qc + 1 + sy] + crse[qc + 1 - sy] - crse[qc - 1 + sy] - crse[qc - 1 - sy]);
        fine[qf]         = fine[qf] + dxf1 * dxf2 * mm0 + dxf2 * dxf0 * mm1 + dxf0 * dxf1 * mm2;
    }
}
// Fast path for refinement ratios of 1 or 2 (all the MG schedules use): one thread per coarse
// cell of a plane updates its R0 x r1 children, so the coarse stencil and the slopes are formed
// once; the expression per fine cell is the one of prolong_k, operation for operation.
template <int R0>
__global__ void prolong_c_k(Lay Lf, Lay Lc, int r1, int r2, double* __restrict__ fine, const double* __restrict__ crse, int order)
{
    const int ic = blockIdx.x * blockDim.x + threadIdx.x;
    const int jc = blockIdx.y * blockDim.y + threadIdx.y;
    const int kc = blockIdx.z;
    if (ic >= Lc.nx || jc >= Lc.ny) return;
    const long long qc = Lc.idx(ic, jc, kc);
    const double    c  = crse[qc];
    double          m0 = 0.0, m1 = 0.0, m2 = 0.0;
    const double    sc0 = R0 == 1 ? 0.0 : 1.0 / R0, sc1 = r1 == 1 ? 0.0 : 1.0 / r1, sc2 = r2 == 1 ? 0.0 : 1.0 / r2;
    if (order == 1) {
        const long long s0 = R0 == 1 ? 0 : 1, s1 = r1 == 1 ? 0 : Lc.sy, s2 = r2 == 1 ? 0 : Lc.sz;
        m0 = 0.5 * (crse[qc + s0] - crse[qc - s0]);
        m1 = 0.5 * (crse[qc + s1] - crse[qc - s1]);
        m2 = 0.5 * (crse[qc + s2] - crse[qc - s2]);
    } else if (order == 2) {
        const double mid = -2.0 * c;
        m0 = 0.25 * (crse[qc + 1] + mid + crse[qc - 1]);
        m1 = 0.25 * (crse[qc + Lc.sy] + mid + crse[qc - Lc.sy]);
        m2 = 0.25 * (crse[qc + Lc.sz] + mid + crse[qc - Lc.sz]);
    }
    for (int kk = 0; kk < r2; ++kk) {
        const int    k    = kc * r2 + kk;
        const double dxf2 = order == 2 ? -0.5 + ((kk + 0.5) / r2) : -0.5 + ((kk + 0.5) * sc2);
        for (int jj = 0; jj < r1; ++jj) {
            const double    dxf1 = order == 2 ? -0.5 + ((jj + 0.5) / r1) : -0.5 + ((jj + 0.5) * sc1);
            const long long qf   = Lf.idx(ic * R0, jc * r1 + jj, k);
            double          f[R0];
            if (R0 == 2) { const double2 v = *reinterpret_cast<const double2*>(fine + qf); f[0] = v.x; f[R0 - 1] = v.y; }
            else f[0] = fine[qf];
#pragma unroll
            for (int ii = 0; ii < R0; ++ii) {
                double g = f[ii];
                if (order == 0 || order == 1) g = g + c;
                if (order == 1) {
                    const double dxf0 = -0.5 + ((ii + 0.5) * sc0);
                    g                 = g + dxf0 * m0 + dxf1 * m1 + dxf2 * m2;
                }
                if (order == 2) {
                    const double dxf0 = -0.5 + ((ii + 0.5) / R0);
                    g                 = g + dxf0 * dxf0 * m0 + dxf1 * dxf1 * m1 + dxf2 * dxf2 * m2;
                }
                f[ii] = g;
            }
            if (R0 == 2) *reinterpret_cast<double2*>(fine + qf) = make_double2(f[0], f[R0 - 1]);
            else fine[qf] = f[0];
        }
    }
}
static void launch_prolong(cudaStream_t st, const Lay& Lf, const Lay& Lc, const int ref[3], double* fine, const double* crse, int order)
{
    const bool fast = (ref[0] == 1 || ref[0] == 2) && (ref[1] == 1 || ref[1] == 2) && (ref[2] == 1 || ref[2] == 2) &&
                      Lc.nx * ref[0] == Lf.nx && Lc.ny * ref[1] == Lf.ny && Lc.nz * ref[2] == Lf.nz;
    if (!fast) {
        prolong_k<<<grid3(Lf.nx, Lf.ny, Lf.nz, B3), B3, 0, st>>>(Lf, Lc, ref[0], ref[1], ref[2], fine, crse, order);
    } else if (ref[0] == 2) {
        prolong_c_k<2><<<grid3(Lc.nx, Lc.ny, Lc.nz, B3), B3, 0, st>>>(Lf, Lc, ref[1], ref[2], fine, crse, order);
    } else {
        prolong_c_k<1><<<grid3(Lc.nx, Lc.ny, Lc.nz, B3), B3, 0, st>>>(Lf, Lc, ref[1], ref[2], fine, crse, order);
    }
    LAUNCHED();
}
void prolong_const(cudaStream_t st, const Lay& Lf, const Lay& Lc, const int ref[3], double* fine, const double* crse)
{
    launch_prolong(st, Lf, Lc, ref, fine, crse, 0);
}
void prolong_linear(cudaStream_t st, const Lay& Lf, const Lay& Lc, const int ref[3], double* fine, const double* crse)
{
    launch_prolong(st, Lf, Lc, ref, fine, crse, 1);
}
void prolong_quad1(cudaStream_t st, const Lay& Lf, const Lay& Lc, const int ref[3], double* fine, const double* crse)
{
    launch_prolong(st, Lf, Lc, ref, fine, crse, 2);
}
void prolong_quad2(cudaStream_t st, const Lay& Lf, const Lay& Lc, const int ref[3], double* fine, const double* crse, int dim)
{
    prolong_quad2_k<<<grid3(Lf.nx, Lf.ny, Lf.nz, B3), B3, 0, st>>>(Lf, Lc, ref[0], ref[1], ref[2], fine, crse, dim);
    LAUNCHED();
}

// ------------------------------------------------------------------------------------------
// Setup: Dinv = 1/(J*(alpha - beta*(MxL+MxR+yzDiags))) (PoissonOpF.ChF:106-144) and the
// Robin terms folded into the tridiagonal end rows (PoissonOpF.ChF:635-694).
// ------------------------------------------------------------------------------------------
__global__ void dinv_k(Lay L, Coef c, double alpha, double* __restrict__ Dinv, int dim)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= L.nx || j >= L.ny) return;
    double yz;
    if (dim == 3) yz = c.myl[j] + c.myr[j] + c.mzl[k] + c.mzr[k];
    else yz = c.mzl[k] + c.mzr[k];  // 2-D build: "My" is the vertical
    const long long q = L.idx(i, j, k);
    Dinv[q]           = 1.0 / (c.J[q] * (alpha - c.beta * (c.mxl[i] + c.mxr[i] + yz)));
}
void compute_dinv(cudaStream_t st, const Lay& L, const Coef& c, double alpha, double* Dinv, int dim)
{
    dinv_k<<<grid3(L.nx, L.ny, L.nz, B3), B3, 0, st>>>(L, c, alpha, Dinv, dim);
    LAUNCHED();
}
__global__ void vert_bcs_k(Lay L, Coef c, double sLo, double sHi, double* __restrict__ loBC, double* __restrict__ hiBC)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= L.nx || j >= L.ny) return;
    const long long slab = (long long)(OX + i) + L.sy * (long long)(1 + j);
    loBC[slab]           = -c.beta * c.J[L.idx(i, j, 0)] * c.mzl[0] * sLo;
    hiBC[slab]           = -c.beta * c.J[L.idx(i, j, L.nz - 1)] * c.mzr[L.nz - 1] * sHi;
}
void compute_vert_bcs(cudaStream_t st, const Lay& L, const Coef& c, double sLo, double sHi, double* loBC, double* hiBC)
{
    vert_bcs_k<<<grid3(L.nx, L.ny, 1, B3), B3, 0, st>>>(L, c, sLo, sHi, loBC, hiBC);
    LAUNCHED();
}

// J = ((1*dxdXi_x)*dxdXi_y)*dxdXi_z and Jgup_d likewise with a divide for d
// (GeoSourceInterface.cpp:206-222, 326-350), from the per-box 1-D tables that
// LevelGeometry::createMetricCache's per-box fills amount to.  Tables are indexed from the
// box's small end (cell tables n entries, node tables n+1).
__global__ void metric_k(Lay L, int b0, int b1, int b2, int n0, int n1, int n2, const double* cx, const double* cy,
                         const double* cz, const double* fx, const double* fy, const double* fz, double* J, double* Jg0,
                         double* Jg1, double* Jg2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i > n0 || j > n1 || k > n2) return;
    const long long q  = L.idx(b0 + i, b1 + j, b2 + k);
    const bool      ci = i < n0, cj = j < n1, ck = k < n2;
    if (ci && cj && ck) J[q] = 1.0 * cx[i] * cy[j] * cz[k];
    if (cj && ck) Jg0[q] = 1.0 / fx[i] * cy[j] * cz[k];
    if (ci && ck) Jg1[q] = 1.0 * cx[i] / fy[j] * cz[k];
    if (ci && cj) Jg2[q] = 1.0 * cx[i] * cy[j] / fz[k];
}
void fill_metric_box(cudaStream_t st, const Lay& L, const int blo[3], const int bhi[3], const double* cx, const double* cy,
                     const double* cz, const double* fx, const double* fy, const double* fz, double* J, double* Jg0,
                     double* Jg1, double* Jg2)
{
    const int n0 = bhi[0] - blo[0] + 1, n1 = bhi[1] - blo[1] + 1, n2 = bhi[2] - blo[2] + 1;
    metric_k<<<grid3(n0 + 1, n1 + 1, n2 + 1, B3), B3, 0, st>>>(L, blo[0], blo[1], blo[2], n0, n1, n2, cx, cy, cz, fx, fy, fz, J,
                                                                Jg0, Jg1, Jg2);
    LAUNCHED();
}

// ------------------------------------------------------------------------------------------
// AMRNSLevel::sendToAdvectingVelocity / sendToCartesianVelocity (Grade5_SOMAR/AMRNSLevelFill.cpp:
// 194-280): the faces of one box in direction `dir` are multiplied (divided) by dx/dXi of the other
// directions, first mu = (dir + 1) % D, then (dir + 2) % D -- two FArrayBox::mult / divide calls in
// the reference, two roundings here.  t1 / t2: cell-centred 1-D tables over the box in those two
// directions (t2 null in 2-D).  n[]: faces handled in each direction (box-local, from blo).
// ------------------------------------------------------------------------------------------
__global__ void scale_faces_box_k(Lay L, int b0, int b1, int b2, int n0, int n1, int n2, int mu1, int mu2,
                                  const double* __restrict__ t1, const double* __restrict__ t2, double* __restrict__ vel, int divide)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= n0 || j >= n1 || k >= n2) return;
    const int       idx[3] = {i, j, k};
    const long long q      = L.idx(b0 + i, b1 + j, b2 + k);
    double          v      = vel[q];
    if (divide) { v = v / t1[idx[mu1]]; if (t2) v = v / t2[idx[mu2]]; }
    else { v = v * t1[idx[mu1]]; if (t2) v = v * t2[idx[mu2]]; }
    vel[q] = v;
}
void scale_faces_box(cudaStream_t st, const Lay& L, const int blo[3], const int n[3], int mu1, int mu2, const double* t1,
                     const double* t2, double* vel, bool divide)
{
    scale_faces_box_k<<<grid3(n[0], n[1], n[2], B3), B3, 0, st>>>(L, blo[0], blo[1], blo[2], n[0], n[1], n[2], mu1, mu2, t1, t2, vel,
                                                                 divide ? 1 : 0);
    LAUNCHED();
}

// ------------------------------------------------------------------------------------------
// Leaves of the leptic solver (Elliptic/LevelLepticSolver.cpp).  One thread per column; the
// vertical loops keep the reference's order, so the results agree operation for operation.
// ------------------------------------------------------------------------------------------
// computeVerticalExcess (:725-770): excess = hiBC - loBC (identically 0: never modified after
// bdryData.setVal(0)) and then FORT_UNMAPPEDVERTINTEGRAL with dz = -dXi_z (SubspaceF.ChF:33-58).
__global__ void vert_excess_k(Lay L, Lay F, double* __restrict__ excess, const double* __restrict__ hiBC,
                              const double* __restrict__ rhs, double dzScale)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= L.nx || j >= L.ny) return;
    const long long f = F.idx(i, j, 0);
    double          e = hiBC[f];
    for (int k = 0; k < L.nz; ++k) e = e + rhs[L.idx(i, j, k)] * dzScale;
    excess[f] = e;
}
void vert_excess(cudaStream_t st, const Lay& L, const Lay& F, double* excess, const double* hiBC, const double* rhs, double dzScale)
{
    const dim3 b(64, 4, 1);
    vert_excess_k<<<grid3(L.nx, L.ny, 1, b), b, 0, st>>>(L, F, excess, hiBC, rhs, dzScale);
    LAUNCHED();
}
// FORT_TRIDIAGPOISSONNN1DFAB (PoissonOpF.ChF:1329-1410): Neumann-Neumann Thomas solve with the
// reference's special-cased last row (a(r-1) and the plain b(r)), zero-mean solution.  sigma is
// Jg^{zz} on vertical faces (face r = lower face of cell r); x lives in phi, gam in scratch.
__global__ void tridiag_nn_k(Lay L, Lay F, double* __restrict__ phi, const double* __restrict__ rhs,
                             const double* __restrict__ upperBC, const double* __restrict__ sigma, double* __restrict__ gam,
                             double dx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= L.nx || j >= L.ny) return;
    const int       N    = L.nz;
    const double    dxsq = dx * dx;
    const long long q0   = L.idx(i, j, 0), sz = L.sz;
    double c = sigma[q0 + sz];
    double b = -c;
    double x = rhs[q0] * dxsq;
    double bet = b;
    x          = x / bet;
    double g   = c / bet;
    phi[q0] = x; gam[q0] = g;
    double aPrev = 1.2345e10;  // a(0), PoissonOpF.ChF:1356
    int    r;
    for (r = 1; r <= N - 2; ++r) {
        const long long q = q0 + r * sz;
        const double    a = sigma[q];
        c                 = sigma[q + sz];
        b                 = -(a + c);
        bet               = b - a * g;
        x                 = (rhs[q] * dxsq - a * x) / bet;
        g                 = c / bet;
        phi[q] = x; gam[q] = g;
        aPrev = a;
    }
    {
        const long long q = q0 + r * sz;  // r == N - 1
        const double    a = sigma[q];
        b                 = -a;
        const double xr   = rhs[q] * dxsq - upperBC[F.idx(i, j, 0)] * dx;
        x                 = (xr - aPrev * x) / b;
        phi[q]            = x;
    }
    double avg = x;
    for (r = N - 2; r >= 0; --r) {
        const long long q = q0 + r * sz;
        x                 = phi[q] - gam[q] * x;
        phi[q]            = x;
        avg               = avg + x;
    }
    avg = avg / (double)N;
    for (r = 0; r <= N - 1; ++r) phi[q0 + r * sz] = phi[q0 + r * sz] - avg;
}
void tridiag_nn(cudaStream_t st, const Lay& L, const Lay& F, double* phi, const double* rhs, const double* upperBC,
                const double* sigma, double* gam, double dx)
{
    const dim3 b(64, 4, 1);
    tridiag_nn_k<<<grid3(L.nx, L.ny, 1, b), b, 0, st>>>(L, F, phi, rhs, upperBC, sigma, gam, dx);
    LAUNCHED();
}
void add_vertical_extrusion(cudaStream_t st, const Lay& L, const Lay& F, double* dest, const double* flat)
{
    launch_valid(st, L, -1, [=] __device__(long long c, int i, int j, int) { dest[c] = dest[c] + flat[F.idx(i, j, 0)]; });
}

// ------------------------------------------------------------------------------------------
// Divergence of the advecting velocity (FiniteDiffF.ChF:42-115) and the face gradient
// Jg^{dd} * (phi(i) - phi(i-e_d)) / dXi_d (FiniteDiffF.ChF:123-147 + FArrayBox::mult,
// PoissonOp.cpp:1508-1534; the optional *beta of :1537-1541).
// ------------------------------------------------------------------------------------------
__global__ void div_k(Lay L, double* __restrict__ div, const double* __restrict__ u0, const double* __restrict__ u1,
                      const double* __restrict__ u2, double dxinv0, double dxinv1, double dxinv2, int dim)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= L.nx || j >= L.ny) return;
    const long long q = L.idx(i, j, k);
    if (dim == 3)
        div[q] = (u0[q + 1] - u0[q]) * dxinv0 + (u1[q + L.sy] - u1[q]) * dxinv1 + (u2[q + L.sz] - u2[q]) * dxinv2;
    else
        div[q] = (u0[q + 1] - u0[q]) * dxinv0 + (u2[q + L.sz] - u2[q]) * dxinv2;
}
void divergence(cudaStream_t st, const Lay& L, double* div, const double* u0, const double* u1, const double* u2,
                double dxinv0, double dxinv1, double dxinv2, int dim)
{
    div_k<<<grid3(L.nx, L.ny, L.nz, B3), B3, 0, st>>>(L, div, u0, u1, u2, dxinv0, dxinv1, dxinv2, dim);
    LAUNCHED();
}
__global__ void grad_k(Lay L, double* __restrict__ g, const double* __restrict__ phi, const double* __restrict__ Jgup, int dir,
                       double oneOnDx, double beta, int scaleBeta)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= L.nx + (dir == 0) || j >= L.ny + (dir == 1) || k >= L.nz + (dir == 2)) return;
    const long long q = L.idx(i, j, k);
    const long long s = dir == 0 ? 1 : (dir == 1 ? L.sy : L.sz);
    double          v = (phi[q] - phi[q - s]) * oneOnDx;
    v                 = v * Jgup[q];
    if (scaleBeta) v = v * beta;
    g[q] = v;
}
void gradient(cudaStream_t st, const Lay& L, double* g, const double* phi, const double* Jgup, int dir, double oneOnDx,
              double beta, int scaleBeta)
{
    grad_k<<<grid3(L.nx + 1, L.ny + 1, L.nz + 1, B3), B3, 0, st>>>(L, g, phi, Jgup, dir, oneOnDx, beta, scaleBeta);
    LAUNCHED();
}

// ------------------------------------------------------------------------------------------
// Reductions per reference box (LDFABOps.cpp:98-164 sums box by box; the 2-norm is
// sqrt(Sum_box (Sum x^2)/numPts_box), FArrayBox.cpp:138-141).  Two deterministic stages:
// RCH chunks per box, then one block per box adds the chunk partials in a fixed order.
// ------------------------------------------------------------------------------------------
constexpr int RCH = 64;
__device__ __forceinline__ double block_reduce(double v, int op, double* sm)
{
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_down_sync(0xffffffffu, v, o);
        v              = op == 0 ? fmax(v, w) : v + w;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) {
            const double w = __shfl_down_sync(0xffffffffu, v, o);
            v              = op == 0 ? fmax(v, w) : v + w;
        }
    }
    __syncthreads();
    return v;  // valid in thread 0
}
// Threads are arranged as bxw columns (a power of two >= 32 covering the box width, at most the
// block) by blockDim.x / bxw rows, and every thread keeps four rows in flight; the assignment of
// cells to threads is fixed, so the result is reproducible run to run.
// mlo / mhi: a box (tile-local indices, empty if mlo0 > mhi0) whose cells count as zero -- the cells
// covered by the finer AMR level in PoissonOp::AMRNormLevel (PoissonOp.cpp:1250-1276).
__global__ void reduce1_k(Lay L, BoxList bl, int op, const double* __restrict__ x, const double* __restrict__ y, double dv,
                          double* __restrict__ partial, int bxw, int mlo0, int mlo1, int mlo2, int mhi0, int mhi1, int mhi2,
                          const double* __restrict__ ytab)
{
    __shared__ double sm[32];
    const int bx = blockIdx.y, ch = blockIdx.x;
    const int lo0 = bl.lo[3 * bx], lo1 = bl.lo[3 * bx + 1], lo2 = bl.lo[3 * bx + 2];
    const int n0 = bl.hi[3 * bx] - lo0 + 1, n1 = bl.hi[3 * bx + 1] - lo1 + 1, n2 = bl.hi[3 * bx + 2] - lo2 + 1;
    const long long rows = (long long)n1 * n2;
    const int tx = threadIdx.x & (bxw - 1), ty = threadIdx.x / bxw, by = blockDim.x / bxw;
    double    a = 0.0, b = 0.0;
    auto acc = [&](double v, double w) {
        if (op == 0) a = fmax(a, fabs(v));
        else if (op == 1) a = a + fabs(v);
        else if (op == 2) a = a + v * v;
        else if (op == 3) a = a + v * w;
        else { const double s = w * dv; a = a + s * v; b = b + s; }  // IntegralF.ChF:37-58
    };
    const bool two = op >= 3;
    const bool masked = mlo0 <= mhi0;
    for (long long r0 = ch + (long long)RCH * ty; r0 < rows; r0 += (long long)RCH * by * 4) {
        for (int i = tx; i < n0; i += bxw) {
            double v[4], w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long r = r0 + (long long)RCH * by * u;
                v[u] = 0.0; w[u] = 0.0;
                if (r < rows) {
                    const int       jj = lo1 + (int)(r % n1), kk = lo2 + (int)(r / n1);
                    const long long q  = L.idx(lo0 + i, jj, kk);
                    v[u] = x[q];
                    if (two) w[u] = ytab ? ytab[kk] : y[q];
                    if (masked && lo0 + i >= mlo0 && lo0 + i <= mhi0 && jj >= mlo1 && jj <= mhi1 && kk >= mlo2 && kk <= mhi2) v[u] = 0.0;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r0 + (long long)RCH * by * u < rows) acc(v[u], w[u]);
        }
    }
    a = block_reduce(a, op, sm);
    if (op == 4) b = block_reduce(b, 1, sm);
    if (threadIdx.x == 0) {
        partial[2 * ((long long)bx * RCH + ch)]     = a;
        partial[2 * ((long long)bx * RCH + ch) + 1] = b;
    }
}
__global__ void reduce2_k(int op, const double* __restrict__ partial, double* __restrict__ out)
{
    __shared__ double sm[32];
    const int bx = blockIdx.x;
    double    a = 0.0, b = 0.0;
    if (threadIdx.x < RCH) {
        a = partial[2 * ((long long)bx * RCH + threadIdx.x)];
        b = partial[2 * ((long long)bx * RCH + threadIdx.x) + 1];
    }
    a = block_reduce(a, op, sm);
    if (op == 4) b = block_reduce(b, 1, sm);
    if (threadIdx.x == 0) {
        if (op == 4) { out[2 * bx] = a; out[2 * bx + 1] = b; }
        else out[bx] = a;
    }
}
// out[c] = sum over boxes of in[ncomp * b + c], boxes in ascending order (the order of the reference's loop over the
// DataIterator, Integral.cpp:249-266): one thread, a handful of adds -- what used to be a round trip to the host.
__global__ void sum_boxes_k(const double* __restrict__ in, int nboxes, int ncomp, double* __restrict__ out)
{
    const int c = threadIdx.x;
    if (c >= ncomp) return;
    double s = 0.0;
    for (int b = 0; b < nboxes; ++b) s += in[ncomp * b + c];
    out[c] = s;
}
void sum_boxes(cudaStream_t st, const double* in, int nboxes, int ncomp, double* out)
{
    sum_boxes_k<<<1, 32, 0, st>>>(in, nboxes, ncomp, out);
    LAUNCHED();
}
int  reduce_partial_len(int nboxes) { return 2 * RCH * nboxes; }
void reduce_boxes(cudaStream_t st, const Lay& L, const BoxList& boxes, int op, const double* x, const double* y, double dv,
                  double* partial, double* out, const Box3* mask, const double* ytab)
{
    int bxw = 32;
    while (bxw < 256 && bxw < boxes.maxnx) bxw <<= 1;
    const Box3 none{{1, 1, 1}, {0, 0, 0}};
    const Box3& m = mask ? *mask : none;
    reduce1_k<<<dim3(RCH, boxes.n), 256, 0, st>>>(L, boxes, op, x, y, dv, partial, bxw, m.lo[0], m.lo[1], m.lo[2], m.hi[0], m.hi[1],
                                                  m.hi[2], ytab);
    LAUNCHED();
    reduce2_k<<<boxes.n, 64, 0, st>>>(op, partial, out);
    LAUNCHED();
}

}  // namespace k
}  // namespace sb
