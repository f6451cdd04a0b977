// sb_line.cu -- relaxation on colour-split storage (sm_100a, fp64): vertical line relaxation
// (vertline_fused_k and its predecessors / diagnostics) and point red-black Gauss-Seidel
// (gsrb_split_k), plus the layout conversions, ghost fills and face packing that go with it.
//
// Reference: PoissonOp::vertLineGSRB_relax (Grade3_Calculus/Elliptic/PoissonOp.cpp:1927-2010),
// FORT_POISSONOP_VERTLINEGSRB_3D (Elliptic/PoissonOpF.ChF:851-1019) + LAPACK dgtsv;
// PoissonOp::gsrb_relax (:1833-1921) + FORT_POISSONOP_GSRB (PoissonOpF.ChF:420-474).
//
// Why a second layout.  A colour pass reads only cells of the other colour and writes only its
// own.  In the natural layout (x contiguous) that is stride-2 access: every 32-byte sector that
// moves is half wasted, and the measured DRAM traffic of a pass was 7.1 GB against 3.2 GB of
// bytes that are actually needed (profiles/r1_v2_summary.md).  Here the two colours
// (i + j) & 1 (global indices) of a field live in two arrays (sb_core.h: SLay); cell (i, j, k) is
// element SOX + (i >> 1) of row (j, k) of its colour's array, so every access of a pass is
// unit-stride and every sector moved is fully used.  Op::relax converts cor/res once per call
// (split_field_k / unsplit_field_k) and runs all its iterations on the split arrays.
//
// Why a chunked Thomas.  All columns of a depth share one tridiagonal matrix (see
// Op::buildLineTables), so the two sweeps are first-order linear recurrences with known
// coefficients, y_k = b_k + a_k y_{k-1} and x_k = z_k + c_k x_{k+1}.  Each of the NW warps of a
// CTA owns a chunk of CL = ceil(nz / NW) levels of the CTA's 32 columns: it runs the recurrence
// locally from a zero start and the true value is local + (prefix product of the coefficients)
// * (carry from the neighbouring chunk); prefix products are 1-D tables, |a_k|, |c_k| < 1 for
// the diagonally dominant matrices of this operator, so the correction is stable.  This makes
// the sweeps NW times shorter than a thread-per-column Thomas and keeps every warp loading.
// Same mathematics as dgtsv's no-interchange branch; rounding differs at the 1e-16 level.
#include "sb_core.h"

namespace sb {
namespace k {

void note_launch();  // sb_kernels.cu (launch counter)

// cell (i, j, k) of a colour-split field: which array, which element
__device__ __forceinline__ double* scell(const SLay& S, double* s0, double* s1, int i, int j, int k)
{
    return (S.colour(i, j) ? s1 : s0) + S.idx(i, j, k);
}

// ------------------------------------------------------------------------------------------
// natural -> split (valid cells).  One thread moves the pair (2t, 2t+1) of a row; scale (may be
// null) multiplies level k by scale[k] (the 1/(beta J_k) of the row-scaled system, applied to the
// right-hand side once instead of in every pass).
// ------------------------------------------------------------------------------------------
// shift (may be null): device pair (sum, vol); the field is reduced by sum / vol on the way -- the
// null-space removal of PoissonOp::removeKernel (PoissonOp.cpp:838-842) that would otherwise be a
// pass of its own.
__global__ void split_field_k(Lay L, SLay S, const double* __restrict__ nat, double* __restrict__ s0, double* __restrict__ s1,
                              const double* __restrict__ scale, const double* __restrict__ shift, const double* __restrict__ scaleJ,
                              double beta)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    const int i = 2 * t;
    if (i >= L.nx || j >= L.ny) return;
    const double    f = scale ? scale[k] : 1.0;
    const long long q = L.idx(i, j, k);
    double          e, o = 0.0;
    const bool      two = i + 1 < L.nx;
    if (two) { const double2 v = *reinterpret_cast<const double2*>(nat + q); e = v.x; o = v.y; }
    else e = nat[q];
    if (scale) { e = e * f; o = o * f; }
    if (scaleJ) {  // mapped grid: 1 / (beta J) differs from cell to cell
        e = e * (1.0 / (beta * scaleJ[q]));
        if (two) o = o * (1.0 / (beta * scaleJ[q + 1]));
    }
    if (shift) { const double avg = shift[0] / shift[1]; e = e - avg; o = o - avg; }
    const int       ce = S.colour(i, j);
    const long long d  = S.idx(i, j, k);
    (ce ? s1 : s0)[d] = e;
    if (two) (ce ? s0 : s1)[d] = o;  // (i+1) >> 1 == i >> 1 for even i
}
// PoissonOp::preCond(cor, res, 0) (cor = res * Dinv, PoissonOp.cpp:893-911) fused with the two
// conversions that start a relaxation: writes split(cor) and split(res * scale) from one read of
// res and Dinv.
__global__ void split_precond_k(Lay L, SLay S, const double* __restrict__ res, const double* __restrict__ Dinv,
                                const double* __restrict__ scale, double* __restrict__ c0, double* __restrict__ c1,
                                double* __restrict__ r0, double* __restrict__ r1, const double* __restrict__ scaleJ, double beta,
                                const double* __restrict__ tabD)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    const int i = 2 * t;
    if (i >= L.nx || j >= L.ny) return;
    const long long q = L.idx(i, j, k);
    double          f = scale ? scale[k] : 0.0, f2 = f;
    if (scaleJ) { f = 1.0 / (beta * scaleJ[q]); f2 = i + 1 < L.nx ? 1.0 / (beta * scaleJ[q + 1]) : 0.0; }
    double          re, ro = 0.0, de, dd = 0.0;
    const bool      two = i + 1 < L.nx;
    if (two) {
        const double2 v = *reinterpret_cast<const double2*>(res + q);
        re = v.x; ro = v.y;
        if (tabD) de = dd = tabD[k];
        else { const double2 w = *reinterpret_cast<const double2*>(Dinv + q); de = w.x; dd = w.y; }
    } else { re = res[q]; de = tabD ? tabD[k] : Dinv[q]; }
    const int       ce = S.colour(i, j);
    const long long d  = S.idx(i, j, k);
    (ce ? c1 : c0)[d] = re * de;
    (ce ? r1 : r0)[d] = re * f;
    if (two) { (ce ? c0 : c1)[d] = ro * dd; (ce ? r0 : r1)[d] = ro * f2; }
}
__global__ void unsplit_field_k(Lay L, SLay S, double* __restrict__ nat, const double* __restrict__ s0,
                                const double* __restrict__ s1)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    const int i = 2 * t;
    if (i >= L.nx || j >= L.ny) return;
    const long long q  = L.idx(i, j, k);
    const int       ce = S.colour(i, j);
    const long long d  = S.idx(i, j, k);
    const double    e  = (ce ? s1 : s0)[d];
    if (i + 1 < L.nx) {
        const double o = (ce ? s0 : s1)[d];
        *reinterpret_cast<double2*>(nat + q) = make_double2(e, o);
    } else nat[q] = e;
}
void split_field(cudaStream_t st, const Lay& L, const SLay& S, const double* nat, double* s0, double* s1, const double* scale,
                 const double* shift, const double* scaleJ, double beta)
{
    const dim3 b(64, 4, 1);
    const dim3 g(((L.nx + 1) / 2 + 63) / 64, (L.ny + 3) / 4, L.nz);
    split_field_k<<<g, b, 0, st>>>(L, S, nat, s0, s1, scale, shift, scaleJ, beta);
    note_launch();
}
void split_precond(cudaStream_t st, const Lay& L, const SLay& S, const double* res, const double* Dinv, const double* scale,
                   double* c0, double* c1, double* r0, double* r1, const double* scaleJ, double beta, const double* tabD)
{
    const dim3 b(64, 4, 1);
    const dim3 g(((L.nx + 1) / 2 + 63) / 64, (L.ny + 3) / 4, L.nz);
    split_precond_k<<<g, b, 0, st>>>(L, S, res, Dinv, scale, c0, c1, r0, r1, scaleJ, beta, tabD);
    note_launch();
}
void unsplit_field(cudaStream_t st, const Lay& L, const SLay& S, double* nat, const double* s0, const double* s1)
{
    const dim3 b(64, 4, 1);
    const dim3 g(((L.nx + 1) / 2 + 63) / 64, (L.ny + 3) / 4, L.nz);
    unsplit_field_k<<<g, b, 0, st>>>(L, S, nat, s0, s1);
    note_launch();
}

// ------------------------------------------------------------------------------------------
// Ghost fill of the x and y sides on split storage, one launch: the same Robin / periodic
// formulas as fill_ghosts_dir_k (BCToolsF.ChF:222-337).  physToo = false refreshes only the
// periodic images (what LevelData::exchange does between the colours, PoissonOp.cpp:1957-1965).
// blockIdx.y = k; threads [0, ny) do the two x sides of row j, threads [ny, ny + nx) the two y
// sides of column i.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ghost_one(const SLay& S, double* s0, double* s1, int dir, int side, int t, int k,
                                          const SideBC& bc, bool physToo)
{
    const int nn = dir == 0 ? S.nx : S.ny;
    int       g[2], p0[2], p1[2], w[2];  // (i, j) of ghost, first and second interior, periodic source
    const int a = dir, b = 1 - dir;
    g[a]  = side ? nn : -1;         g[b] = t;
    p0[a] = side ? nn - 1 : 0;      p0[b] = t;
    p1[a] = side ? nn - 2 : 1;      p1[b] = t;
    w[a]  = side ? 0 : nn - 1;      w[b] = t;
    if (sideIsBC(bc.kind)) {
        if (!physToo) return;
        const double v = sideGhost(bc, *scell(S, s0, s1, p0[0], p0[1], k), *scell(S, s0, s1, p1[0], p1[1], k));
        *scell(S, s0, s1, g[0], g[1], k) = v;
    } else if (bc.kind == SIDE_PERIODIC_SELF) {
        *scell(S, s0, s1, g[0], g[1], k) = *scell(S, s0, s1, w[0], w[1], k);
    }
}
__global__ void fill_ghosts_split_k(SLay S, double* s0, double* s1, SideBC xlo, SideBC xhi, SideBC ylo, SideBC yhi, int doY,
                                    int physToo)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (t < S.ny) {
        ghost_one(S, s0, s1, 0, 0, t, k, xlo, physToo);
        ghost_one(S, s0, s1, 0, 1, t, k, xhi, physToo);
    } else if (doY && t < S.ny + S.nx) {
        ghost_one(S, s0, s1, 1, 0, t - S.ny, k, ylo, physToo);
        ghost_one(S, s0, s1, 1, 1, t - S.ny, k, yhi, physToo);
    }
}
void fill_ghosts_split(cudaStream_t st, const SLay& S, double* s0, double* s1, const SideBC bc[3][2], int dim, bool physToo)
{
    auto idle = [&](const SideBC& b) { return b.kind == SIDE_NEIGHBOR || b.kind < 0 || (sideIsBC(b.kind) && !physToo); };
    const bool doY = dim == 3;
    if (idle(bc[0][0]) && idle(bc[0][1]) && (!doY || (idle(bc[1][0]) && idle(bc[1][1])))) return;
    const int n = S.ny + (doY ? S.nx : 0);
    fill_ghosts_split_k<<<dim3((n + 127) / 128, S.nz), 128, 0, st>>>(S, s0, s1, bc[0][0], bc[0][1], bc[1][0], bc[1][1], doY,
                                                                    physToo);
    note_launch();
}

// Vertical sides (split storage with z ghosts): same formulas as fill_ghosts_dir_k.
__global__ void fill_ghosts_split_z_k(SLay S, double* s0, double* s1, SideBC lo, SideBC hi, int physToo)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= S.nx || j >= S.ny) return;
    double* a = S.colour(i, j) ? s1 : s0;
    for (int side = 0; side < 2; ++side) {
        const SideBC& bc = side ? hi : lo;
        const int g = side ? S.nz : -1, p0 = side ? S.nz - 1 : 0, p1 = side ? S.nz - 2 : 1, wv = side ? 0 : S.nz - 1;
        if (sideIsBC(bc.kind)) {
            if (!physToo) continue;
            const double v = sideGhost(bc, a[S.idx(i, j, p0)], a[S.idx(i, j, p1)]);
            a[S.idx(i, j, g)] = v;
        } else if (bc.kind == SIDE_PERIODIC_SELF) {
            a[S.idx(i, j, g)] = a[S.idx(i, j, wv)];
        }
    }
}
void fill_ghosts_split_z(cudaStream_t st, const SLay& S, double* s0, double* s1, const SideBC& lo, const SideBC& hi, bool physToo)
{
    auto idle = [&](const SideBC& b) { return b.kind == SIDE_NEIGHBOR || b.kind < 0 || (sideIsBC(b.kind) && !physToo); };
    if (S.zg == 0 || (idle(lo) && idle(hi))) return;
    const dim3 b(64, 4, 1);
    fill_ghosts_split_z_k<<<dim3((S.nx + 63) / 64, (S.ny + 3) / 4), b, 0, st>>>(S, s0, s1, lo, hi, physToo ? 1 : 0);
    note_launch();
}

// Point red-black Gauss-Seidel on split storage.  The cells of one pass, (i + j + k + pass) even in
// global indices, lie at level k in the colour array a = (pass + k + lo2) & 1; their horizontal
// neighbours are in the other array at the same level, their vertical neighbours in the same array at
// k -+ 1.  Every access is unit-stride, and only what the pass needs moves: 20 B per grid cell per
// pass against 40 B in the natural layout (where every sector brings the other colour along).
// Same expression as gsrb_k / stencil7, so the result is bit-identical.
// Two adjacent cells of the pass per thread (16-byte accesses; the east neighbour of the first is the
// west neighbour of the second); the arrays of the level are picked once per CTA.
__global__ void gsrb_split_k(SLay S, Coef c, double* __restrict__ p0, double* __restrict__ p1, const double* __restrict__ r0,
                             const double* __restrict__ r1, const double* __restrict__ J0, const double* __restrict__ J1,
                             const double* __restrict__ D0, const double* __restrict__ D1, int pass)
{
    const int m = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    const int a  = (pass + k + S.lo2p) & 1;             // array of the cells being updated at this level
    const int i0 = (a + S.par + j) & 1;                 // colour(i, j) == a  <=>  i = 2 m + i0
    const int i  = 2 * m + i0;
    if (i >= S.nx || j >= S.ny) return;
    double* const       own = a ? p1 : p0;
    const double* const oth = a ? p0 : p1;
    const double* const rr  = a ? r1 : r0;
    const double* const JJ  = a ? J1 : J0;
    const double* const DD  = a ? D1 : D0;
    const long long q  = S.idx(i, j, k);                // element SOX + m of the row: 16-byte aligned
    const bool      two = i + 2 < S.nx;
    const double mzl = c.mzl[k], mzr = c.mzr[k], myl = c.myl[j], myr = c.myr[j], beta = c.beta;
    const double2 dn = *reinterpret_cast<const double2*>(own + q - S.sz), up = *reinterpret_cast<const double2*>(own + q + S.sz);
    const double2 so = *reinterpret_cast<const double2*>(oth + q - S.sy), no = *reinterpret_cast<const double2*>(oth + q + S.sy);
    const double2 rv = *reinterpret_cast<const double2*>(rr + q);
    // J / Dinv of the level from the [nz] tables when they are functions of the level only (the same bits, Coef::tabJ)
    const double2 Jv = c.tabJ ? make_double2(c.tabJ[k], c.tabJ[k]) : *reinterpret_cast<const double2*>(JJ + q);
    const double2 Dv = c.tabD ? make_double2(c.tabD[k], c.tabD[k]) : *reinterpret_cast<const double2*>(DD + q);
    // horizontal neighbours in the other array: elements m + i0 - 1, m + i0, m + i0 + 1
    double w0, e0, e1;
    if (i0) { const double2 t = *reinterpret_cast<const double2*>(oth + q); w0 = t.x; e0 = t.y; e1 = oth[q + 2]; }
    else { const double2 t = *reinterpret_cast<const double2*>(oth + q); w0 = oth[q - 1]; e0 = t.x; e1 = t.y; }
    double s = c.mxl[i] * w0 + c.mxr[i] * e0 + myl * so.x + myr * no.x;
    s        = s + mzl * dn.x + mzr * up.x;
    const double x0 = (rv.x - beta * Jv.x * s) * Dv.x;
    if (two) {
        double s2 = c.mxl[i + 2] * e0 + c.mxr[i + 2] * e1 + myl * so.y + myr * no.y;
        s2        = s2 + mzl * dn.y + mzr * up.y;
        const double x1 = (rv.y - beta * Jv.y * s2) * Dv.y;
        *reinterpret_cast<double2*>(own + q) = make_double2(x0, x1);
    } else own[q] = x0;
}
void gsrb_split_pass(cudaStream_t st, const SLay& S, const Coef& c, double* const p[2], const double* const r[2],
                     const double* const J[2], const double* const Di[2], int pass)
{
    const dim3 b(64, 4, 1);
    const int  pairs = ((S.nx + 1) / 2 + 1) / 2;
    gsrb_split_k<<<dim3((pairs + 63) / 64, (S.ny + 3) / 4, S.nz), b, 0, st>>>(S, c, p[0], p[1], r[0], r[1], J[0], J[1], Di[0], Di[1],
                                                                            pass);
    note_launch();
}

// Face layer pack / unpack for the neighbour exchange on split storage; buffer order as
// pack_face_k: (tangential index, k), tangential = j for dir 0, i for dir 1.
struct FaceBufs { double* b[2][2]; };  // [dir][side], null = side not exchanged
__global__ void pack_face_split_k(SLay S, double* s0, double* s1, FaceBufs fb, int unpack)
{
    const int dir = blockIdx.z >> 1, side = blockIdx.z & 1;
    double*   buf = fb.b[dir][side];
    if (!buf) return;
    const int t  = blockIdx.x * blockDim.x + threadIdx.x;
    const int k  = blockIdx.y;
    const int nt = dir == 0 ? S.ny : S.nx, nn = dir == 0 ? S.nx : S.ny;
    if (t >= nt) return;
    const int layer = unpack ? (side ? nn : -1) : (side ? nn - 1 : 0);
    double* c = dir == 0 ? scell(S, s0, s1, layer, t, k) : scell(S, s0, s1, t, layer, k);
    const long long m = t + (long long)nt * k;
    if (unpack) *c = buf[m];
    else buf[m] = *c;
}
// all exchanged sides in one launch; bufs[dir][side] (x and y only), null where there is no neighbour
void pack_faces_split(cudaStream_t st, const SLay& S, double* s0, double* s1, double* const bufs[2][2], bool unpack)
{
    FaceBufs fb;
    bool     any = false;
    for (int d = 0; d < 2; ++d)
        for (int sd = 0; sd < 2; ++sd) { fb.b[d][sd] = bufs[d][sd]; any = any || bufs[d][sd]; }
    if (!any) return;
    const int nt = S.nx > S.ny ? S.nx : S.ny;
    pack_face_split_k<<<dim3((nt + 127) / 128, S.nz, 4), 128, 0, st>>>(S, s0, s1, fb, unpack ? 1 : 0);
    note_launch();
}

// ------------------------------------------------------------------------------------------
// The relaxation pass.  own / oth: the colour being updated and the other one; rhs: this colour's
// right-hand side, already multiplied by s_k = 1 / (beta J_k).  Two kernels:
//
//  * vertline_fused_k (default).  In terms of z_k = g_k y_k the Thomas sweeps are
//        z_k = g_k b_k + a'_k z_{k-1},   x_k = z_k + c_k x_{k+1},   a'_k = -MzL_k g_k, c_k = -MzR_k g_k, g_k = 1 / d'_k.
//    A warp runs the forward recurrence over its chunk from zero (zl) and, in the same loop,
//    accumulates acc = sum_k R_k zl_k with R_k = prod_{m = chunk start}^{k-1} c_m: the local part of
//    the chunk's first unknown.  The true values are z_k = zl_k + P'_k Z and
//    x_first = acc + Z T + Rend X, where Z / X are the carries entering the chunk from below /
//    above and P'_k = prod a', T = sum_k R_k P'_k, Rend = prod c are tables.  So after ONE barrier
//    every warp resolves both carry chains from the NW chunk summaries and its backward sweep
//    produces final values that go straight to HBM: two passes over shared memory per level
//    instead of three, two barriers instead of three.
//    tab = {a', g}[N] | {P', c}[N] | R[N] | Pend[NW] | T[NW] | Rend[NW].
//  * vertline_split_k (SB_LINE_VARIANT=1): the earlier three-phase form (local forward, local
//    backward, correction), tab = [6][N]: s, a, P, g, c, Q.
// ------------------------------------------------------------------------------------------
// Launch shape.  The shipped library has one: 8 warps, 4 levels per load batch, 2 CTAs per SM (measured best on B200,
// profiles/r1_v4_summary.md).  Building with -DSB_DIAG_KERNELS adds the development knob SB_LINE_VARIANT (read once) and
// the diagnostic / rejected kernels it selects (vertline_split_k, vertline_pers_k, vertline_probe_k).
struct LineVariant { int nw, u, minb, fused, pers; };
static LineVariant line_variant()
{
#ifdef SB_DIAG_KERNELS
    static int v = -1;
    if (v < 0) { const char* e = getenv("SB_LINE_VARIANT"); v = e ? atoi(e) : 0; }
    switch (v) {
        case 1: return {8, 4, 2, 0, 0};
        case 2: return {8, 2, 2, 1, 0};
        case 3: return {8, 8, 1, 1, 0};
        case 4: return {16, 4, 1, 1, 0};
        case 5: return {8, 4, 1, 1, 0};
        case 7: return {8, 4, 2, 2, 0};
        case 8: return {8, 4, 2, 1, 1};   // persistent CTAs with cross-tile prefetch: measured slower (0.776 ms, see header of vertline_pers_k)
        default: break;
    }
#endif
    return {8, 4, 2, 1, 0};
}
int  vertline_split_chunk(int nz) { const int nw = line_variant().nw; return (nz + nw - 1) / nw; }
int  vertline_split_nw() { return line_variant().nw; }
bool vertline_split_fused() { return line_variant().fused != 0; }
size_t vertline_split_smem(int nz)
{
    const LineVariant v = line_variant();
    return ((size_t)nz * 32 + 5 * (size_t)nz + 2 * (size_t)v.nw * 32 + 3 * (size_t)v.nw) * sizeof(double);
}
bool vertline_split_fits(int nz) { return vertline_split_smem(nz) <= 108 * 1024; }

// ALIGNED: nz = NW * CL and CL a multiple of 2 U (true for the power-of-two depths of the bench
// grid): no clamps or per-level predicates, 32-bit element offsets (one IMAD.WIDE per load).
template <int NW, int U, int MINB, bool ALIGNED>
__global__ void __launch_bounds__(NW * 32, MINB)
    vertline_fused_k(SLay S, const double* __restrict__ mx, const double* __restrict__ my, const double* __restrict__ tab,
                     double* __restrict__ own, const double* __restrict__ oth, const double* __restrict__ rhs, int pass, int CL,
                     int region, int nbMask)
{
    // region 1 / 2: only the CTAs that do / do not own cells of a face layer that is sent to a
    // neighbouring tile (nbMask bits: x-lo, x-hi, y-lo, y-hi), so that the exchange of those layers
    // can overlap the rest of the pass (Op::relaxLineSplit).
    if (region != 0) {
        const bool edge = ((nbMask & 1) && blockIdx.x == 0) || ((nbMask & 2) && blockIdx.x == gridDim.x - 1) ||
                          ((nbMask & 4) && blockIdx.y == 0) || ((nbMask & 8) && blockIdx.y == gridDim.y - 1);
        if (edge != (region == 1)) return;
    }
    extern __shared__ double sm[];
    const int      N  = S.nz;
    double* const  sy = sm;                         // [N][32] local forward sweep
    double* const  cz = sm + (size_t)N * 32;        // [NW][32] its value at the chunk end
    double* const  ca = cz + NW * 32;               // [NW][32] acc of the chunk
    double* const  ts = ca + NW * 32;               // tables
    const double2* const T1 = reinterpret_cast<const double2*>(ts);          // {a', g}
    const double2* const T2 = reinterpret_cast<const double2*>(ts + 2 * N);  // {P', c}
    const double*  const tR = ts + 4 * N;
    const double*  const tPend = ts + 5 * N;
    const double*  const tT    = tPend + NW;
    const double*  const tRend = tT + NW;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int j    = blockIdx.y;
    const int i0   = (pass + S.par + j) & 1;   // own cells of this row: i = 2 m + i0
    const int m    = blockIdx.x * 32 + lane;
    const bool act = 2 * m + i0 < S.nx;
    const int  mm  = act ? m : 0;              // idle lanes shadow column 0 and never store
    const int  ii  = 2 * mm + i0;
    const double mxl = mx[ii], mxr = mx[S.nx + ii], myl = my[j], myr = my[S.ny + j];
    for (int k = threadIdx.x; k < 5 * N + 3 * NW; k += NW * 32) ts[k] = tab[k];
    const long long base = (long long)(SOX + mm) + S.sy * (long long)(1 + j);
    const double*   pw = oth + base + (i0 - 1);  // west neighbour (east = pw[1])
    const double*   ps = oth + base - S.sy;      // south
    const double*   pn = oth + base + S.sy;      // north
    const double*   pr = rhs + base;
    const int       szi = (int)S.sz;             // one colour array has fewer than 2^31 elements (checked by the launcher)

    const int k0 = w * CL, k1 = ALIGNED ? k0 + CL : min(N, k0 + CL);
    double    a0[U][5], a1[U][5];
    // running pointers of the four load streams (levels are issued in ascending order)
    const long long szl = S.sz;
    const double *qw = pw + (long long)k0 * szl, *qs = ps + (long long)k0 * szl, *qn = pn + (long long)k0 * szl,
                 *qr = pr + (long long)k0 * szl;
    auto issue = [&](double(&a)[U][5], int kk) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (ALIGNED) {
                a[u][0] = qw[0]; a[u][1] = qw[1]; a[u][2] = qs[0]; a[u][3] = qn[0]; a[u][4] = qr[0];
                qw += szl; qs += szl; qn += szl; qr += szl;
            } else {
                const int o = min(kk + u, N - 1) * szi;
                a[u][0] = pw[o]; a[u][1] = pw[o + 1]; a[u][2] = ps[o]; a[u][3] = pn[o]; a[u][4] = pr[o];
            }
        }
    };
    if (k0 < k1) issue(a0, k0);
    __syncthreads();  // tables visible

    // P1: right-hand sides, local forward sweep in z, and the chunk's contribution to its first unknown.
    double zl = 0.0, acc = 0.0;
    auto consume = [&](double(&a)[U][5], int kk) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = kk + u;
            if (ALIGNED || k < k1) {
                const double  lphi = fma(myr, a[u][3], fma(myl, a[u][2], fma(mxr, a[u][1], mxl * a[u][0])));
                const double  b    = a[u][4] - lphi;
                const double2 t    = T1[k];
                zl                 = fma(t.x, zl, t.y * b);
                acc                = fma(tR[k], zl, acc);
                sy[k * 32 + lane]  = zl;
            }
        }
    };
    if (ALIGNED) {
        // rolling prefetch: 2 U levels stay in flight; a slot is reloaded as soon as it is consumed
        auto roll = [&](double(&a)[U][5], int kk, bool more) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int     k    = kk + u;
                const double  lphi = fma(myr, a[u][3], fma(myl, a[u][2], fma(mxr, a[u][1], mxl * a[u][0])));
                const double  b    = a[u][4] - lphi;
                const double2 t    = T1[k];
                zl                 = fma(t.x, zl, t.y * b);
                acc                = fma(tR[k], zl, acc);
                sy[k * 32 + lane]  = zl;
                if (more) {
                    a[u][0] = qw[0]; a[u][1] = qw[1]; a[u][2] = qs[0]; a[u][3] = qn[0]; a[u][4] = qr[0];
                    qw += szl; qs += szl; qn += szl; qr += szl;
                }
            }
        };
        issue(a1, k0 + U);
        for (int kk = k0; kk < k1; kk += 2 * U) {
            const bool more = kk + 2 * U < k1;
            roll(a0, kk, more);
            roll(a1, kk + U, more);
        }
    } else {
        for (int kk = k0; kk < k1;) {
            if (kk + U < k1) issue(a1, kk + U);
            consume(a0, kk);
            kk += U;
            if (kk >= k1) break;
            if (kk + U < k1) issue(a0, kk + U);
            consume(a1, kk);
            kk += U;
        }
    }
    cz[w * 32 + lane] = zl;
    ca[w * 32 + lane] = acc;
    __syncthreads();
    if (k0 >= k1) return;

    // Carries.  Zs[v]: true z just below chunk v; X: true x just above this warp's chunk.
    const int nch = ALIGNED ? NW : (N + CL - 1) / CL;
    double    Zs[NW];
    double    Z = 0.0, Zm = 0.0;
#pragma unroll
    for (int v = 0; v < NW; ++v) {
        Zs[v] = Z;
        if (v == w) Zm = Z;
        if (v < nch) Z = fma(tPend[v], Z, cz[v * 32 + lane]);
    }
    double X = 0.0;
#pragma unroll
    for (int v = NW - 1; v >= 1; --v)
        if (v < nch && v > w) X = fma(tRend[v], X, fma(Zs[v], tT[v], ca[v * 32 + lane]));

    // P2: true backward sweep of this chunk, straight to HBM.
    double* po = own + base + (long long)(k1 - 1) * szl;
    double  xl = X;
    if (!act) return;
#pragma unroll 4
    for (int k = k1 - 1; k >= k0; --k) {
        const double2 t = T2[k];
        xl              = fma(t.y, xl, fma(t.x, Zm, sy[k * 32 + lane]));
        *po             = xl;
        po -= szl;
    }
}

#ifdef SB_DIAG_KERNELS
// Persistent form of vertline_fused_k for aligned shapes (nz = NW * CL, CL a multiple of 2 U): a
// CTA walks over tiles (32 columns of one row) with stride gridDim.x.  The tables are staged once per
// CTA instead of once per tile, and the first 2 U levels of the NEXT tile are requested before the
// barrier of the current one, so HBM requests stay in flight while the carries are resolved and the
// backward sweep streams its results out (the load registers are idle in that phase anyway).  The
// chunk summaries are double-buffered: the one barrier per tile bounds the skew between warps to
// less than a tile.  Per-cell arithmetic is that of vertline_fused_k.
// Measured on S5 (B200): 0.776 ms per pass against 0.709 ms for one CTA per tile, although the pure
// access pattern (vertline_probe_k) runs in 0.557 ms (0.620 ms at the same residency of 2 CTAs per
// SM).  Why it loses is not established: starting half of the CTAs 3-9 us late changes nothing, so
// it is not lock-step between CTAs.  Kept as SB_LINE_VARIANT=8 for the next round.
template <int NW, int U>
__global__ void __launch_bounds__(NW * 32, 2)
    vertline_pers_k(SLay S, const double* __restrict__ mx, const double* __restrict__ my, const double* __restrict__ tab,
                    double* __restrict__ own, const double* __restrict__ oth, const double* __restrict__ rhs, int pass, int CL,
                    int region, int nbMask, int nbx, int ntiles)
{
    extern __shared__ double sm[];
    const int      N  = S.nz;
    double* const  sy = sm;                         // [N][32] local forward sweep
    double* const  cs = sm + (size_t)N * 32;        // [2 sets][cz, ca][NW][32]
    double* const  ts = cs + 4 * NW * 32;           // tables
    const double2* const T1 = reinterpret_cast<const double2*>(ts);          // {a', g}
    const double2* const T2 = reinterpret_cast<const double2*>(ts + 2 * N);  // {P', c}
    const double*  const tR = ts + 4 * N;
    const double*  const tPend = ts + 5 * N;
    const double*  const tT    = tPend + NW;
    const double*  const tRend = tT + NW;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nby  = S.ny;
    auto in_region = [&](int t) -> bool {
        if (region == 0) return true;
        const int  bx = t % nbx, j = t / nbx;
        const bool edge = ((nbMask & 1) && bx == 0) || ((nbMask & 2) && bx == nbx - 1) || ((nbMask & 4) && j == 0) ||
                          ((nbMask & 8) && j == nby - 1);
        return edge == (region == 1);
    };
    int t = blockIdx.x;
    while (t < ntiles && !in_region(t)) t += gridDim.x;
    if (t >= ntiles) return;
    for (int k = threadIdx.x; k < 5 * N + 3 * NW; k += NW * 32) ts[k] = tab[k];

    const long long szl = S.sz;
    const int       k0 = w * CL, k1 = k0 + CL;
    // state of the tile whose loads are being issued
    double        mxl, mxr, myl, myr;
    bool          act;
    long long     base;
    const double *qw, *qs, *qn, *qr;
    auto setup = [&](int tile) {
        const int bx = tile % nbx, j = tile / nbx;
        const int i0 = (pass + S.par + j) & 1;   // own cells of this row: i = 2 m + i0
        const int m  = bx * 32 + lane;
        act          = 2 * m + i0 < S.nx;
        const int mm = act ? m : 0;              // idle lanes shadow column 0 and never store
        const int ii = 2 * mm + i0;
        mxl = mx[ii]; mxr = mx[S.nx + ii]; myl = my[j]; myr = my[S.ny + j];
        base = (long long)(SOX + mm) + S.sy * (long long)(1 + j);
        const long long o = base + (long long)k0 * szl;
        qw = oth + o + (i0 - 1);  // west neighbour (east = qw[1])
        qs = oth + o - S.sy;      // south
        qn = oth + o + S.sy;      // north
        qr = rhs + o;
    };
    double a0[U][5], a1[U][5];
    auto issue = [&](double(&a)[U][5]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            a[u][0] = qw[0]; a[u][1] = qw[1]; a[u][2] = qs[0]; a[u][3] = qn[0]; a[u][4] = qr[0];
            qw += szl; qs += szl; qn += szl; qr += szl;
        }
    };
    setup(t);
    issue(a0);
    issue(a1);
    __syncthreads();  // tables visible

    int set = 0;
    while (true) {
        double* const cz = cs + set * (2 * NW * 32);
        double* const ca = cz + NW * 32;
        // P1: right-hand sides, local forward sweep in z, the chunk's contribution to its first unknown.
        double zl = 0.0, acc = 0.0;
        auto roll = [&](double(&a)[U][5], int kk, bool more) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int     k    = kk + u;
                const double  lphi = fma(myr, a[u][3], fma(myl, a[u][2], fma(mxr, a[u][1], mxl * a[u][0])));
                const double  b    = a[u][4] - lphi;
                const double2 tt   = T1[k];
                zl                 = fma(tt.x, zl, tt.y * b);
                acc                = fma(tR[k], zl, acc);
                sy[k * 32 + lane]  = zl;
                if (more) {
                    a[u][0] = qw[0]; a[u][1] = qw[1]; a[u][2] = qs[0]; a[u][3] = qn[0]; a[u][4] = qr[0];
                    qw += szl; qs += szl; qn += szl; qr += szl;
                }
            }
        };
        for (int kk = k0; kk < k1; kk += 2 * U) {
            const bool more = kk + 2 * U < k1;
            roll(a0, kk, more);
            roll(a1, kk + U, more);
        }
        cz[w * 32 + lane] = zl;
        ca[w * 32 + lane] = acc;
        // what the backward sweep of THIS tile needs
        double*    po     = own + base + (long long)(k1 - 1) * szl;
        const bool actCur = act;
        // request the head of the next tile before waiting for the other warps
        int tn = t + gridDim.x;
        while (tn < ntiles && !in_region(tn)) tn += gridDim.x;
        const bool hasNext = tn < ntiles;
        if (hasNext) {
            setup(tn);
            issue(a0);
            issue(a1);
        }
        __syncthreads();

        // Carries.  Zs[v]: true z just below chunk v; X: true x just above this warp's chunk.
        double Zs[NW];
        double Z = 0.0, Zm = 0.0;
#pragma unroll
        for (int v = 0; v < NW; ++v) {
            Zs[v] = Z;
            if (v == w) Zm = Z;
            Z = fma(tPend[v], Z, cz[v * 32 + lane]);
        }
        double X = 0.0;
#pragma unroll
        for (int v = NW - 1; v >= 1; --v)
            if (v > w) X = fma(tRend[v], X, fma(Zs[v], tT[v], ca[v * 32 + lane]));

        // P2: true backward sweep of this chunk, straight to HBM.
        double xl = X;
#pragma unroll 4
        for (int k = k1 - 1; k >= k0; --k) {
            const double2 tt = T2[k];
            xl               = fma(tt.y, xl, fma(tt.x, Zm, sy[k * 32 + lane]));
            if (actCur) *po = xl;
            po -= szl;
        }
        if (!hasNext) break;
        t = tn;
        set ^= 1;
    }
}

// Diagnostic only (SB_LINE_VARIANT=7, never used by the solver): the memory access pattern of
// vertline_fused_k -- same five load streams per level, same chunking, same store -- without the
// recurrences, shared memory or barriers.  Its time is the floor the access pattern allows.
template <int NW, int U>
__global__ void __launch_bounds__(NW * 32, 2)
    vertline_probe_k(SLay S, double* __restrict__ own, const double* __restrict__ oth, const double* __restrict__ rhs, int pass, int CL)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int j    = blockIdx.y;
    const int i0   = (pass + S.par + j) & 1;
    const int m    = blockIdx.x * 32 + lane;
    const bool act = 2 * m + i0 < S.nx;
    const int  mm  = act ? m : 0;
    const long long base = (long long)(SOX + mm) + S.sy * (long long)(1 + j);
    const long long szl  = S.sz;
    const int k0 = w * CL, k1 = min(S.nz, k0 + CL);
    const double *qw = oth + base + (i0 - 1) + (long long)k0 * szl, *qs = oth + base - S.sy + (long long)k0 * szl,
                 *qn = oth + base + S.sy + (long long)k0 * szl, *qr = rhs + base + (long long)k0 * szl;
    double* po = own + base + (long long)k0 * szl;
    for (int kk = k0; kk < k1; kk += U) {
        double v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            v[u] = qw[0] + qw[1] + qs[0] + qn[0] + qr[0];
            qw += szl; qs += szl; qn += szl; qr += szl;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (act) *po = v[u];
            po += szl;
        }
    }
}

template <int NW, int U, int MINB, bool TG>
__global__ void __launch_bounds__(NW * 32, MINB)
    vertline_split_k(SLay S, const double* __restrict__ mx, const double* __restrict__ my, const double* __restrict__ tab,
                     double* __restrict__ own, const double* __restrict__ oth, const double* __restrict__ rhs, int pass, int CL,
                     int region, int nbMask)
{
    // region 1 / 2: only the CTAs that do / do not own cells of a face layer that is sent to a
    // neighbouring tile (nbMask bits: x-lo, x-hi, y-lo, y-hi), so that the exchange of those layers
    // can overlap the rest of the pass (Op::relaxLineSplit).
    if (region != 0) {
        const bool edge = ((nbMask & 1) && blockIdx.x == 0) || ((nbMask & 2) && blockIdx.x == gridDim.x - 1) ||
                          ((nbMask & 4) && blockIdx.y == 0) || ((nbMask & 8) && blockIdx.y == gridDim.y - 1);
        if (edge != (region == 1)) return;
    }
    extern __shared__ double sm[];
    const int     N  = S.nz;
    double* const sy = sm;                     // [N][32] local sweeps, in place
    double* const cy = sm + (size_t)N * 32;    // [NW][32] chunk-end values of the forward sweep
    double* const cx = cy + NW * 32;           // [NW][32] chunk-start values of the backward sweep
    double* const ts = cx + NW * 32;           // tables a, P, g, c, Q staged here unless TG
    const double* const ta = TG ? tab + N : ts;
    const double* const tP = ta + N;
    const double* const tg = tP + N;
    const double* const tc = tg + N;
    const double* const tQ = tc + N;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int j    = blockIdx.y;
    const int i0   = (pass + S.par + j) & 1;   // own cells of this row: i = 2 m + i0
    const int m    = blockIdx.x * 32 + lane;
    const bool act = 2 * m + i0 < S.nx;
    const int  mm  = act ? m : 0;
    const int  ii  = 2 * mm + i0;
    // tile-local 1-D tables: mx = [mxl | mxr] (nx each), my = [myl | myr] (ny each)
    const double mxl = mx[ii], mxr = mx[S.nx + ii], myl = my[j], myr = my[S.ny + j];
    if (!TG) for (int k = threadIdx.x; k < 5 * N; k += NW * 32) ts[k] = tab[N + k];
    const long long sz = S.sz, sy_ = S.sy;
    const long long base = (long long)(SOX + mm) + sy_ * (long long)(1 + j);
    const double*   pw = oth + base + (i0 - 1);  // west neighbour (east = pw[1])
    const double*   pc = oth + base;             // south = pc[-sy_], north = pc[+sy_]
    const double*   pr = rhs + base;

    const int k0 = w * CL, k1 = min(N, k0 + CL);
    double    a0[U][5], a1[U][5];
    auto issue = [&](double(&a)[U][5], int kk) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long o = (long long)min(kk + u, N - 1) * sz;
            a[u][0] = pw[o]; a[u][1] = pw[o + 1]; a[u][2] = pc[o - sy_]; a[u][3] = pc[o + sy_]; a[u][4] = pr[o];
        }
    };
    if (k0 < k1) issue(a0, k0);
    __syncthreads();  // tables visible

    // P1: right-hand sides and the local forward sweep of this warp's chunk.
    double yl = 0.0;
    auto consume = [&](double(&a)[U][5], int kk) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = kk + u;
            if (k < k1) {
                const double lphi = mxl * a[u][0] + mxr * a[u][1] + myl * a[u][2] + myr * a[u][3];
                const double b    = act ? a[u][4] - lphi : 0.0;
                yl                = fma(ta[k], yl, b);
                sy[k * 32 + lane] = yl;
            }
        }
    };
    for (int kk = k0; kk < k1;) {
        if (kk + U < k1) issue(a1, kk + U);
        consume(a0, kk);
        kk += U;
        if (kk >= k1) break;
        if (kk + U < k1) issue(a0, kk + U);
        consume(a1, kk);
        kk += U;
    }
    cy[w * 32 + lane] = yl;
    __syncthreads();

    // P2: true forward values from the carry, scaled, and the local backward sweep.
    if (k0 < k1) {
        double Y = 0.0;
        for (int v = 0; v < w; ++v) Y = fma(tP[min(N, (v + 1) * CL) - 1], Y, cy[v * 32 + lane]);
        double xl = 0.0;
#pragma unroll 4
        for (int k = k1 - 1; k >= k0; --k) {
            const double y = fma(tP[k], Y, sy[k * 32 + lane]);
            xl             = fma(tc[k], xl, y * tg[k]);
            sy[k * 32 + lane] = xl;
        }
        cx[w * 32 + lane] = xl;
    }
    __syncthreads();

    // P3: true solution from the backward carry; store.
    if (k0 < k1) {
        const int nch = (N + CL - 1) / CL;
        double    X   = 0.0;
        for (int v = nch - 1; v > w; --v) X = fma(tQ[v * CL], X, cx[v * 32 + lane]);
        double* po = own + base;
        if (act) {
#pragma unroll 4
            for (int k = k0; k < k1; ++k) po[(long long)k * sz] = fma(tQ[k], X, sy[k * 32 + lane]);
        }
    }
}

#endif  // SB_DIAG_KERNELS

void vertline_split_pass(cudaStream_t st, const SLay& S, const Coef& c, const double* tab, double* own, const double* oth,
                         const double* rhs, int pass, int region, int nbMask)
{
    const size_t sh = vertline_split_smem(S.nz);
    const int    CL = vertline_split_chunk(S.nz);
    const dim3   g(((S.nx + 1) / 2 + 31) / 32, S.ny);
    const LineVariant v = line_variant();
#define SB_LAUNCH(KERNEL)                                                                                    \
    {                                                                                                        \
        static size_t configured = 0;                                                                        \
        if (sh > configured) {                                                                               \
            cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);              \
            configured = sh;                                                                                 \
        }                                                                                                    \
        KERNEL<<<g, v.nw * 32, sh, st>>>(S, c.mxl, c.myl, tab, own, oth, rhs, pass, CL, region, nbMask);     \
    }
    if ((long long)S.sz * S.nz >= (1LL << 31)) SB_FAIL("colour-split field too large for 32-bit element offsets");
    const bool al = S.nz == v.nw * CL && CL % (2 * v.u) == 0;
#ifdef SB_DIAG_KERNELS
    if (v.fused == 1 && v.pers && al && v.nw == 8 && v.u == 4) {
        static int    nsm = 0;
        static size_t configured = 0;
        if (!nsm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev); }
        const size_t shp = sh + 2 * (size_t)v.nw * 32 * sizeof(double);  // second set of chunk summaries
        if (shp > configured) {
            cudaFuncSetAttribute(vertline_pers_k<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shp);
            configured = shp;
        }
        const int nbx = (int)g.x, ntiles = (int)(g.x * g.y);
        const int grid = ntiles < 2 * nsm ? ntiles : 2 * nsm;
        vertline_pers_k<8, 4><<<grid, 256, shp, st>>>(S, c.mxl, c.myl, tab, own, oth, rhs, pass, CL, region, nbMask, nbx, ntiles);
    } else if (v.fused == 2) {  // access-pattern probe (diagnostic)
        if (S.nz % (8 * 8) != 0) SB_FAIL("probe kernel needs nz to be a multiple of 64");
        static const int psm = [] { const char* e = getenv("SB_PROBE_SMEM"); return e ? atoi(e) : 0; }();  // limits residency
        static bool      set = false;
        if (psm > 48 * 1024 && !set) { cudaFuncSetAttribute(vertline_probe_k<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm); set = true; }
        vertline_probe_k<8, 8><<<g, 256, psm, st>>>(S, own, oth, rhs, pass, CL);
    } else if (!v.fused) SB_LAUNCH((vertline_split_k<8, 4, 2, false>))
    else if (v.nw == 8 && v.u == 2 && al) SB_LAUNCH((vertline_fused_k<8, 2, 2, true>))
    else if (v.nw == 8 && v.u == 2) SB_LAUNCH((vertline_fused_k<8, 2, 2, false>))
    else if (v.nw == 8 && v.u == 8 && al) SB_LAUNCH((vertline_fused_k<8, 8, 1, true>))
    else if (v.nw == 16 && v.u == 4 && al) SB_LAUNCH((vertline_fused_k<16, 4, 1, true>))
    else if (v.nw == 16 && v.u == 4) SB_LAUNCH((vertline_fused_k<16, 4, 1, false>))
    else if (v.nw == 8 && v.u == 4 && v.minb == 1 && al) SB_LAUNCH((vertline_fused_k<8, 4, 1, true>))
    else
#endif
    if (al && v.nw == 8) SB_LAUNCH((vertline_fused_k<8, 4, 2, true>))
    else if (v.nw == 8) SB_LAUNCH((vertline_fused_k<8, 4, 2, false>))
    else SB_FAIL("SB_LINE_VARIANT: no kernel instance for this shape");
#undef SB_LAUNCH
    note_launch();
}

}  // namespace k
}  // namespace sb
