// sb_amr_kernels.cu -- fp64 CUDA kernels (sm_100a) of the coarse-fine coupling between two AMR levels of the
// pressure projection: quadratic coarse-fine ghost interpolation, the two sides of the flux register,
// and region copies between the arrays of different levels.
//
// Reference (paths relative to /root/reference/src):
//   Grade2_AnisotropicChombo/QuadCFInterp/MappedQuadCFInterp.cpp:271-505 (getPhiStar), :513-573 (interpOnIVS),
//   MappedQuadCFInterpF.ChF:9-46 (MAPPEDQUADINTERP), :50-127 (MAPPEDPHISTAR),
//   MappedCFStencil.cpp:379-598 (derivative evaluation; the stencil SELECTION is done on the host, sb_amr_plan.cpp),
//   Grade3_Calculus/Elliptic/PoissonOp.cpp:1295-1418 (getFlux, reflux),
//   Grade2_AnisotropicChombo/AnisotropicFluxRegister.cpp:329-384 (incrementCoarse), :411-534 (incrementFine),
//   :632-652 (reflux), AnisotropicFluxRegisterF.ChF (ANISOTROPICINCREMENTFINE).
// Compiled with -fmad=false like the other kernels: sums run in the reference's order.
#include "sb_core.h"
#include "sb_amr.h"

namespace sb {
namespace k {

void note_launch();  // sb_kernels.cu (launch counter)

// ------------------------------------------------------------------------------------------
// dst(box) = src(box) or dst(box) += scale * src(box); box in global indices, both arrays
// addressed through their own layouts (valid cells or ghosts).
// ------------------------------------------------------------------------------------------
__global__ void copy_region_k(Lay Ls, const double* __restrict__ src, Lay Ld, double* __restrict__ dst, int lo0, int lo1, int lo2,
                              int n0, int n1, int n2, int mode, double scale)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= n0 || j >= n1 || k >= n2) return;
    const long long qs = Ls.idx(lo0 + i - Ls.lo0, lo1 + j - Ls.lo1, lo2 + k - Ls.lo2);
    const long long qd = Ld.idx(lo0 + i - Ld.lo0, lo1 + j - Ld.lo1, lo2 + k - Ld.lo2);
    if (mode == 0) dst[qd] = src[qs];
    else dst[qd] = dst[qd] + src[qs] * scale;  // AddOp::linearIn: argR += *buffer * scale
}
void copy_region(cudaStream_t st, const Lay& Ls, const double* src, const Lay& Ld, double* dst, const Box3& b, int mode, double scale)
{
    const dim3 blk(32, 4, 1);
    const dim3 g((b.size(0) + 31) / 32, (b.size(1) + 3) / 4, b.size(2));
    copy_region_k<<<g, blk, 0, st>>>(Ls, src, Ld, dst, b.lo[0], b.lo[1], b.lo[2], b.size(0), b.size(1), b.size(2), mode, scale);
    note_launch();
}
// the same through a dense staging buffer (x fastest) for regions that travel between ranks
__global__ void stage_region_k(Lay L, double* __restrict__ field, double* __restrict__ buf, int lo0, int lo1, int lo2, int n0, int n1,
                               int n2, int mode, double scale)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= n0 || j >= n1 || k >= n2) return;
    const long long q = L.idx(lo0 + i - L.lo0, lo1 + j - L.lo1, lo2 + k - L.lo2);
    const long long m = i + (long long)n0 * (j + (long long)n1 * k);
    if (mode < 0) buf[m] = field[q];                      // pack
    else if (mode == 0) field[q] = buf[m];                // unpack, assign
    else field[q] = field[q] + buf[m] * scale;            // unpack, add
}
void stage_region(cudaStream_t st, const Lay& L, double* field, double* buf, const Box3& b, int mode, double scale)
{
    const dim3 blk(32, 4, 1);
    const dim3 g((b.size(0) + 31) / 32, (b.size(1) + 3) / 4, b.size(2));
    stage_region_k<<<g, blk, 0, st>>>(L, field, buf, b.lo[0], b.lo[1], b.lo[2], b.size(0), b.size(1), b.size(2), mode, scale);
    note_launch();
}

// ------------------------------------------------------------------------------------------
// Quadratic coarse-fine ghost interpolation of one side of the fine tile.  One thread per fine ghost
// cell.  The coarse data sit in the buffer array (coarsened tile grown by 2, layout P.B); the record
// of the coarse cell under the ghost holds the weights of its tangential first / second / mixed
// derivative stencils (centred, one-sided, dropped -- resolved on the host).  Then
//   phistar = phic + sum_t (slope_t x_t + curv_t x_t^2 / 2) + mixed x_0 x_1          (MAPPEDPHISTAR)
// with x_t the offset of the fine cell centre from the coarse cell centre, and the ghost is the
// parabola through the two interior fine cells and phistar                           (MAPPEDQUADINTERP).
// ------------------------------------------------------------------------------------------
__global__ void cf_interp_k(CFSideParams P, Lay Lf, double* __restrict__ fine, const double* __restrict__ buf)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;  // fine index along t0 (tile-local)
    const int b = blockIdx.y * blockDim.y + threadIdx.y;  // fine index along t1 (tile-local; 0 in 2-D)
    if (a >= P.nf0 || b >= P.nf1) return;
    const int ca = a / P.r0, cb = b / P.r1;  // coarse cell (tile lows are multiples of the ratio)
    const int rec = ca + P.nc0 * cb;
    const long long st[3] = {1, P.B.sy, P.B.sz};
    // coarse cell under the ghost, buffer-local
    int c[3];
    c[P.dir] = P.cn - P.blo[P.dir];
    c[P.t0]  = P.clo[P.t0] + ca - P.blo[P.t0];
    if (P.t1 >= 0) c[P.t1] = P.clo[P.t1] + cb - P.blo[P.t1];
    else c[3 - P.dir - P.t0] = 0;
    const long long qc = P.B.idx(c[0], c[1], c[2]);
    const double    pc = buf[qc];
    const long long s0 = st[P.t0], s1 = P.t1 >= 0 ? st[P.t1] : 0;

    const double* w1 = P.w1 + 10 * (long long)rec;
    const double* w2 = P.w2 + 10 * (long long)rec;
    double slope0 = 0.0, curv0 = 0.0, slope1 = 0.0, curv1 = 0.0, mixed = 0.0;
#pragma unroll
    for (int o = 0; o < 5; ++o) {
        const double wa = w1[o], wb = w2[o];
        if (wa != 0.0 || wb != 0.0) {
            const double v = buf[qc + (o - 2) * s0];
            if (wa != 0.0) slope0 = slope0 + wa * v;
            if (wb != 0.0) curv0 = curv0 + wb * v;
        }
    }
    slope0 = slope0 / P.dxc[P.t0];
    curv0  = curv0 / (P.dxc[P.t0] * P.dxc[P.t0]);
    const double x0 = (P.flo[P.t0] + a + 0.5) * P.dxf[P.t0] - (P.clo[P.t0] + ca + 0.5) * P.dxc[P.t0];
    double       ps;
    if (P.t1 >= 0) {
#pragma unroll
        for (int o = 0; o < 5; ++o) {
            const double wa = w1[5 + o], wb = w2[5 + o];
            if (wa != 0.0 || wb != 0.0) {
                const double v = buf[qc + (o - 2) * s1];
                if (wa != 0.0) slope1 = slope1 + wa * v;
                if (wb != 0.0) curv1 = curv1 + wb * v;
            }
        }
        slope1 = slope1 / P.dxc[P.t1];
        curv1  = curv1 / (P.dxc[P.t1] * P.dxc[P.t1]);
        const double* wm = P.wm + 9 * (long long)rec;
#pragma unroll
        for (int o1 = 0; o1 < 3; ++o1)
#pragma unroll
            for (int o0 = 0; o0 < 3; ++o0) {
                const double w = wm[3 * o1 + o0];
                if (w != 0.0) mixed = mixed + w * buf[qc + (o0 - 1) * s0 + (o1 - 1) * s1];
            }
        mixed = mixed / (P.dxc[P.t1] * P.dxc[P.t0]);
        const double x1 = (P.flo[P.t1] + b + 0.5) * P.dxf[P.t1] - (P.clo[P.t1] + cb + 0.5) * P.dxc[P.t1];
        ps = pc + (slope0 * x0 + curv0 * x0 * x0 * 0.5) + (slope1 * x1 + curv1 * x1 * x1 * 0.5) + mixed * x0 * x1;
    } else {
        // CH_SPACEDIM == 2: MappedQuadCFInterp.cpp:455-480 (the C++ branch; MAPPEDPHISTAR is 3-D only)
        const double update1 = x0 * slope0 + 0.5 * x0 * x0 * curv0;
        ps = pc + update1 + 0.0 + 0.0;
    }

    // the ghost cell and the two fine cells inside of it along the normal
    int f[3];
    f[P.dir] = P.side ? P.nfn : -1;
    f[P.t0]  = a;
    if (P.t1 >= 0) f[P.t1] = b;
    else f[3 - P.dir - P.t0] = 0;
    const long long sf[3] = {1, Lf.sy, Lf.sz};
    const long long qg    = Lf.idx(f[0], f[1], f[2]);
    const long long in    = P.side ? -sf[P.dir] : sf[P.dir];
    const double    pb = fine[qg + in], pa = fine[qg + 2 * in];
    const double    h = P.dxf[P.dir];
    const double    nref = (double)P.rn;
    const double    mult = (2.0 / (h * h)) / (nref * nref + 4.0 * nref + 3.0);
    const double    aa = mult * (2.0 * ps + (nref + 1.0) * pa - (nref + 3.0) * pb);
    const double    bb = (pb - pa) * (1.0 / h) - aa * h;
    fine[qg]           = (4.0 * h * h) * aa + bb * (2.0 * h) + pa;
}
void cf_interp(cudaStream_t st, const CFSideParams& P, const Lay& Lf, double* fine, const double* buf)
{
    const dim3 blk(32, 4, 1);
    const dim3 g((P.nf0 + 31) / 32, (P.nf1 + 3) / 4, 1);
    cf_interp_k<<<g, blk, 0, st>>>(P, Lf, fine, buf);
    note_launch();
}

// ------------------------------------------------------------------------------------------
// Coarse side of the flux register, applied to the coarse cells just outside side (dir, side) of the
// refined patch: res -= (-sign / dXi_dir) * F_c with F_c = beta Jg^{dd} (phi_hi - phi_lo) / dXi_dir at
// the shared face (getFlux + incrementCoarse + the first half of AnisotropicFluxRegister::reflux),
// or with F_c read from a given flux field (PoissonOp::reflux(div, flux, fineFlux)).
// slab: the cells in tile-local indices of the coarse layout.
// ------------------------------------------------------------------------------------------
__global__ void reflux_coarse_k(Lay L, double* __restrict__ res, const double* __restrict__ phi, const double* __restrict__ Jgup,
                                const double* __restrict__ flux, int dir, int side, int lo0, int lo1, int lo2, int n0, int n1, int n2,
                                double oneOnDx, double beta)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= n0 || j >= n1 || k >= n2) return;
    const long long q  = L.idx(lo0 + i, lo1 + j, lo2 + k);
    const long long s  = dir == 0 ? 1 : (dir == 1 ? L.sy : L.sz);
    const long long qf = side ? q : q + s;  // the face towards the patch: low face of this cell (hi side) or of the next one
    double F;
    if (flux) F = flux[qf];
    else {
        F = (phi[qf] - phi[qf - s]) * oneOnDx;  // FINITEDIFF_PARTIALD_CC2NC
        F = F * Jgup[qf];
        F = F * beta;
    }
    const double sign  = side ? 1.0 : -1.0;
    const double scale = -sign * oneOnDx;
    const double reg   = scale * F;
    res[q]             = res[q] + (-1.0) * reg;
}
void reflux_coarse(cudaStream_t st, const Lay& L, double* res, const double* phi, const double* Jgup, const double* flux, int dir,
                   int side, const int lo[3], const int n[3], double oneOnDx, double beta)
{
    const dim3 blk(32, 4, 1);
    const dim3 g((n[0] + 31) / 32, (n[1] + 3) / 4, n[2]);
    reflux_coarse_k<<<g, blk, 0, st>>>(L, res, phi, Jgup, flux, dir, side, lo[0], lo[1], lo[2], n[0], n[1], n[2], oneOnDx, beta);
    note_launch();
}

// Fine side of the flux register for side (dir, side) of the fine tile: one thread per coarse cell
// outside the coarsened tile; its fine faces are added in the Fortran loop order of
// ANISOTROPICINCREMENTFINE (x fastest) with scale = sign / dXi_dir(coarse) / (prod(ref) / ref_dir).
// The sum lands in the ghost face of the coarsened-tile array `reg` (layout Lc).
__global__ void fine_register_k(CFSideParams P, Lay Lf, Lay Lc, double* __restrict__ reg, const double* __restrict__ phi,
                                const double* __restrict__ Jgup, const double* __restrict__ flux, double oneOnDxf, double beta,
                                double scale)
{
    const int ca = blockIdx.x * blockDim.x + threadIdx.x;
    const int cb = blockIdx.y * blockDim.y + threadIdx.y;
    if (ca >= P.nc0 || cb >= P.nc1) return;
    const long long sf[3] = {1, Lf.sy, Lf.sz};
    const long long sn    = sf[P.dir];
    double          acc   = 0.0;
    for (int b1 = 0; b1 < P.r1; ++b1)
        for (int b0 = 0; b0 < P.r0; ++b0) {
            int f[3];
            f[P.dir] = P.side ? P.nfn : 0;  // face index = low face of this cell
            f[P.t0]  = ca * P.r0 + b0;
            if (P.t1 >= 0) f[P.t1] = cb * P.r1 + b1;
            else f[3 - P.dir - P.t0] = 0;
            const long long qf = Lf.idx(f[0], f[1], f[2]);
            double F;
            if (flux) F = flux[qf];
            else {
                F = (phi[qf] - phi[qf - sn]) * oneOnDxf;
                F = F * Jgup[qf];
                F = F * beta;
            }
            acc = acc + scale * F;
        }
    int c[3];
    c[P.dir] = P.side ? P.ncn : -1;
    c[P.t0]  = ca;
    if (P.t1 >= 0) c[P.t1] = cb;
    else c[3 - P.dir - P.t0] = 0;
    reg[Lc.idx(c[0], c[1], c[2])] = acc;
}
void fine_register(cudaStream_t st, const CFSideParams& P, const Lay& Lf, const Lay& Lc, double* reg, const double* phi,
                   const double* Jgup, const double* flux, double oneOnDxf, double beta, double scale)
{
    const dim3 blk(32, 4, 1);
    const dim3 g((P.nc0 + 31) / 32, (P.nc1 + 3) / 4, 1);
    fine_register_k<<<g, blk, 0, st>>>(P, Lf, Lc, reg, phi, Jgup, flux, oneOnDxf, beta, scale);
    note_launch();
}

}  // namespace k
}  // namespace sb
