// sb_amr.h -- types shared by the AMR host code (sb_amr.cpp) and its kernels (sb_amr_kernels.cu).
#pragma once
#include "sb_core.h"

namespace sb {

// One coarse-fine side of a refined level's tile, as the kernels see it.
struct CFSideParams {
    int dir, side;     // normal direction; 0 low / 1 high
    int t0, t1;        // tangential directions, ascending; t1 = -1 in a 2-D build
    int nf0, nf1, nfn; // fine cells of the tile along t0, t1 and the normal
    int nc0, nc1, ncn; // cells of the coarsened tile along t0, t1 and the normal
    int r0, r1, rn;    // refinement ratio along t0, t1 and the normal
    int flo[3];        // global fine index of the tile's first cell
    int clo[3];        // global coarse index of the coarsened tile's first cell
    int blo[3];        // global coarse index of the buffer's first cell
    int cn;            // global coarse index (normal direction) of the coarse cells under the ghosts
    double dxf[3], dxc[3];
    Lay B;             // layout of the coarse buffer
    const double* w1;  // [nc1 * nc0][2][5] first-derivative weights   (sb_amr_plan.cpp)
    const double* w2;  // [nc1 * nc0][2][5] second-derivative weights
    const double* wm;  // [nc1 * nc0][3][3] mixed-derivative weights
};

namespace k {
void copy_region(cudaStream_t st, const Lay& Ls, const double* src, const Lay& Ld, double* dst, const Box3& b, int mode, double scale);
// mode -1: buf = field(b); 0: field(b) = buf; 1: field(b) += scale * buf
void stage_region(cudaStream_t st, const Lay& L, double* field, double* buf, const Box3& b, int mode, double scale);
void cf_interp(cudaStream_t st, const CFSideParams& P, const Lay& Lf, double* fine, const double* buf);
void reflux_coarse(cudaStream_t st, const Lay& L, double* res, const double* phi, const double* Jgup, const double* flux, int dir,
                   int side, const int lo[3], const int n[3], double oneOnDx, double beta);
void fine_register(cudaStream_t st, const CFSideParams& P, const Lay& Lf, const Lay& Lc, double* reg, const double* phi,
                   const double* Jgup, const double* flux, double oneOnDxf, double beta, double scale);
}  // namespace k

}  // namespace sb
