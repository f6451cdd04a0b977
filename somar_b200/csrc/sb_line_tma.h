// sb_line_tma.h -- host interface of the TMA-staged line relaxation kernel (sb_line_tma.cu).
#pragma once
#include "sb_core.h"
#include "sb_halo.h"

namespace sb {

struct LineTmaMap { alignas(64) unsigned char bytes[128]; };  // a CUtensorMap (opaque here: cuda.h stays out of the host headers)

struct LineTmaArgs {
    const double* mx;      // [MxL | MxR] over the tile (nx each)
    const double* my;      // [MyL | MyR] (ny each)
    const double* tab;     // shared matrix: the tables of vertline_fused_k; general: [MzL | MzR] (nz each)
    double*       own;     // colour being updated
    const double* gstart;  // general: g = 1 / d' at the level below chunks 1 .. NW-1, [NW-1][ny][nx] (natural i)
    double        aob, sLo, sHi;  // general: alpha / beta and the vertical BC factors (PoissonOpF.ChF:676-688)
    int           pass, region, nbMask;
    int           nbx, ntiles;    // filled by the launcher
    int*          fault;          // mapped host record written before a watchdog trap (may be null)
    // Neighbouring tiles, fused exchange (halo != null, region == 0): the tiles that touch an exchanged side (nbMask) run
    // first, their backward sweeps also store the face layer into the neighbour's ghost cells (peer memory), and the CTA
    // that completes the last tile of a side publishes the side's arrival counter (sb_halo.cu: same protocol as
    // halo_post_k).  The schedule below (filled by the launcher) enumerates: row 0, row ny-1, column 0, column nbx-1, interior.
    const HaloDev* halo;
    int           nA, nB, nC, nD, jlo, bxlo, nbxi;
};

namespace k {
int  vertline_tma_nw();                      // chunks per column (consumer warps)
// the order in which a pass with the fused exchange walks its tiles: order[3 u] = (bx, j, touched sides) of slot u; returns the
// number of tiles (host only; what sb_plan_line_tile_order exports for the CPU tests)
int  line_tma_tile_order(int nx, int ny, int nbMask, int* order, int capacity);
bool vertline_tma_fits(int nz, bool general);
void vertline_tma_make_maps(const SLay& S, const double* oth, const double* rhs, LineTmaMap* mapOth, LineTmaMap* mapRhs);
void vertline_tma_pass(cudaStream_t st, const SLay& S, const LineTmaMap& mapOth, const LineTmaMap& mapRhs, const LineTmaArgs& args,
                       bool general);
void line_gstart(cudaStream_t st, const Lay& L, const double* J, const double* mx, const double* my, const double* mz, double aob,
                 double sLo, double sHi, int CL, double* gstart, int* flag);
}  // namespace k
}  // namespace sb
