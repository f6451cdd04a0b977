// sb_core.h -- internal types shared by the host driver (sb_op.cpp, sb_mg.cpp, sb_capi.cpp)
// and the CUDA kernels (sb_kernels.cu).  Nothing here crosses the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/somar_b200.h"

namespace sb {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
#define SB_FAIL(msg) throw ::sb::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg))
#define SB_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) SB_FAIL(std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

struct Box3 {
    int lo[3], hi[3];
    int size(int d) const { return hi[d] - lo[d] + 1; }
    long long numPts() const { return (long long)size(0) * size(1) * size(2); }
    bool operator==(const Box3& o) const
    {
        for (int d = 0; d < 3; ++d)
            if (lo[d] != o.lo[d] || hi[d] != o.hi[d]) return false;
        return true;
    }
};

// floor division (Chombo's coarsen for negative indices, BoxTools/IntVect.H coarsen)
inline int fdiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline Box3 coarsen(const Box3& b, const int r[3])
{
    Box3 c;
    for (int d = 0; d < 3; ++d) { c.lo[d] = fdiv(b.lo[d], r[d]); c.hi[d] = fdiv(b.hi[d], r[d]); }
    return c;
}
inline Box3 refine(const Box3& b, const int r[3])
{
    Box3 c;
    for (int d = 0; d < 3; ++d) { c.lo[d] = b.lo[d] * r[d]; c.hi[d] = (b.hi[d] + 1) * r[d] - 1; }
    return c;
}
inline bool coarsenable(const Box3& b, const int r[3]) { return refine(coarsen(b, r), r) == b; }

// Data layout of every field of one rank at one MG depth ("fused tile"): the union of the
// rank's boxes is one rectangle [tile.lo, tile.hi]; a field is a pitched 3-D array with one
// ghost layer in y and z, OX pad cells (>= 1 ghost) left of x and >= 1 right.  The same shape
// serves cell- and face-centred data (a face array needs n+1 entries in its own direction).
constexpr int OX = 4;  // first valid x sits on a 32-byte sector boundary
struct Lay {
    int nx, ny, nz;     // valid cells of the tile
    int lo0, lo1, lo2;  // global index of the first valid cell
    int px, py, pz;     // allocated extents
    long long sy, sz;   // strides (elements)
    long long n;        // total elements
    __host__ __device__ long long idx(int i, int j, int k) const  // local indices, ghosts = -1 / n
    {
        return (long long)(OX + i) + sy * (long long)(1 + j) + sz * (long long)(1 + k);
    }
};
inline Lay makeLay(const Box3& tile)
{
    Lay L;
    L.nx = tile.size(0); L.ny = tile.size(1); L.nz = tile.size(2);
    L.lo0 = tile.lo[0]; L.lo1 = tile.lo[1]; L.lo2 = tile.lo[2];
    L.px = ((L.nx + 8 + 3) / 4) * 4;
    L.py = L.ny + 2;
    L.pz = L.nz + 2;
    L.sy = L.px;
    L.sz = (long long)L.px * L.py;
    L.n  = L.sz * L.pz;
    return L;
}

// Colour-split ("checkerboard-compact") storage of a cell field, used inside the vertical line
// relaxation.  A colour pass reads only cells of the other colour and writes only its own, so in
// the natural layout every 32-byte sector moved is half wasted (stride-2 access in x).  Here the
// two colours (i + j + lo0 + lo1) & 1 of the tile live in two separate arrays; cell (i, j, k) sits
// at index SOX + floor(i / 2) of row (j, k) of its colour's array, so every access of a colour pass
// is unit-stride.  One ghost in x and y, none in z (the line solve folds the vertical BCs).
constexpr int SOX = 4;
struct SLay {
    int nx, ny, nz;
    int par;           // (lo0 + lo1) & 1
    int lo2p;          // lo2 & 1 (point red-black colours also count k)
    int zg;            // ghost layers in z: 0 for the line relaxation, 1 for point GSRB
    int px, py;        // allocated extents of one colour array
    long long sy, sz;  // strides (elements)
    long long n;       // elements of one colour array
    __host__ __device__ long long idx(int i, int j, int k) const
    {
        return (long long)(SOX + (i >> 1)) + sy * (long long)(1 + j) + sz * (long long)(k + zg);
    }
    __host__ __device__ int colour(int i, int j) const { return (par + i + j) & 1; }
};
inline SLay makeSLay(const Lay& L, int zg = 0)
{
    SLay S;
    S.nx = L.nx; S.ny = L.ny; S.nz = L.nz;
    S.par = (L.lo0 + L.lo1) & 1;
    S.lo2p = L.lo2 & 1;
    S.zg = zg;
    S.px = ((SOX + (L.nx + 1) / 2 + 2 + 3) / 4) * 4;
    S.py = L.ny + 2;
    S.sy = S.px;
    S.sz = (long long)S.px * S.py;
    S.n  = S.sz * (L.nz + 2 * zg);
    return S;
}

// What a side of the tile touches.
// SIDE_CF: the side of a refined AMR level's patch that borders the coarser level.  Its ghosts are
// either interpolated from coarse data (inhomogeneous, sb_amr_kernels.cu: cf_interp_k) or, in every
// relaxation / residual of a level solve, filled by the homogeneous formula of
// CFInterp::homogInterpAtCFI (Grade3_Calculus/CFInterp.cpp:429-515, CFInterpF.ChF:305-370).
enum SideKind { SIDE_PHYS = 0, SIDE_PERIODIC_SELF = 1, SIDE_NEIGHBOR = 2, SIDE_CF = 3 };

// Ghost fill constants of one side.  SIDE_PHYS: Robin (BCToolsF.ChF:222-337), a = alpha/8 (2 cells)
// or alpha/2 (1 cell), bb = beta/dx.  SIDE_CF: ghost = a p1 + bb p2 with a = c1 = 2(dxc-dxf)/(dxc+dxf),
// bb = c2 = -(dxc-dxf)/(dxc+3dxf) (twoCells) or ghost = a p1 with a = 1 - 2dxf/(dxf+dxc).
struct SideBC {
    int    kind;       // SideKind
    int    twoCells;   // numValidCells >= 2 (BCTools.cpp:381; CFInterp.cpp:445,455)
    double a, bb;      // as above
    int    neighbor;   // rank (SIDE_NEIGHBOR)
};
// a side whose ghosts are a function of the interior cells next to it (refreshed once per relaxation
// iteration, PoissonOp.cpp:1957-1962), as opposed to copies of other valid cells
__host__ __device__ inline bool sideIsBC(int kind) { return kind == SIDE_PHYS || kind == SIDE_CF; }
// value of the ghost cell from the first (p0) and second (p1) interior cell
__host__ __device__ inline double sideGhost(const SideBC& bc, double p0, double p1)
{
    if (bc.kind == SIDE_PHYS) {
        if (bc.twoCells) {  // BCToolsF.ChF:313-331 (homogeneous branch)
            const double cg = 3.0 * bc.a + bc.bb;
            const double c0 = 6.0 * bc.a - bc.bb;
            const double c1 = -1.0 * bc.a;
            return -(c0 * p0 + c1 * p1) / cg;
        }
        const double cg = bc.a + bc.bb;  // BCToolsF.ChF:255-270
        const double c0 = bc.a - bc.bb;
        return -(c0 * p0) / cg;
    }
    // SIDE_CF: CFInterpF.ChF:305-336 (quadratic) / :340-370 (linear)
    if (bc.twoCells) return bc.a * p0 + bc.bb * p1;
    return bc.a * p0;
}

// Per-depth coefficient pointers handed to kernels.
struct Coef {
    const double* J;
    const double* Dinv;
    const double* mxl; const double* mxr;  // [nx] lower / upper off-diagonals, tile-local index
    const double* myl; const double* myr;  // [ny]
    const double* mzl; const double* mzr;  // [nz]
    const double* loBC; const double* hiBC;  // [py*px] slabs (index = OX+i + sy*(1+j))
    double beta;
    // J and Dinv as functions of the level only ([nz] each), set when every column of the tile holds the same values
    // BIT FOR BIT (Op::detectColumnCoefficients): kernels then read the two tables instead of the two arrays
    const double* tabJ; const double* tabD;
};

struct BoxList {  // boxes of this rank in tile-local coordinates, on the device
    int  n;
    int* lo;  // [n][3]
    int* hi;  // [n][3]
    int  maxnx;  // widest box (host-side hint for the reduction's thread shape)
};

// ---- kernel launchers (sb_kernels.cu).  All asynchronous on `st`. -------------------------
namespace k {
void fill(cudaStream_t st, double* a, long long n, double v);
void copy_valid(cudaStream_t st, const Lay& L, double* dst, const double* src);
void scale_valid(cudaStream_t st, const Lay& L, double* a, double s, int faceDir);
void incr_valid(cudaStream_t st, const Lay& L, double* y, const double* x, double s, int faceDir);
void axby_valid(cudaStream_t st, const Lay& L, double* z, const double* x, const double* y, double a, double b);
void add_scalar_valid(cudaStream_t st, const Lay& L, double* a, const double* sumvol /* dev: sum, vol */);
void mult_valid(cudaStream_t st, const Lay& L, double* dst, const double* src, const double* m);  // dst = src*m

// ghost fill of the six tile sides (physical Robin BC or periodic self-copy)
void fill_ghosts(cudaStream_t st, const Lay& L, double* phi, const SideBC bc[3][2], int dim);
// faces, then edges (needed only by the quadratic prolongation's mixed differences)
void fill_ghosts_with_edges(cudaStream_t st, const Lay& L, double* phi, const SideBC bc[3][2], int dim);
long long launch_count();
void note_launches(long long n);  // kernels replayed by a CUDA graph launch
// pack / unpack of one side's face layer for neighbour exchange
void fill_ghosts_dir(cudaStream_t st, const Lay& L, double* phi, int dir, const SideBC& lo, const SideBC& hi, int ext0, int ext1);
void extrap_domain_edges(cudaStream_t st, const Lay& L, double* phi, const SideBC bc[3][2], int dim);
size_t face_count(const Lay& L, int dir, int ext0, int ext1);
void pack_face(cudaStream_t st, const Lay& L, const double* phi, int dir, int side, double* buf, int ext0 = 0, int ext1 = 0);
void unpack_face(cudaStream_t st, const Lay& L, double* phi, int dir, int side, const double* buf, int ext0 = 0, int ext1 = 0);

void apply_op(cudaStream_t st, const Lay& L, const Coef& c, double* lhs, const double* phi);
void residual(cudaStream_t st, const Lay& L, const Coef& c, double* res, const double* phi, const double* rhs);
void gsrb_pass(cudaStream_t st, const Lay& L, const Coef& c, double* phi, const double* rhs, int pass);
void jacobi(cudaStream_t st, const Lay& L, const Coef& c, double* phi, const double* res, int pass /* -1 all */);
// one colour of vertical line relaxation; wd/wb are scratch of nz*ny*((nx+1)/2) doubles;
// *pivotFlag is set to 1 if LAPACK dgtsv would have interchanged rows.
void vertline_pass(cudaStream_t st, const Lay& L, const Coef& c, double* phi, const double* rhs, int pass,
                   double* wd, double* wb, int* pivotFlag);

// shared-matrix fast path (see sb_kernels.cu); tab = [4][nz]
size_t vertline_smem_bytes(int nz);
void vertline_smem_pass(cudaStream_t st, const Lay& L, const Coef& c, const double* tab, double* phi, const double* rhs, int pass);
// colour-split storage (SLay): conversion, ghost fill / face pack of directions x and y, and the
// line relaxation on it.  s[c] = array of colour c.
// scale (per level) or scaleJ / beta (per cell: value / (beta * scaleJ)) may be given, not both
void split_field(cudaStream_t st, const Lay& L, const SLay& S, const double* nat, double* s0, double* s1, const double* scale,
                 const double* shift = nullptr, const double* scaleJ = nullptr, double beta = 1.0);
// tabD (may be null): Dinv as a function of the level only, read instead of the array
void split_precond(cudaStream_t st, const Lay& L, const SLay& S, const double* res, const double* Dinv, const double* scale,
                   double* c0, double* c1, double* r0, double* r1, const double* scaleJ = nullptr, double beta = 1.0,
                   const double* tabD = nullptr);
void unsplit_field(cudaStream_t st, const Lay& L, const SLay& S, double* nat, const double* s0, const double* s1);
void fill_ghosts_split(cudaStream_t st, const SLay& S, double* s0, double* s1, const SideBC bc[3][2], int dim, bool physToo);
// vertical sides of a split field with z ghosts (S.zg = 1): Robin / periodic, as fill_ghosts_dir_k
void fill_ghosts_split_z(cudaStream_t st, const SLay& S, double* s0, double* s1, const SideBC& lo, const SideBC& hi, bool physToo);
// one colour of point red-black Gauss-Seidel on split storage (PoissonOpF.ChF:420-474); p / r / J / Dinv: the two
// colour arrays of phi, rhs, J and Dinv
void gsrb_split_pass(cudaStream_t st, const SLay& S, const Coef& c, double* const p[2], const double* const r[2],
                     const double* const J[2], const double* const Di[2], int pass);
void pack_faces_split(cudaStream_t st, const SLay& S, double* s0, double* s1, double* const bufs[2][2], bool unpack);
void vertline_split_pass(cudaStream_t st, const SLay& S, const Coef& c, const double* tab, double* own, const double* oth,
                         const double* rhs, int pass, int region = 0, int nbMask = 0);
bool vertline_split_fits(int nz);
int  vertline_split_chunk(int nz);  // levels per warp chunk (the tables depend on it)
int  vertline_split_nw();           // chunks per column
bool vertline_split_fused();        // which of the two kernels (and table layouts) is in use
void j_deviation(cudaStream_t st, const Lay& L, const double* J, const double* jcol, double* out);

void restrict_avg(cudaStream_t st, const Lay& Lf, const Lay& Lc, const int ref[3], double* crse, const double* fine);
void prolong_const(cudaStream_t st, const Lay& Lf, const Lay& Lc, const int ref[3], double* fine, const double* crse);
void prolong_linear(cudaStream_t st, const Lay& Lf, const Lay& Lc, const int ref[3], double* fine, const double* crse);
void prolong_quad1(cudaStream_t st, const Lay& Lf, const Lay& Lc, const int ref[3], double* fine, const double* crse);
void prolong_quad2(cudaStream_t st, const Lay& Lf, const Lay& Lc, const int ref[3], double* fine, const double* crse, int dim);
void restrict_face(cudaStream_t st, const Lay& Lf, const Lay& Lc, const int ref[3], int dir, double* crse, const double* fine);

void compute_dinv(cudaStream_t st, const Lay& L, const Coef& c, double alpha, double* Dinv, int dim);
void compute_vert_bcs(cudaStream_t st, const Lay& L, const Coef& c, double sLo, double sHi, double* loBC, double* hiBC);
// J / Jgup from per-box 1-D dx/dXi tables (cell tables c*, node tables f*), box in tile-local indices
void fill_metric_box(cudaStream_t st, const Lay& L, const int blo[3], const int bhi[3], const double* cx, const double* cy,
                     const double* cz, const double* fx, const double* fy, const double* fz, double* J, double* Jg0,
                     double* Jg1, double* Jg2);

void divergence(cudaStream_t st, const Lay& L, double* div, const double* u0, const double* u1, const double* u2,
                double dxinv0, double dxinv1, double dxinv2, int dim);
void gradient(cudaStream_t st, const Lay& L, double* g, const double* phi, const double* Jgup, int dir, double oneOnDx,
              double beta, int scaleBeta);

void scale_faces_box(cudaStream_t st, const Lay& L, const int blo[3], const int n[3], int mu1, int mu2, const double* t1,
                     const double* t2, double* vel, bool divide);

// Leptic solver leaves.  F is the layout of the flattened (one-layer) fields.
// excess = hiBC - sum_k rhs*dz, k ascending (LevelLepticSolver.cpp:725-770, SubspaceF.ChF:33-58)
void vert_excess(cudaStream_t st, const Lay& L, const Lay& F, double* excess, const double* hiBC, const double* rhs, double dzScale);
// FORT_TRIDIAGPOISSONNN1DFAB (PoissonOpF.ChF:1329-1410); gam: cell-sized scratch
void tridiag_nn(cudaStream_t st, const Lay& L, const Lay& F, double* phi, const double* rhs, const double* upperBC,
                const double* sigma, double* gam, double dx);
// FORT_ADDVERTICALEXTRUSION (SubspaceF.ChF:66-110)
void add_vertical_extrusion(cudaStream_t st, const Lay& L, const Lay& F, double* dest, const double* flat);

// Reductions.  op: 0 max|x|, 1 sum|x|, 2 sum x^2, 3 sum x*y, 4 sum (J*dv)*x and sum J*dv (2 outputs).
// One result per box (or 2 for op 4) lands in out[] (device); partial is scratch.
// mask (may be null): a box in tile-local indices whose cells count as zero (AMRNormLevel).
// ytab (may be null): y as a function of the level only, read instead of the array y
void reduce_boxes(cudaStream_t st, const Lay& L, const BoxList& boxes, int op, const double* x, const double* y,
                  double dv, double* partial, double* out, const Box3* mask = nullptr, const double* ytab = nullptr);
int  reduce_partial_len(int nboxes);

// The tail of a V-cycle in one launch (sb_tiny.cu): vCycle_residualEq over the deepest depths (as many as fit a shared-memory
// arena together, owned by one rank), bottom smooths and BiCGStab solve included.
struct TinyLevel {
    Lay           L;
    Coef          c;
    SideBC        side[3][2];
    int           dim, relaxMethod;
    const int*    boxLo;   // [nboxes][3], tile-local
    const int*    boxHi;
    int           nboxes;
    int           hasNullSpace;
    double        dv;      // cell volume in index space (removeKernel's weights are J dv)
    int           ref[3];  // refinement ratio to the next (coarser) level of the tail
    double *      cor, *res, *tmp;
    int*          pivotFlag;
    const double* lineTab;  // [4][nz] factorisation tables of the shared-matrix line relaxation (Op::lineTab), or null
};
constexpr int TINY_MAXLEV = 4;
struct TinyTailArgs {
    int               nlev;            // lev[nlev - 1] is the bottom of the hierarchy
    TinyLevel         lev[TINY_MAXLEV];
    sb_bottom_options opt;
    int               numSmoothDown, numSmoothUp, numSmoothBottom, prolongOrder, corIsPreCond;
    double*           w[8];            // BiCGStab work vectors on the bottom level: r, r_tilde, e, p, p_tilde, s_tilde, t, v
    double*           out;             // status, initResNorm, finalResNorm, iterations, restarts of the bottom solve (may be null)
};
bool   tiny_level_fits(const Lay& L, int nboxes);
size_t tiny_level_bytes(const Lay& L, int nboxes, bool nonBottom, bool bottom);  // shared memory its staged copy takes
size_t tiny_arena_limit();
void   tiny_tail(cudaStream_t st, const TinyTailArgs& args, size_t arenaBytes);
void sum_boxes(cudaStream_t st, const double* in, int nboxes, int ncomp, double* out);
}  // namespace k

}  // namespace sb
