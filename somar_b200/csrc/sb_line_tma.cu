// sb_line_tma.cu -- vertical line relaxation, one colour pass, as a persistent warp-specialised kernel whose
// operands arrive by TMA (cp.async.bulk.tensor, sm_100a): vertline_tma_k.
//
// Reference: PoissonOp::vertLineGSRB_relax (Grade3_Calculus/Elliptic/PoissonOp.cpp:1927-2010),
// FORT_POISSONOP_VERTLINEGSRB_3D (Elliptic/PoissonOpF.ChF:851-1019) + LAPACK dgtsv (no-interchange branch).
//
// Why.  vertline_fused_k (sb_line.cu) moves exactly the bytes it has to (12 B per grid cell per pass) but
// reaches 0.68 of the copy bandwidth: its loads are held in registers (128 per thread, 2 CTAs per SM) and
// stop while a CTA resolves its carries and streams its backward sweep out (profiles/r1_v4_summary.md).
// Here one CTA per SM walks over tiles (32 columns of one colour in one grid row, all levels):
//   * a producer thread issues 3-D tensor-map loads -- the other colour's rows j-1, j, j+1 (36 wide: west
//     and east neighbours come out of the same box) and this colour's right-hand side -- KZ levels per box,
//     one ring of S stages per consumer warp, completion on mbarriers.  Bytes in flight cost shared memory,
//     not registers, and the ring runs ahead across tile boundaries, so HBM requests never stop;
//   * NW consumer warps each own a chunk of nz / NW levels and run the chunked Thomas sweeps of
//     vertline_fused_k (same arithmetic, operation for operation: results are bit-identical).
//
// GENERAL = true is the mapped-grid form of the same kernel: with a horizontally varying metric the row-scaled
// tridiagonal matrix of a column is T_z - h(i, j) I, h = MxL + MxR + MyL + MyR, so its Thomas factorisation
// differs from column to column.  It is recomputed in the kernel, per lane (one reciprocal per cell), from the
// 1-D tables M_x, M_y, M_z and the value of g = 1 / d' at the level below each chunk (a small 2-D table built
// once per operator by line_gstart_k); the prefix products that vertline_fused_k reads from 1-D tables become
// running products, and the chunk's g values are parked in shared memory for the fix-up and backward sweeps.
// No J / Dinv operand, no HBM scratch: still 12 B per grid cell per pass (vertline_k: 43 B measured).
#include <cuda.h>

#include "sb_core.h"
#include "sb_line_tma.h"

namespace sb {
namespace k {

void note_launch();  // sb_kernels.cu (launch counter)

namespace {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Blocks until the phase with the given parity has completed.  A watchdog turns a protocol error (a load that never
// lands) into a launch failure instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* fault = nullptr, int code = 0)
{
    const uint32_t a = smem_u32(bar);
    uint32_t       done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (spin > (1u << 22)) {
            if (fault) {  // mapped host memory: readable after the launch has failed
                fault[1] = code; fault[2] = (int)blockIdx.x; fault[3] = (int)threadIdx.x; fault[4] = (int)parity;
                fault[0] = 1;
                __threadfence_system();
            }
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void consumer_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

__device__ __forceinline__ unsigned int ld_acquire_cta_shared(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_cta_shared_add(unsigned int* p, unsigned int v)
{
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Schedule slot u -> tile (bx, j).  Plain order without a fused exchange; otherwise the tiles touching an exchanged side
// first: row 0 (nA), row ny-1 (nB), column 0 (nC), column nbx-1 (nD), then the interior rectangle.
__host__ __device__ __forceinline__ void tile_of(const LineTmaArgs& A, int ny, int u, int& bx, int& j)
{
    if (!A.halo) { bx = u % A.nbx; j = u / A.nbx; return; }
    if (u < A.nA) { bx = u; j = 0; return; }
    u -= A.nA;
    if (u < A.nB) { bx = u; j = ny - 1; return; }
    u -= A.nB;
    if (u < A.nC) { bx = 0; j = A.jlo + u; return; }
    u -= A.nC;
    if (u < A.nD) { bx = A.nbx - 1; j = A.jlo + u; return; }
    u -= A.nD;
    bx = A.bxlo + u % A.nbxi;
    j  = A.jlo + u / A.nbxi;
}
// sides of the tile an exchanged face layer lies on: bit 2 * dir + side
__host__ __device__ __forceinline__ int tile_touch(const LineTmaArgs& A, int ny, int bx, int j)
{
    if (!A.halo) return 0;
    return (((A.nbMask & 1) && bx == 0) ? 1 : 0) | (((A.nbMask & 2) && bx == A.nbx - 1) ? 2 : 0) | (((A.nbMask & 4) && j == 0) ? 4 : 0) |
           (((A.nbMask & 8) && j == ny - 1) ? 8 : 0);
}
// One CTA has finished a tile on the sides in `touch` (its stores fenced, its threads past a barrier): count it, and
// publish the side's arrival counter in the neighbour's memory when it was the last one.
__device__ __forceinline__ void halo_signal(const LineTmaArgs& A, int ny, int touch)
{
    const HaloDev& H = *A.halo;
    for (int s = 0; s < 4; ++s) {
        if (!((touch >> s) & 1)) continue;
        const unsigned int count = s < 2 ? (unsigned int)ny : (unsigned int)A.nbx;  // tiles along an x side / a y side
        __threadfence_system();
        const unsigned int old = atomicAdd(H.done + s, 1u);
        if (old == count - 1) {
            H.done[s] = 0;
            const unsigned long long e = H.posted[s] + 1;
            H.posted[s] = e;
            __threadfence_system();
            st_release_sys_u64(H.side[s].rflag, e);
        }
    }
}

constexpr int OTHW = 36;  // box width of the other colour's rows: cells m0 - 2 .. m0 + 33 (the innermost TMA coordinate must be
                          // 16-byte aligned, i.e. even for doubles -- measured: an odd start is an illegal instruction)
// One stage holds the boxes of ALL chunks: the tensor maps view z as (level in chunk, chunk), so a single 4-D box
// [NW chunks][KZ levels][3 rows][36] (and [NW][KZ][1][32] of the right-hand side) serves the eight consumer warps at
// once -- 16 bulk copies per tile instead of 128 (one issuing thread could not keep up with the small ones: measured).
__host__ __device__ constexpr int othBytes(int KZ, int NW) { return ((OTHW * 3 * KZ * NW * 8 + 127) / 128) * 128; }
__host__ __device__ constexpr int rhsBytes(int KZ, int NW) { return ((32 * KZ * NW * 8 + 127) / 128) * 128; }
}  // namespace

// tab (shared matrix): {a', g}[N] | {P', c}[N] | R[N] | Pend[NW] | T[NW] | Rend[NW]   (Op::buildLineTables)
// tab (general):       MzL[N] | MzR[N]
template <int NW, int KZ, int S, bool GENERAL>
__global__ void __launch_bounds__((NW + 2) * 32, 1)
    vertline_tma_k(const __grid_constant__ CUtensorMap mapOth, const __grid_constant__ CUtensorMap mapRhs, SLay Sl, LineTmaArgs A)
{
    extern __shared__ __align__(128) unsigned char smraw[];
    const int N  = Sl.nz;
    const int CL = N / NW;
    constexpr int OB = othBytes(KZ, NW), RB = rhsBytes(KZ, NW), SB = OB + RB;
    unsigned char* const stages = smraw;                                   // [S][SB]
    double* const  sy  = reinterpret_cast<double*>(smraw + (size_t)S * SB);  // [N][32]
    double* const  sg  = sy + (size_t)N * 32;                              // [N][32] (GENERAL)
    double* const  sum = sg + (GENERAL ? (size_t)N * 32 : 0);              // [2 sets][NSUM][NW][32]
    constexpr int  NSUM = GENERAL ? 5 : 2;
    double* const  ts  = sum + 2 * NSUM * NW * 32;                         // tables
    const int      ntab = GENERAL ? 2 * N : 5 * N + 3 * NW;
    uint64_t* const bars = reinterpret_cast<uint64_t*>(ts + ((ntab + 1) & ~1));  // full[S], empty[S]
    uint64_t* const full = bars;
    uint64_t* const empt = bars + S;
    unsigned int* const tilesDone = reinterpret_cast<unsigned int*>(bars + 2 * S);  // consumer warps that have finished an exchanged tile

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nbx = A.nbx, ntiles = A.ntiles, nby = Sl.ny;
    auto in_region = [&](int t) -> bool {
        if (A.region == 0) return true;
        const int  bx = t % nbx, j = t / nbx;
        const bool edge = ((A.nbMask & 1) && bx == 0) || ((A.nbMask & 2) && bx == nbx - 1) || ((A.nbMask & 4) && j == 0) ||
                          ((A.nbMask & 8) && j == nby - 1);
        return edge == (A.region == 1);
    };
    auto next_tile = [&](int t) { do { t += gridDim.x; } while (t < ntiles && !in_region(t)); return t; };
    int t0 = blockIdx.x;
    while (t0 < ntiles && !in_region(t0)) t0 += gridDim.x;

    for (int k = threadIdx.x; k < ntab; k += (NW + 2) * 32) ts[k] = A.tab[k];
    if (threadIdx.x == 0) {
        *tilesDone = 0;
        for (int i = 0; i < S; ++i) { mbar_init(full + i, 1); mbar_init(empt + i, NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nslab = CL / KZ;

    if (w == NW) {
        // ===== producer: one thread feeds every consumer warp's ring =====
        if (lane == 0) {
            int      stage = 0;
            uint32_t phase = 0;
            for (int t = t0; t < ntiles; t = next_tile(t)) {
                int bx, j;
                tile_of(A, nby, t, bx, j);
                const int x0 = SOX + bx * 32 - 2;  // element of cell m0 - 2 in a row of a colour array (even)
                for (int sl = 0; sl < nslab; ++sl) {
                    uint64_t* fb = full + stage;
                    mbar_wait(empt + stage, phase ^ 1, A.fault, 1000 + stage);
                    unsigned char* dst = stages + (size_t)stage * SB;
                    mbar_expect_tx(fb, (OTHW * 3 + 32) * KZ * NW * 8);
                    tma_load_4d(dst, &mapOth, x0, j, sl * KZ, 0, fb);               // rows j-1 .. j+1 (array rows j .. j+2), every chunk
                    tma_load_4d(dst + OB, &mapRhs, x0 + 2, j + 1, sl * KZ, 0, fb);  // row j
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }
    if (w == NW + 1) {
        // ===== signaller (fused exchange): counts this CTA's tiles on exchanged sides as the consumers finish them and
        // publishes a side's arrival counter in the neighbour's memory when its last tile is in.  The system-scope fences
        // this takes (they wait for the SM's outstanding stores) stay off the consumer warps.
        if (lane == 0 && A.halo) {
            unsigned int seen = 0;
            for (int t = t0; t < ntiles; t = next_tile(t)) {
                int bx, j;
                tile_of(A, nby, t, bx, j);
                const int touch = tile_touch(A, nby, bx, j);
                if (!touch) continue;
                seen += NW;
                unsigned int spin = 0;
                while (ld_acquire_cta_shared(tilesDone) < seen) {
                    __nanosleep(100);
                    if (++spin > (1u << 24)) {
                        if (A.fault) { A.fault[1] = 4000; A.fault[2] = (int)blockIdx.x; A.fault[3] = (int)seen; A.fault[4] = 0; A.fault[0] = 1; __threadfence_system(); }
                        __trap();
                    }
                }
                halo_signal(A, nby, touch);
            }
        }
        return;
    }

    // ===== consumers =====
    const double2* const T1 = reinterpret_cast<const double2*>(ts);          // shared: {a', g}
    const double2* const T2 = reinterpret_cast<const double2*>(ts + 2 * N);  // shared: {P', c}
    const double*  const tR = ts + 4 * N;
    const double*  const tPend = ts + 5 * N;
    const double*  const tT    = tPend + NW;
    const double*  const tRend = tT + NW;
    const double*  const zl_ = ts;       // general: MzL
    const double*  const zr_ = ts + N;   // general: MzR
    const int       k0 = w * CL, k1 = k0 + CL;
    const long long szl = Sl.sz;
    int             stage = 0;
    uint32_t        phase = 0;
    int             set = 0;
    for (int t = t0; t < ntiles; t = next_tile(t)) {
        int bx, j;
        tile_of(A, nby, t, bx, j);
        const int  touch = tile_touch(A, nby, bx, j);
        const int  i0 = (A.pass + Sl.par + j) & 1;  // own cells of this row: i = 2 m + i0
        const int  m  = bx * 32 + lane;
        const bool act = 2 * m + i0 < Sl.nx;
        const int  ii = 2 * (act ? m : 0) + i0;     // idle lanes shadow column 0 (their boxes are zero-filled) and never store
        const double mxl = A.mx[ii], mxr = A.mx[Sl.nx + ii], myl = A.my[j], myr = A.my[Sl.ny + j];
        double* const cz = sum + (size_t)set * NSUM * NW * 32;
        double* const ca = cz + NW * 32;
        double zl = 0.0, acc = 0.0;
        // general: running products of the chunk and the factorisation carried in from the level below it
        double P = 1.0, R = 1.0, T = 0.0, g = 0.0, h = 0.0;
        if (GENERAL) {
            h = mxl + mxr + myl + myr;
            if (w > 0) g = A.gstart[(size_t)(w - 1) * Sl.nx * Sl.ny + (size_t)j * Sl.nx + ii];
        }
        // P1: right-hand sides, local forward sweep, the chunk's contribution to its first unknown
        for (int sl = 0; sl < nslab; ++sl) {
            const double* const bo = reinterpret_cast<const double*>(stages + (size_t)stage * SB) + (size_t)w * KZ * 3 * OTHW;
            const double* const br = reinterpret_cast<const double*>(stages + (size_t)stage * SB + OB) + (size_t)w * KZ * 32;
            mbar_wait(full + stage, phase, A.fault, 2000 + stage * 16 + w);
            double a[KZ][5];
#pragma unroll
            for (int u = 0; u < KZ; ++u) {
                a[u][0] = bo[(u * 3 + 1) * OTHW + lane + i0 + 1];  // west  (box element 0 is cell m0 - 2)
                a[u][1] = bo[(u * 3 + 1) * OTHW + lane + i0 + 2];  // east
                a[u][2] = bo[(u * 3 + 0) * OTHW + lane + 2];       // south
                a[u][3] = bo[(u * 3 + 2) * OTHW + lane + 2];       // north
                a[u][4] = br[u * 32 + lane];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empt + stage);  // this warp's part of the box is in registers: hand the slot back
            if (++stage == S) { stage = 0; phase ^= 1; }
#pragma unroll
            for (int u = 0; u < KZ; ++u) {
                const int    k    = k0 + sl * KZ + u;
                const double lphi = fma(myr, a[u][3], fma(myl, a[u][2], fma(mxr, a[u][1], mxl * a[u][0])));
                const double b    = a[u][4] - lphi;
                if (!GENERAL) {
                    const double2 tt = T1[k];
                    zl  = fma(tt.x, zl, tt.y * b);
                    acc = fma(tR[k], zl, acc);
                } else {
                    // row k of the column system divided by beta J_k (sb_op.cpp: Op::buildLineTables, the same statements)
                    const double ml = zl_[k], mr = zr_[k];
                    double       diag = A.aob - h - ml - mr;
                    if (k == 0) diag += -ml * A.sLo;
                    if (k == N - 1) diag += -mr * A.sHi;
                    double d = diag;
                    if (k > 0) d = diag - (ml * g) * zr_[k - 1];
                    g = __drcp_rn(d);  // correctly rounded, as 1.0 / d
                    const double aa = k > 0 ? -(ml * g) : 0.0;
                    const double cc = k < N - 1 ? -(mr * g) : 0.0;
                    zl  = fma(aa, zl, g * b);
                    P   = P * aa;
                    acc = fma(R, zl, acc);
                    T   = fma(R, P, T);
                    R   = R * cc;
                    sg[k * 32 + lane] = g;
                }
                sy[k * 32 + lane] = zl;
            }
        }
        cz[w * 32 + lane] = zl;
        ca[w * 32 + lane] = acc;
        if (GENERAL) {
            cz[(2 * NW + w) * 32 + lane] = P;
            cz[(3 * NW + w) * 32 + lane] = T;
            cz[(4 * NW + w) * 32 + lane] = R;
        }
        consumer_bar(NW * 32);

        // Carries.  Zs[v]: true z just below chunk v; X: true x just above this warp's chunk.
        double Zs[NW];
        double Z = 0.0, Zm = 0.0;
#pragma unroll
        for (int v = 0; v < NW; ++v) {
            Zs[v] = Z;
            if (v == w) Zm = Z;
            const double pe = GENERAL ? cz[(2 * NW + v) * 32 + lane] : tPend[v];
            Z = fma(pe, Z, cz[v * 32 + lane]);
        }
        double X = 0.0;
#pragma unroll
        for (int v = NW - 1; v >= 1; --v)
            if (v > w) {
                const double re = GENERAL ? cz[(4 * NW + v) * 32 + lane] : tRend[v];
                const double tv = GENERAL ? cz[(3 * NW + v) * 32 + lane] : tT[v];
                X = fma(re, X, fma(Zs[v], tv, ca[v * 32 + lane]));
            }

        // P2: true backward sweep of this chunk, straight to HBM -- and, for cells of an exchanged face layer, into the
        // neighbour's ghost cell as well (rq[0]: across a y side, rq[1]: across an x side; a corner cell feeds both).
        const long long base = (long long)(SOX + (act ? m : 0)) + Sl.sy * (long long)(1 + j);
        double* po = A.own + base + (long long)(k1 - 1) * szl;
        double  xl = X;
        double*   rq[2]  = {nullptr, nullptr};
        long long rqs[2] = {0, 0};
        if (touch && act) {
            const HaloDev& H = *A.halo;
            if (touch & 12) {
                const HaloPeerSide& Pq = H.side[(touch & 4) ? 2 : 3];
                rq[0]  = Pq.rs[A.pass] + (long long)(SOX + m) + Pq.rsy * (long long)Pq.rrow + Pq.rsz * (long long)(k1 - 1);
                rqs[0] = Pq.rsz;
            }
            const int sx = ((touch & 1) && ii == 0) ? 0 : ((touch & 2) && ii == Sl.nx - 1) ? 1 : -1;
            if (sx >= 0) {
                const HaloPeerSide& Pq = H.side[sx];
                rq[1]  = Pq.rs[A.pass] + (long long)Pq.rx + Pq.rsy * (long long)(1 + j) + Pq.rsz * (long long)(k1 - 1);
                rqs[1] = Pq.rsz;
            }
        }
        // warp-uniform choice of the loop: a lane with a neighbour-bound cell drags its warp through the predicated stores
        const bool remote = touch && __any_sync(0xffffffffu, rq[0] || rq[1]);
        if (GENERAL) {
            // true forward values first: z_k = zl_k + P'_k Z (P'_k recomputed exactly as in P1)
            double Pp = 1.0;
            for (int k = k0; k < k1; ++k) {
                const double aa = k > 0 ? -(zl_[k] * sg[k * 32 + lane]) : 0.0;
                Pp = Pp * aa;
                sy[k * 32 + lane] = fma(Pp, Zm, sy[k * 32 + lane]);
            }
#pragma unroll 4
            for (int k = k1 - 1; k >= k0; --k) {
                const double cc = k < N - 1 ? -(zr_[k] * sg[k * 32 + lane]) : 0.0;
                xl = fma(cc, xl, sy[k * 32 + lane]);
                if (act) *po = xl;
                po -= szl;
                if (remote) {
                    if (rq[0]) { *rq[0] = xl; rq[0] -= rqs[0]; }
                    if (rq[1]) { *rq[1] = xl; rq[1] -= rqs[1]; }
                }
            }
        } else if (!remote) {
#pragma unroll 4
            for (int k = k1 - 1; k >= k0; --k) {
                const double2 tt = T2[k];
                xl = fma(tt.y, xl, fma(tt.x, Zm, sy[k * 32 + lane]));
                if (act) *po = xl;
                po -= szl;
            }
        } else {
#pragma unroll 4
            for (int k = k1 - 1; k >= k0; --k) {
                const double2 tt = T2[k];
                xl = fma(tt.y, xl, fma(tt.x, Zm, sy[k * 32 + lane]));
                if (act) *po = xl;
                po -= szl;
                if (rq[0]) { *rq[0] = xl; rq[0] -= rqs[0]; }
                if (rq[1]) { *rq[1] = xl; rq[1] -= rqs[1]; }
            }
        }
        // hand the tile to the signaller: release at CTA scope, so that its system-scope fence covers this warp's stores
        if (touch) {
            __syncwarp();
            if (lane == 0) red_release_cta_shared_add(tilesDone, 1u);
        }
        set ^= 1;
    }
}

// g = 1 / d' of the column factorisation at the level below every chunk but the first (levels w CL - 1), one thread per
// column; flag: dgtsv would have interchanged rows (|D'_k| < |DL_{k+1}| on the unscaled system) or met a zero pivot.
__global__ void line_gstart_k(Lay L, const double* __restrict__ J, const double* __restrict__ mx, const double* __restrict__ my,
                              const double* __restrict__ mz, double aob, double sLo, double sHi, int CL, double* __restrict__ gstart,
                              int* __restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= L.nx || j >= L.ny) return;
    const int    N = L.nz;
    const double h = mx[i] + mx[L.nx + i] + my[j] + my[L.ny + j];
    double       g = 0.0, dprev = 0.0;
    int          bad = 0;
    for (int k = 0; k < N; ++k) {
        const double ml = mz[k], mr = mz[N + k];
        double       diag = aob - h - ml - mr;
        if (k == 0) diag += -ml * sLo;
        if (k == N - 1) diag += -mr * sHi;
        double d = diag;
        if (k > 0) {
            const double jk = J[L.idx(i, j, k)], jp = J[L.idx(i, j, k - 1)];
            if (!(fabs(jp * dprev) >= fabs(jk * ml)) || dprev == 0.0) bad = 1;
            d = diag - (ml * g) * mz[N + k - 1];
        }
        if (d == 0.0) bad = 1;
        g     = 1.0 / d;
        dprev = d;
        if ((k + 1) % CL == 0 && k + 1 < N) gstart[(size_t)((k + 1) / CL - 1) * L.nx * L.ny + (size_t)j * L.nx + i] = g;
    }
    if (bad) atomicOr(flag, 1);
}
void line_gstart(cudaStream_t st, const Lay& L, const double* J, const double* mx, const double* my, const double* mz, double aob,
                 double sLo, double sHi, int CL, double* gstart, int* flag)
{
    const dim3 b(32, 4, 1);
    line_gstart_k<<<dim3((L.nx + 31) / 32, (L.ny + 3) / 4), b, 0, st>>>(L, J, mx, my, mz, aob, sLo, sHi, CL, gstart, flag);
    note_launch();
}

// ------------------------------------------------------------------------------------------
// Host side: tensor maps and launch.
// ------------------------------------------------------------------------------------------
namespace {
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);
EncodeTiled encoder()
{
    static EncodeTiled fn = nullptr;
    if (!fn) {
        void*                           p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
            SB_FAIL("cuTensorMapEncodeTiled is not available from this driver");
        fn = (EncodeTiled)p;
    }
    return fn;
}
constexpr int TNW = 8, TKZ = 4;
size_t tma_smem(int nz, bool general, int S)
{
    const int    nsum = general ? 5 : 2;
    const size_t ntab = general ? 2 * (size_t)nz : 5 * (size_t)nz + 3 * TNW;
    return (size_t)S * (othBytes(TKZ, TNW) + rhsBytes(TKZ, TNW)) + ((size_t)nz * 32 * (general ? 2 : 1) + 2 * nsum * TNW * 32 + ((ntab + 1) & ~(size_t)1)) * 8 +
           2 * (size_t)S * 8 + 16;
}
}  // namespace

int  vertline_tma_nw() { return TNW; }
bool vertline_tma_fits(int nz, bool general)
{
    if (nz % (TNW * TKZ) != 0) return false;
    return tma_smem(nz, general, 2) <= 227 * 1024;
}
void vertline_tma_make_map(const SLay& S, const double* array, int boxw, int boxrows, LineTmaMap* out)
{
    static_assert(sizeof(LineTmaMap) >= sizeof(CUtensorMap), "LineTmaMap too small");
    // z is viewed as (level inside a chunk, chunk): one box then spans every chunk's KZ levels
    if (S.zg != 0 || S.nz % TNW != 0) SB_FAIL("vertline_tma: the colour arrays must have no z ghosts and nz a multiple of the chunk count");
    const int        CL = S.nz / TNW;
    const cuuint64_t dims[4]    = {(cuuint64_t)S.px, (cuuint64_t)S.py, (cuuint64_t)CL, (cuuint64_t)TNW};
    const cuuint64_t strides[3] = {(cuuint64_t)S.sy * 8, (cuuint64_t)S.sz * 8, (cuuint64_t)S.sz * 8 * (cuuint64_t)CL};
    const cuuint32_t box[4]     = {(cuuint32_t)boxw, (cuuint32_t)boxrows, (cuuint32_t)TKZ, (cuuint32_t)TNW};
    const cuuint32_t estr[4]    = {1, 1, 1, 1};
    const CUresult   r = encoder()(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(array), dims,
                                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) SB_FAIL("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
}
void vertline_tma_make_maps(const SLay& S, const double* oth, const double* rhs, LineTmaMap* mapOth, LineTmaMap* mapRhs)
{
    vertline_tma_make_map(S, oth, OTHW, 3, mapOth);
    vertline_tma_make_map(S, rhs, 32, 1, mapRhs);
}

// Tile count and, for the fused exchange, the edge-first schedule of a pass (see tile_of)
static void line_tma_schedule(const SLay& S, LineTmaArgs& A)
{
    A.nbx    = ((S.nx + 1) / 2 + 31) / 32;
    A.ntiles = A.nbx * S.ny;
    A.nA = A.nB = A.nC = A.nD = A.jlo = A.bxlo = 0;
    A.nbxi = A.nbx;
    if (A.halo) {
        if (A.region != 0 || S.ny < 2 || S.nx < 2) SB_FAIL("vertline_tma: the fused exchange needs region 0 and a tile of at least 2 x 2 columns");
        const bool xlo = A.nbMask & 1, xhi = A.nbMask & 2, ylo = A.nbMask & 4, yhi = A.nbMask & 8;
        A.nA  = ylo ? A.nbx : 0;
        A.nB  = yhi ? A.nbx : 0;
        A.jlo = ylo ? 1 : 0;
        const int jhi = yhi ? S.ny - 2 : S.ny - 1, nrows = jhi - A.jlo + 1 > 0 ? jhi - A.jlo + 1 : 0;
        A.nC   = xlo ? nrows : 0;
        A.nD   = xhi && !(xlo && A.nbx == 1) ? nrows : 0;
        A.bxlo = A.nC ? 1 : 0;
        const int bxhi = A.nD ? A.nbx - 2 : A.nbx - 1;
        A.nbxi = bxhi - A.bxlo + 1 > 0 ? bxhi - A.bxlo + 1 : 0;
        if (A.nA + A.nB + A.nC + A.nD + A.nbxi * nrows != A.ntiles) SB_FAIL("vertline_tma: tile schedule does not cover the tile");
    }
}
// Host-side view of that schedule (sb_plan_line_tile_order): (bx, j, touched sides) of every slot, in order
int line_tma_tile_order(int nx, int ny, int nbMask, int* order, int capacity)
{
    SLay S{};
    S.nx = nx; S.ny = ny;
    LineTmaArgs A{};
    A.nbMask = nbMask; A.region = 0;
    static const HaloDev dummy{};
    A.halo = nbMask ? &dummy : nullptr;
    line_tma_schedule(S, A);
    for (int u = 0; u < A.ntiles && u < capacity; ++u) {
        int bx, j;
        tile_of(A, ny, u, bx, j);
        order[3 * u] = bx; order[3 * u + 1] = j; order[3 * u + 2] = tile_touch(A, ny, bx, j);
    }
    return A.ntiles;
}

void vertline_tma_pass(cudaStream_t st, const SLay& S, const LineTmaMap& mapOth, const LineTmaMap& mapRhs, const LineTmaArgs& args,
                       bool general)
{
    static int nsm = 0;
    if (!nsm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev); }
    LineTmaArgs A = args;
    line_tma_schedule(S, A);
    const int grid = A.ntiles < nsm ? A.ntiles : nsm;
    const CUtensorMap& mo = reinterpret_cast<const CUtensorMap&>(mapOth);
    const CUtensorMap& mr = reinterpret_cast<const CUtensorMap&>(mapRhs);
#define SB_TMA_LAUNCH(STG, GEN)                                                                                                  \
    {                                                                                                                            \
        const size_t  sh = tma_smem(S.nz, GEN, STG);                                                                             \
        static size_t configured = 0;                                                                                            \
        if (sh > configured) {                                                                                                   \
            SB_CUDA(cudaFuncSetAttribute(vertline_tma_k<TNW, TKZ, STG, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh)); \
            configured = sh;                                                                                                     \
        }                                                                                                                        \
        vertline_tma_k<TNW, TKZ, STG, GEN><<<grid, (TNW + 2) * 32, sh, st>>>(mo, mr, S, A);                                       \
    }
    // as deep a ring as the 227 KB of shared memory allow
    if (!general) {
        if (tma_smem(S.nz, false, 4) <= 227 * 1024) SB_TMA_LAUNCH(4, false)
        else if (tma_smem(S.nz, false, 3) <= 227 * 1024) SB_TMA_LAUNCH(3, false)
        else SB_TMA_LAUNCH(2, false)
    } else {
        if (tma_smem(S.nz, true, 3) <= 227 * 1024) SB_TMA_LAUNCH(3, true)
        else SB_TMA_LAUNCH(2, true)
    }
#undef SB_TMA_LAUNCH
    note_launch();
}

}  // namespace k
}  // namespace sb
