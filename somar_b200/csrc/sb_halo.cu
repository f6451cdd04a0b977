// sb_halo.cu -- see sb_halo.h.
//
// Protocol, per exchanged side, between this tile (A) and the neighbour across it (B):
//   * B owns an arrival counter `flag[side]` inside its exported block; only A writes it, with the number of posts A
//     has made towards B so far.  A counts its posts in `posted[side]`, B the arrivals it has consumed in
//     `waited[side]`, both in private device memory, so the counters survive CUDA-graph replays without host help.
//   * post = halo_post_k: every CTA copies its share of the face layer into B's ghost cells (remote stores), fences
//     to system scope and bumps a completion counter; the CTA that finishes last publishes ++posted in B's flag.
//   * wait = halo_wait_k: one thread per side spins (acquire, system scope) until flag >= ++waited.  The kernels that
//     follow in the stream see the ghost values (kernel boundary).
// Write-after-read safety.  A post towards B overwrites ghosts B may still be reading.  Inside the pass loop of a
// relaxation the previous arrival from B is the licence: B posts colour c only after the kernel that read its colour
// 1 - c ghosts has finished (the post follows it in B's stream), and A consumes that arrival before it runs the pass
// whose result it posts.  At the start of a relaxation call there is no such arrival, so the call opens with a bare
// post (no payload) / wait pair: "everything I did with my ghosts in the previous call is over".
#include <cuda.h>

#include <cstring>
#include <map>

#include "sb_comm.h"
#include "sb_halo.h"

namespace sb {

namespace k {
void note_launch();
}

namespace {
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// colourMask: which of the two arrays' face cells travel (bit c); gsrbPass >= 0: the cells a point red-black pass has
// just updated instead -- at level k they sit in array (gsrbPass + k + lo2) & 1 (gsrb_split_k).
__global__ void __launch_bounds__(256) halo_post_k(SLay S, const double* __restrict__ s0, const double* __restrict__ s1, HaloDev H,
                                                   int colourMask, int gsrbPass)
{
    const int           sideIx = blockIdx.y;
    const HaloPeerSide& P      = H.side[sideIx];
    if (!P.rflag) return;
    const int dir = sideIx >> 1, sd = sideIx & 1;
    if (colourMask || gsrbPass >= 0) {
        const int       nt = dir == 0 ? S.ny : S.nx, nn = dir == 0 ? S.nx : S.ny;
        const int       layer = sd ? nn - 1 : 0;
        const long long total = (long long)nt * S.nz;
        for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < total; m += (long long)gridDim.x * blockDim.x) {
            const int t = (int)(m % nt), kk = (int)(m / nt);
            const int i = dir == 0 ? layer : t, j = dir == 0 ? t : layer;
            const int c = S.colour(i, j);
            if (gsrbPass >= 0 ? c != ((gsrbPass + kk + S.lo2p) & 1) : !((colourMask >> c) & 1)) continue;
            const double    v = (c ? s1 : s0)[S.idx(i, j, kk)];
            const long long q = dir == 0 ? (long long)P.rx + P.rsy * (long long)(1 + t) : (long long)(SOX + (t >> 1)) + P.rsy * (long long)P.rrow;
            P.rs[c][q + P.rsz * (long long)(kk + S.zg)] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();  // release: the CTA's stores (ordered before this point by the barrier) before the count
        const unsigned int old = atomicAdd(H.done + sideIx, 1u);
        if (old == gridDim.x - 1) {
            H.done[sideIx] = 0;
            const unsigned long long e = H.posted[sideIx] + 1;
            H.posted[sideIx] = e;
            __threadfence_system();
            st_release_sys(P.rflag, e);
        }
    }
}

// A neighbour that never posts (a rank that failed, a call sequence that differs between ranks) must not hang the
// GPU: after 20 s the kernel leaves its last words in mapped host memory and traps.
__global__ void halo_wait_k(HaloDev H, int mask)
{
    const int sideIx = threadIdx.x;
    if (sideIx >= 4 || !((mask >> sideIx) & 1)) return;
    const unsigned long long want = H.waited[sideIx] + 1;
    H.waited[sideIx] = want;
    const unsigned long long t0 = globaltimer_ns();
    unsigned int             spin = 0;
    while (ld_acquire_sys(H.flag + sideIx) < want) {
        if ((++spin & 0xfff) == 0 && globaltimer_ns() - t0 > 20000000000ull) {
            if (H.fault) {
                H.fault[1] = 3000 + sideIx; H.fault[2] = (int)want; H.fault[3] = (int)ld_acquire_sys(H.flag + sideIx); H.fault[4] = 0;
                H.fault[0] = 1;
                __threadfence_system();
            }
            __trap();
        }
    }
}

// ---- IPC mappings, shared by every PeerHalo of the process ----
struct Mapping { void* base; int refs; };
std::map<std::string, Mapping>& mappings()
{
    static std::map<std::string, Mapping> m;
    return m;
}
void* openMapping(const cudaIpcMemHandle_t& h)
{
    const std::string key(reinterpret_cast<const char*>(&h), sizeof(h));
    auto              it = mappings().find(key);
    if (it != mappings().end()) { ++it->second.refs; return it->second.base; }
    void* base = nullptr;
    if (cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    mappings()[key] = Mapping{base, 1};
    return base;
}
void closeMapping(void* base)
{
    for (auto it = mappings().begin(); it != mappings().end(); ++it)
        if (it->second.base == base) {
            if (--it->second.refs == 0) { cudaIpcCloseMemHandle(base); mappings().erase(it); }
            return;
        }
}
typedef CUresult (*GetAddressRange)(CUdeviceptr*, size_t*, CUdeviceptr);
GetAddressRange addressRange()
{
    static GetAddressRange fn = nullptr;
    if (!fn) {
        void*                           p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
            SB_FAIL("cuMemGetAddressRange is not available from this driver");
        fn = (GetAddressRange)p;
    }
    return fn;
}

// what a rank tells its neighbours about its block (16 doubles on the wire)
struct Desc {
    cudaIpcMemHandle_t h;          // 64 bytes
    long long          off;        // of the block inside the exported allocation
    long long          offFlag;    // of the flags inside the block (bytes)
    long long          n;          // elements of one colour array
    long long          sy, sz;
    int                nx, ny, ok, pad;
};
static_assert(sizeof(Desc) <= 16 * sizeof(double), "Desc must fit the 16-double message");
}  // namespace

bool PeerHalo::setup(Op& op, const SLay& S, double** s0, double** s1)
{
    ctx = op.ctx;
    Context* root = ctx->parent ? ctx->parent : ctx;
    mask = 0;
    for (int d = 0; d < 2; ++d)
        for (int s = 0; s < 2; ++s)
            if (op.side[d][s].kind == SIDE_NEIGHBOR) mask |= 1 << (2 * d + s);
    const size_t arrBytes  = ((size_t)S.n * sizeof(double) + 255) & ~(size_t)255;
    const size_t flagBytes = 256;
    const size_t privBytes = 256;  // posted[4] | waited[4] | done[4]
    const size_t bytes     = std::max<size_t>(2 * arrBytes + flagBytes + privBytes, (size_t)2 << 20);
    SB_CUDA(cudaMalloc(&block, bytes));
    SB_CUDA(cudaMemsetAsync(block, 0, bytes, ctx->st));
    char* const b = static_cast<char*>(block);
    *s0 = reinterpret_cast<double*>(b);
    *s1 = reinterpret_cast<double*>(b + arrBytes);
    dev        = HaloDev{};
    dev.flag   = reinterpret_cast<unsigned long long*>(b + 2 * arrBytes);
    dev.posted = reinterpret_cast<unsigned long long*>(b + 2 * arrBytes + flagBytes);
    dev.waited = dev.posted + 4;
    dev.done   = reinterpret_cast<unsigned int*>(dev.waited + 4);
    dev.fault  = root->fault;

    // describe the block, swap descriptions with the neighbours (device staging, one NCCL group)
    Desc mine;
    std::memset(&mine, 0, sizeof(mine));
    mine.ok = cudaIpcGetMemHandle(&mine.h, block) == cudaSuccess ? 1 : 0;
    if (!mine.ok) cudaGetLastError();
    CUdeviceptr base = 0;
    size_t      len  = 0;
    if (addressRange()(&base, &len, (CUdeviceptr)block) != CUDA_SUCCESS) mine.ok = 0;
    mine.off     = (long long)((CUdeviceptr)block - base);
    mine.offFlag = (long long)(2 * arrBytes);
    mine.n       = (long long)(arrBytes / sizeof(double));
    mine.sy = S.sy; mine.sz = S.sz; mine.nx = S.nx; mine.ny = S.ny;
    double* stage = nullptr;  // [send 16 | recv 4 x 16]
    SB_CUDA(cudaMalloc((void**)&stage, 5 * 16 * sizeof(double)));
    double hsend[16] = {0};
    std::memcpy(hsend, &mine, sizeof(mine));
    SB_CUDA(cudaMemcpyAsync(stage, hsend, sizeof(hsend), cudaMemcpyHostToDevice, ctx->st));
    std::vector<Comm::Msg> msgs;
    for (int d = 0; d < 2; ++d) {
        for (int s = 0; s < 2; ++s)
            if (mask & (1 << (2 * d + s))) msgs.push_back({stage, 16, op.side[d][s].neighbor, true});
        for (int s = 1; s >= 0; --s)
            if (mask & (1 << (2 * d + s))) msgs.push_back({stage + 16 * (1 + 2 * d + s), 16, op.side[d][s].neighbor, false});
    }
    root->comm->sendRecv(msgs, ctx->st);
    double hrecv[4][16];
    SB_CUDA(cudaMemcpyAsync(hrecv, stage + 16, sizeof(hrecv), cudaMemcpyDeviceToHost, ctx->st));
    ctx->sync();
    SB_CUDA(cudaFree(stage));

    double bad = mine.ok ? 0.0 : 1.0;
    for (int d = 0; d < 2; ++d)
        for (int s = 0; s < 2; ++s) {
            const int ix = 2 * d + s;
            if (!(mask & (1 << ix))) continue;
            Desc theirs;
            std::memcpy(&theirs, hrecv[ix], sizeof(theirs));
            // the face layers must line up: same extent along the side, same vertical extent
            if ((d == 0 && theirs.ny != S.ny) || (d == 1 && theirs.nx != S.nx)) SB_FAIL("peer halo: the neighbouring tile does not line up");
            void* pbase = theirs.ok ? openMapping(theirs.h) : nullptr;
            if (!pbase) { bad = 1.0; continue; }
            opened.push_back(pbase);
            char* const   pb = static_cast<char*>(pbase) + theirs.off;
            HaloPeerSide& P  = dev.side[ix];
            P.rs[0] = reinterpret_cast<double*>(pb);
            P.rs[1] = reinterpret_cast<double*>(pb) + theirs.n;
            P.rflag = reinterpret_cast<unsigned long long*>(pb + theirs.offFlag) + (2 * d + (1 - s));  // the side of B that faces us
            P.rsy = theirs.sy; P.rsz = theirs.sz;
            P.rx   = s ? SOX - 1 : SOX + (theirs.nx >> 1);   // our hi side feeds B's ghost column -1, our lo side its column nx
            P.rrow = s ? 0 : 1 + theirs.ny;                  // likewise ghost row -1 / ny
        }
    root->allreduceMax(&bad, 1);
    if (bad != 0.0) {
        // some rank could not export or map: everybody keeps the NCCL exchange (the arrays stay where they are)
        for (void* p : opened) closeMapping(p);
        opened.clear();
        for (HaloPeerSide& P : dev.side) P = HaloPeerSide{};
        ready = false;
        return false;
    }
    SB_CUDA(cudaMalloc((void**)&devCopy, sizeof(HaloDev)));
    SB_CUDA(cudaMemcpy(devCopy, &dev, sizeof(HaloDev), cudaMemcpyHostToDevice));
    ready = true;
    return true;
}

PeerHalo::~PeerHalo()
{
    for (void* p : opened) closeMapping(p);
    if (devCopy) cudaFree(devCopy);
    if (block) cudaFree(block);
}

void PeerHalo::post(cudaStream_t st, const SLay& S, const double* s0, const double* s1, int colourMask, int gsrbPass)
{
    if (!mask) return;
    int nblk = 1;
    if (colourMask || gsrbPass >= 0) {
        const long long total = (long long)std::max(S.nx, S.ny) * S.nz;
        nblk = (int)std::min<long long>(48, (total + 2047) / 2048);
        if (nblk < 1) nblk = 1;
    }
    halo_post_k<<<dim3(nblk, 4), 256, 0, st>>>(S, s0, s1, dev, colourMask, gsrbPass);
    k::note_launch();
}
void PeerHalo::wait(cudaStream_t st)
{
    if (!mask) return;
    halo_wait_k<<<1, 32, 0, st>>>(dev, mask);
    k::note_launch();
}

}  // namespace sb
