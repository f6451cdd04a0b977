// sb_comm.cpp -- see sb_comm.h.
#include "sb_comm.h"

#include <dlfcn.h>

#include <cstring>

namespace sb {

namespace {
// Minimal NCCL ABI (nccl.h 2.x): opaque communicator, 128-byte unique id, enum values.
typedef struct { char internal[128]; } ncclUniqueId_t;
enum { kNcclSuccess = 0, kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2 };
struct Api {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId_t*)                                                    = nullptr;
    int (*CommInitRank)(void**, int, ncclUniqueId_t, int)                                  = nullptr;
    int (*CommDestroy)(void*)                                                              = nullptr;
    int (*GroupStart)()                                                                    = nullptr;
    int (*GroupEnd)()                                                                      = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t)                        = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t)                              = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t)            = nullptr;
    const char* (*GetErrorString)(int)                                                     = nullptr;
};
Api& api()
{
    static Api a;
    if (a.lib) return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        a.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // reuse the copy torch loaded
        if (a.lib) break;
    }
    if (!a.lib) {
        const char* env = getenv("SB_NCCL_LIB");
        if (env) a.lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    for (const char* n : names) {
        if (a.lib) break;
        a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!a.lib) SB_FAIL(std::string("cannot load NCCL: ") + dlerror());
#define SYM(f)                                                              \
    *(void**)(&a.f) = dlsym(a.lib, "nccl" #f);                              \
    if (!a.f) SB_FAIL("NCCL symbol nccl" #f " not found")
    SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(GroupStart); SYM(GroupEnd); SYM(Send); SYM(Recv);
    SYM(AllReduce); SYM(GetErrorString);
#undef SYM
    return a;
}
#define SB_NCCL(call)                                                                        \
    do {                                                                                     \
        int r__ = (call);                                                                    \
        if (r__ != kNcclSuccess) SB_FAIL(std::string(#call) + ": " + api().GetErrorString(r__)); \
    } while (0)
}  // namespace

void Comm::getUniqueId(void* id128)
{
    ncclUniqueId_t id;
    SB_NCCL(api().GetUniqueId(&id));
    std::memcpy(id128, &id, 128);
}

Comm::Comm(Context* c, const void* id128) : ctx(c)
{
    ncclUniqueId_t id;
    std::memcpy(&id, id128, 128);
    SB_CUDA(cudaSetDevice(ctx->device));
    SB_NCCL(api().CommInitRank(&comm, ctx->nranks, id, ctx->rank));
    SB_CUDA(cudaMalloc((void**)&dscal, 64 * sizeof(double)));
}
Comm::~Comm()
{
    if (comm) api().CommDestroy(comm);
    if (dscal) cudaFree(dscal);
}

void Comm::allreduceHost(double* v, int n, bool isMax)
{
    if (n > 64) SB_FAIL("allreduceHost: n > 64");
    SB_CUDA(cudaMemcpyAsync(dscal, v, n * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
    SB_NCCL(api().AllReduce(dscal, dscal, n, kNcclFloat64, isMax ? kNcclMax : kNcclSum, comm, ctx->st));
    SB_CUDA(cudaMemcpyAsync(v, dscal, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    ctx->sync();
}

void Comm::allreduceDevice(double* dev, int n, bool isMax, cudaStream_t st)
{
    if (!st) st = ctx->st;
    SB_NCCL(api().AllReduce(dev, dev, n, kNcclFloat64, isMax ? kNcclMax : kNcclSum, comm, st));
}

// One ghost layer, faces only, all neighbour sides in one NCCL group.  Sends are issued lo then
// hi and receives hi then lo, so that two ranks that are each other's neighbour on both sides
// (2 ranks along a periodic direction) pair the messages correctly.
void Comm::exchangeFaces(Op& op, double* phi)
{
    bool any = false;
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s)
            if (op.side[d][s].kind == SIDE_NEIGHBOR) { k::pack_face(ctx->st, op.lay, phi, d, s, op.xbuf[d][s][0]); any = true; }
    if (!any) return;
    SB_NCCL(api().GroupStart());
    for (int d = 0; d < 3; ++d) {
        const size_t n = (size_t)(d == 0 ? op.lay.ny : op.lay.nx) * (d == 2 ? op.lay.ny : op.lay.nz);
        for (int s = 0; s < 2; ++s)
            if (op.side[d][s].kind == SIDE_NEIGHBOR)
                SB_NCCL(api().Send(op.xbuf[d][s][0], n, kNcclFloat64, op.side[d][s].neighbor, comm, ctx->st));
        for (int s = 1; s >= 0; --s)
            if (op.side[d][s].kind == SIDE_NEIGHBOR)
                SB_NCCL(api().Recv(op.xbuf[d][s][1], n, kNcclFloat64, op.side[d][s].neighbor, comm, ctx->st));
    }
    SB_NCCL(api().GroupEnd());
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s)
            if (op.side[d][s].kind == SIDE_NEIGHBOR) k::unpack_face(ctx->st, op.lay, phi, d, s, op.xbuf[d][s][1]);
}

void Comm::exchangeFacesSplit(Op& op, double* s0, double* s1, cudaStream_t st, const SLay* Sp)
{
    if (!st) st = ctx->st;
    const SLay& S = Sp ? *Sp : op.slay;
    double* snd[2][2];
    double* rcv[2][2];
    bool    any = false;
    for (int d = 0; d < 2; ++d)
        for (int s = 0; s < 2; ++s) {
            const bool nb = op.side[d][s].kind == SIDE_NEIGHBOR;
            snd[d][s] = nb ? op.xbuf[d][s][0] : nullptr;
            rcv[d][s] = nb ? op.xbuf[d][s][1] : nullptr;
            any = any || nb;
        }
    if (!any) return;
    k::pack_faces_split(st, S, s0, s1, snd, false);
    SB_NCCL(api().GroupStart());
    for (int d = 0; d < 2; ++d) {
        const size_t n = (size_t)(d == 0 ? op.lay.ny : op.lay.nx) * op.lay.nz;
        for (int s = 0; s < 2; ++s)
            if (snd[d][s]) SB_NCCL(api().Send(snd[d][s], n, kNcclFloat64, op.side[d][s].neighbor, comm, st));
        for (int s = 1; s >= 0; --s)
            if (rcv[d][s]) SB_NCCL(api().Recv(rcv[d][s], n, kNcclFloat64, op.side[d][s].neighbor, comm, st));
    }
    SB_NCCL(api().GroupEnd());
    k::pack_faces_split(st, S, s0, s1, rcv, true);
}

// One direction only, optionally extended over the ghosts of the other directions.
void Comm::exchangeDir(Op& op, double* phi, int d, int ext0, int ext1)
{
    bool any = false;
    for (int s = 0; s < 2; ++s)
        if (op.side[d][s].kind == SIDE_NEIGHBOR) { k::pack_face(ctx->st, op.lay, phi, d, s, op.xbuf[d][s][0], ext0, ext1); any = true; }
    if (!any) return;
    const size_t n = k::face_count(op.lay, d, ext0, ext1);
    SB_NCCL(api().GroupStart());
    for (int s = 0; s < 2; ++s)
        if (op.side[d][s].kind == SIDE_NEIGHBOR) SB_NCCL(api().Send(op.xbuf[d][s][0], n, kNcclFloat64, op.side[d][s].neighbor, comm, ctx->st));
    for (int s = 1; s >= 0; --s)
        if (op.side[d][s].kind == SIDE_NEIGHBOR) SB_NCCL(api().Recv(op.xbuf[d][s][1], n, kNcclFloat64, op.side[d][s].neighbor, comm, ctx->st));
    SB_NCCL(api().GroupEnd());
    for (int s = 0; s < 2; ++s)
        if (op.side[d][s].kind == SIDE_NEIGHBOR) k::unpack_face(ctx->st, op.lay, phi, d, s, op.xbuf[d][s][1], ext0, ext1);
}

void Comm::sendRecv(const std::vector<Msg>& msgs, cudaStream_t st)
{
    if (msgs.empty()) return;
    if (!st) st = ctx->st;
    SB_NCCL(api().GroupStart());
    for (const Msg& m : msgs) {
        if (m.send) SB_NCCL(api().Send(m.p, m.n, kNcclFloat64, m.peer, comm, st));
        else SB_NCCL(api().Recv(m.p, m.n, kNcclFloat64, m.peer, comm, st));
    }
    SB_NCCL(api().GroupEnd());
}

// ---------------------------------------------------------------------------------------------
// Agglomeration traffic.  A tile region (valid cells, plus the far face for face-centred data) is
// copied into a dense staging buffer with a device-to-device 3-D copy, sent to / received from
// rank 0 with ncclSend/ncclRecv in one group, and copied into place on the other side.
namespace {
void tileExtent(const Box3& t, int centering, int n[3])
{
    for (int d = 0; d < 3; ++d) n[d] = t.size(d) + (d == centering ? 1 : 0);
}
// dense [n0][n1][n2] (x fastest) <-> layout L at tile-local offset off (cells)
void copyDense(cudaStream_t st, const Lay& L, double* field, const int off[3], const int n[3], double* dense, bool toDense)
{
    cudaMemcpy3DParms p;
    std::memset(&p, 0, sizeof(p));
    cudaPitchedPtr hp = make_cudaPitchedPtr(dense, (size_t)n[0] * sizeof(double), (size_t)n[0], (size_t)n[1]);
    cudaPitchedPtr dp = make_cudaPitchedPtr(field, (size_t)L.px * sizeof(double), (size_t)L.px, (size_t)L.py);
    cudaPos        dpos = make_cudaPos((size_t)(OX + off[0]) * sizeof(double), (size_t)(1 + off[1]), (size_t)(1 + off[2]));
    p.extent = make_cudaExtent((size_t)n[0] * sizeof(double), (size_t)n[1], (size_t)n[2]);
    p.kind   = cudaMemcpyDeviceToDevice;
    if (toDense) { p.srcPtr = dp; p.srcPos = dpos; p.dstPtr = hp; }
    else { p.srcPtr = hp; p.dstPtr = dp; p.dstPos = dpos; }
    SB_CUDA(cudaMemcpy3DAsync(&p, st));
}
}  // namespace

void Comm::gatherTiles(const Op& dist, const double* tileField, const Lay* full, double* fullField, int centering, double* buf)
{
    const int zero[3] = {0, 0, 0};
    int       n[3];
    if (ctx->rank != 0) {
        tileExtent(dist.tile, centering, n);
        copyDense(ctx->st, dist.lay, const_cast<double*>(tileField), zero, n, buf, true);
        SB_NCCL(api().Send(buf, (size_t)n[0] * n[1] * n[2], kNcclFloat64, 0, comm, ctx->st));
        return;
    }
    std::vector<size_t> at(ctx->nranks, 0);
    size_t              o = 0;
    SB_NCCL(api().GroupStart());
    for (int r = 1; r < ctx->nranks; ++r) {
        tileExtent(dist.tiles[r], centering, n);
        at[r] = o;
        SB_NCCL(api().Recv(buf + o, (size_t)n[0] * n[1] * n[2], kNcclFloat64, r, comm, ctx->st));
        o += (size_t)n[0] * n[1] * n[2];
    }
    SB_NCCL(api().GroupEnd());
    for (int r = 0; r < ctx->nranks; ++r) {
        const Box3& t = dist.tiles[r];
        tileExtent(t, centering, n);
        const int off[3] = {t.lo[0] - full->lo0, t.lo[1] - full->lo1, t.lo[2] - full->lo2};
        if (r == 0) {  // own tile: straight from the tile array, through the head of the staging buffer
            double* own = buf + o;
            copyDense(ctx->st, dist.lay, const_cast<double*>(tileField), zero, n, own, true);
            copyDense(ctx->st, *full, fullField, off, n, own, false);
        } else {
            copyDense(ctx->st, *full, fullField, off, n, buf + at[r], false);
        }
    }
}

void Comm::scatterTiles(const Op& dist, double* tileField, const Lay* full, const double* fullField, double* buf)
{
    const int zero[3] = {0, 0, 0};
    int       n[3];
    if (ctx->rank != 0) {
        tileExtent(dist.tile, SB_CELL, n);
        SB_NCCL(api().Recv(buf, (size_t)n[0] * n[1] * n[2], kNcclFloat64, 0, comm, ctx->st));
        copyDense(ctx->st, dist.lay, tileField, zero, n, buf, false);
        return;
    }
    std::vector<size_t> at(ctx->nranks, 0);
    size_t              o = 0;
    for (int r = 0; r < ctx->nranks; ++r) {
        const Box3& t = dist.tiles[r];
        tileExtent(t, SB_CELL, n);
        const int off[3] = {t.lo[0] - full->lo0, t.lo[1] - full->lo1, t.lo[2] - full->lo2};
        at[r] = o;
        copyDense(ctx->st, *full, const_cast<double*>(fullField), off, n, buf + o, true);
        o += (size_t)n[0] * n[1] * n[2];
    }
    SB_NCCL(api().GroupStart());
    for (int r = 1; r < ctx->nranks; ++r) {
        tileExtent(dist.tiles[r], SB_CELL, n);
        SB_NCCL(api().Send(buf + at[r], (size_t)n[0] * n[1] * n[2], kNcclFloat64, r, comm, ctx->st));
    }
    SB_NCCL(api().GroupEnd());
    tileExtent(dist.tiles[0], SB_CELL, n);
    copyDense(ctx->st, dist.lay, tileField, zero, n, buf + at[0], false);
}

}  // namespace sb
