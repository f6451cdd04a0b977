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

void Comm::exchangeFacesSplit(Op& op, double* s0, double* s1)
{
    bool any = false;
    for (int d = 0; d < 2; ++d)
        for (int s = 0; s < 2; ++s)
            if (op.side[d][s].kind == SIDE_NEIGHBOR) { k::pack_face_split(ctx->st, op.slay, s0, s1, d, s, op.xbuf[d][s][0]); any = true; }
    if (!any) return;
    SB_NCCL(api().GroupStart());
    for (int d = 0; d < 2; ++d) {
        const size_t n = (size_t)(d == 0 ? op.lay.ny : op.lay.nx) * op.lay.nz;
        for (int s = 0; s < 2; ++s)
            if (op.side[d][s].kind == SIDE_NEIGHBOR)
                SB_NCCL(api().Send(op.xbuf[d][s][0], n, kNcclFloat64, op.side[d][s].neighbor, comm, ctx->st));
        for (int s = 1; s >= 0; --s)
            if (op.side[d][s].kind == SIDE_NEIGHBOR)
                SB_NCCL(api().Recv(op.xbuf[d][s][1], n, kNcclFloat64, op.side[d][s].neighbor, comm, ctx->st));
    }
    SB_NCCL(api().GroupEnd());
    for (int d = 0; d < 2; ++d)
        for (int s = 0; s < 2; ++s)
            if (op.side[d][s].kind == SIDE_NEIGHBOR) k::unpack_face_split(ctx->st, op.slay, s0, s1, d, s, op.xbuf[d][s][1]);
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

}  // namespace sb
