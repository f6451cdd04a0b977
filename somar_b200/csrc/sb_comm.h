// sb_comm.h -- NCCL plumbing for the horizontal box decomposition: face-ghost exchange between
// neighbouring tiles and scalar all-reduces.  Stands in for Chombo's MPI layer
// (BoxTools/BoxLayoutDataI.H:665-812 exchange, BaseTools/Comm.cpp:14-50 reduce).
// NCCL is bound at run time with dlopen so the library has no link-time dependency on it.
#pragma once
#include "sb_host.h"

namespace sb {

struct Comm {
    Context* ctx;
    void*    comm = nullptr;   // ncclComm_t
    double*  dscal = nullptr;  // device staging for scalar reductions
    Comm(Context* ctx, const void* id128);
    ~Comm();
    static void getUniqueId(void* id128);
    void allreduceHost(double* v, int n, bool isMax);
    void allreduceDevice(double* dev, int n, bool isMax, cudaStream_t st = nullptr);  // in place, stream-ordered, no host sync
    void exchangeFaces(Op& op, double* phi);
    void exchangeDir(Op& op, double* phi, int dir, int ext0, int ext1);
    void exchangeFacesSplit(Op& op, double* s0, double* s1, cudaStream_t st = nullptr, const SLay* S = nullptr);  // same, on colour-split storage (x and y sides); st: default ctx->st; S: default op.slay
    // Agglomeration: the tiles of `dist` (every rank) <-> one array over the whole domain on rank 0
    // (layout `full`, meaningful on rank 0 only).  buf: staging, sum over ranks of tile sizes on
    // rank 0, one tile elsewhere.  centering: SB_CELL or the face direction.
    void gatherTiles(const Op& dist, const double* tileField, const Lay* full, double* fullField, int centering, double* buf);
    void scatterTiles(const Op& dist, double* tileField, const Lay* full, const double* fullField, double* buf);
    // One NCCL group of point-to-point messages (the motion items of a Copier between ranks): every rank issues
    // its sends and receives in the same global order, so messages between a pair of ranks match up.
    struct Msg { double* p; size_t n; int peer; bool send; };
    void sendRecv(const std::vector<Msg>& msgs, cudaStream_t st = nullptr);
};

}  // namespace sb
