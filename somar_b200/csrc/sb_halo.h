// sb_halo.h -- face-ghost exchange between neighbouring tiles by stores into the neighbour's memory over NVLink
// (CUDA IPC mappings), for fields on colour-split storage.  Stands in for LevelData::exchange between the colours
// of a relaxation (PoissonOp.cpp:1957-1965, BoxTools/BoxLayoutDataI.H:665-812) without a collective launch: the
// producer writes its face layer straight into the ghost cells of the neighbour's arrays and bumps an arrival
// counter there; the consumer's stream spins on its own counter.  See sb_halo.cu for the protocol.
#pragma once
#include "sb_core.h"

namespace sb {

struct Op;
struct Context;

struct HaloPeerSide {
    double*             rs[2];   // the neighbour's two colour arrays, mapped into this process (null: side not exchanged)
    unsigned long long* rflag;   // the neighbour's arrival counter for the side that faces this tile
    long long           rsy, rsz;  // row / level strides of the neighbour's arrays
    int                 rx;      // dir 0: element of the ghost column in a row of the neighbour's arrays
    int                 rrow;    // dir 1: row of the ghost line in the neighbour's arrays
};
struct HaloDev {
    HaloPeerSide        side[4];  // [2 * dir + side]
    unsigned long long* flag;     // [4] arrival counters the neighbours write (in the exported block)
    unsigned long long* posted;   // [4] posts made per side
    unsigned long long* waited;   // [4] arrivals consumed per side
    unsigned int*       done;     // [4] CTA completion counters of halo_post_k
    int*                fault;    // mapped host memory (Context::fault)
};

// One exported block per split field: [colour 0 | colour 1 | flags].  The arrays are what the relaxation kernels
// use; the neighbours hold mappings of them.
struct PeerHalo {
    Context* ctx = nullptr;
    bool     ready = false;
    void*    block = nullptr;
    HaloDev  dev{};
    HaloDev* devCopy = nullptr;        // `dev` in device memory, for kernels that take it by pointer (vertline_tma_k)
    int      mask = 0;                 // bit 2 * dir + side: a neighbouring tile on that side
    std::vector<void*> opened;         // mapped bases of the neighbours' blocks (closed by the destructor)
    // allocates s0 / s1 (S.n doubles each) inside a block exported to the neighbours of `op` and maps theirs;
    // collective over the ranks.  Returns false (and leaves plain allocations) when a mapping could not be made on
    // some rank: every rank then keeps the NCCL exchange.
    bool setup(Op& op, const SLay& S, double** s0, double** s1);
    // post: write the face layers (cells of the colours in colourMask; 0: nothing, a bare "my ghosts may be
    // overwritten" signal) into the neighbours' ghost cells and bump their counters.  wait: block the stream until
    // every neighbour's next post has arrived.
    // gsrbPass >= 0: the cells point red-black pass `gsrbPass` has just updated (their array alternates with the level)
    void post(cudaStream_t st, const SLay& S, const double* s0, const double* s1, int colourMask, int gsrbPass = -1);
    void wait(cudaStream_t st);
    ~PeerHalo();
};

}  // namespace sb
