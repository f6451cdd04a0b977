// sb_op.cpp -- host driver of one MG depth: the B200 counterpart of Elliptic::PoissonOp
// (reference Grade3_Calculus/Elliptic/PoissonOp.cpp).  All field arithmetic is in
// sb_kernels.cu; this file holds geometry/setup logic and the call order of the reference.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "sb_comm.h"
#include "sb_host.h"

namespace sb {

// ---------------------------------------------------------------------------------------------
Context::Context(int dev, int rank_, int nranks_) : device(dev), rank(rank_), nranks(nranks_)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        SB_FAIL("no CUDA device visible: somar_b200 has no CPU fallback");
    SB_CUDA(cudaSetDevice(dev));
    SB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    hpinLen = 1 << 16;
    SB_CUDA(cudaMallocHost((void**)&hpin, hpinLen * sizeof(double)));
    SB_CUDA(cudaHostAlloc((void**)&fault, 16 * sizeof(int), cudaHostAllocMapped));
    std::memset(fault, 0, 16 * sizeof(int));
    launches0 = k::launch_count();
    { const char* e = getenv("SB_PEER_HALO"); peerHalo = !(e && std::string(e) == "0"); }
}
void Context::sync()
{
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) return;
    const int* f = parent ? parent->fault : fault;
    std::string why = std::string("cudaStreamSynchronize(st): ") + cudaGetErrorString(e);
    if (f && f[0]) why += " [kernel watchdog: code " + std::to_string(f[1]) + ", block " + std::to_string(f[2]) + ", thread " + std::to_string(f[3]) +
                          ", parity " + std::to_string(f[4]) + "]";
    SB_FAIL(why);
}
Context::Context(Context& p) : device(p.device), rank(0), nranks(1), st(p.st), parent(&p)
{
    hpinLen = 1 << 16;
    SB_CUDA(cudaMallocHost((void**)&hpin, hpinLen * sizeof(double)));
    launches0 = k::launch_count();
}
Context::~Context()
{
    delete comm;
    if (scratch) cudaFree(scratch);
    if (hpin) cudaFreeHost(hpin);
    if (fault) cudaFreeHost(fault);
    if (evEdge) cudaEventDestroy(evEdge);
    if (evHalo) cudaEventDestroy(evHalo);
    if (evPost) cudaEventDestroy(evPost);
    if (commSt) cudaStreamDestroy(commSt);
    for (cudaStream_t q : copySt) if (q) cudaStreamDestroy(q);
    if (st && !parent) cudaStreamDestroy(st);
}
cudaStream_t Context::stream(int which)
{
    if (which == SB_STREAM_COMPUTE) return st;
    if (which != SB_STREAM_H2D && which != SB_STREAM_D2H) SB_FAIL("bad stream id");
    cudaStream_t& q = copySt[which - 1];
    if (!q) SB_CUDA(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
    return q;
}
void Context::profBegin(const char* key, int depth, cudaEvent_t* e0)
{
    if (parent) { parent->profBegin(key, depth, e0); return; }
    *e0 = nullptr;
    if (!profiling) return;
    SB_CUDA(cudaEventCreate(e0));
    SB_CUDA(cudaEventRecord(*e0, st));
}
void Context::profEnd(const char* key, int depth, cudaEvent_t e0)
{
    if (parent) { parent->profEnd(key, depth, e0); return; }
    if (!profiling || !e0) return;
    cudaEvent_t e1;
    SB_CUDA(cudaEventCreate(&e1));
    SB_CUDA(cudaEventRecord(e1, st));
    prof[std::string(key) + "@" + std::to_string(depth)].ev.emplace_back(e0, e1);
}
void Context::phaseBegin(cudaEvent_t* e0)
{
    if (parent) { parent->phaseBegin(e0); return; }
    *e0 = nullptr;
    if (!phases) return;
    SB_CUDA(cudaEventCreate(e0));
    SB_CUDA(cudaEventRecord(*e0, st));
}
void Context::phaseEnd(const char* key, int depth, cudaEvent_t e0)
{
    if (parent) { parent->phaseEnd((std::string("agg.") + key).c_str(), depth, e0); return; }  // depths of the agglomerated hierarchy count from its top
    if (!phases || !e0) return;
    cudaEvent_t e1;
    SB_CUDA(cudaEventCreate(&e1));
    SB_CUDA(cudaEventRecord(e1, st));
    prof[std::string(key) + "@" + std::to_string(depth)].ev.emplace_back(e0, e1);
}
void Context::profResolve()
{
    sync();
    for (auto& kv : prof) {
        for (auto& pr : kv.second.ev) {
            float ms = 0;
            cudaEventElapsedTime(&ms, pr.first, pr.second);
            kv.second.ms += ms;
            kv.second.count += 1;
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        kv.second.ev.clear();
    }
}
void* Context::getScratch(size_t bytes)
{
    if (bytes > scratchBytes) {
        sync();
        if (scratch) SB_CUDA(cudaFree(scratch));
        SB_CUDA(cudaMalloc(&scratch, bytes));
        scratchBytes = bytes;
    }
    return scratch;
}
void Context::allreduceSum(double* v, int n)
{
    if (nranks > 1) { if (!comm) SB_FAIL("nranks > 1 but sb_comm_init was not called"); comm->allreduceHost(v, n, false); }
}
void Context::allreduceMax(double* v, int n)
{
    if (nranks > 1) { if (!comm) SB_FAIL("nranks > 1 but sb_comm_init was not called"); comm->allreduceHost(v, n, true); }
}

// ---------------------------------------------------------------------------------------------
// Geometry source.
static const double kPi = 3.14159265358979323846264338327950288e0;  // Grade1_Basics/SOMAR_Constants.H:55

void MapSpec::interp(std::vector<double>& x, const std::vector<double>& xi, int mu) const
{
    x.resize(xi.size());
    if (kind == SB_MAP_CARTESIAN) {
        x = xi;
    } else if (kind == SB_MAP_STRETCHED) {
        // maps/StretchedMap.cpp:9-38
        const double A = ampl[mu], kk = 2.0 * kPi / (xmax[mu] - xmin[mu]), x0 = xmin[mu];
        for (size_t n = 0; n < xi.size(); ++n) x[n] = xi[n] + A * std::sin(kk * (xi[n] - x0));
    } else {
        if (!fn) SB_FAIL("SB_MAP_CALLBACK without map_fn");
        fn(x.data(), xi.data(), (int)xi.size(), mu, user);
    }
}
std::vector<double> MapSpec::physCoor(int mu, double dXi_, int lo, int n, int nodeType) const
{
    std::vector<double> x(n);
    if (kind == SB_MAP_CARTESIAN) {
        // maps/CartesianMapF.ChF:31-80: x = dXi*(i + offset)
        const double offset = (1.0 - nodeType) * 0.5;
        for (int m = 0; m < n; ++m) x[m] = dXi_ * ((double)(lo + m) + offset);
        return x;
    }
    // GeoSourceInterface.cpp:41-73
    std::vector<double> vxi(n);
    double              xi = lo * dXi_;
    if (nodeType == 0) xi += 0.5 * dXi_;
    for (int m = 0; m < n; ++m) { vxi[m] = xi; xi += dXi_; }
    interp(x, vxi, mu);
    return x;
}
std::vector<double> MapSpec::dxdXi(int mu, double dXi_, int lo, int n, int nodeType, double scale) const
{
    std::vector<double> d(n);
    if (kind == SB_MAP_CARTESIAN) {  // maps/CartesianMap.cpp:117-130: setVal(scale)
        std::fill(d.begin(), d.end(), scale);
        return d;
    }
    const double scaledDXi = dXi_ / scale;
    const double oneOnDx   = 1.0 / scaledDXi;
    if (nodeType == 0) {
        // cell-centred dest: nodes lo..lo+n, FINITEDIFF_PARTIALD_NC2CC
        std::vector<double> x = physCoor(mu, dXi_, lo, n + 1, 1);
        for (int m = 0; m < n; ++m) d[m] = (x[m + 1] - x[m]) * oneOnDx;
    } else {
        // node-centred dest: cells lo-1..lo+n-1, FINITEDIFF_PARTIALD_CC2NC
        std::vector<double> x = physCoor(mu, dXi_, lo - 1, n + 1, 0);
        for (int m = 0; m < n; ++m) d[m] = (x[m + 1] - x[m]) * oneOnDx;
    }
    return d;
}

// ---------------------------------------------------------------------------------------------
Field::Field(Op* op_, int c) : op(op_), centering(c)
{
    d = op->alloc();
    SB_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    SB_CUDA(cudaEventRecord(ready, op->ctx->st));
}
Field::~Field()
{
    if (ready) cudaEventDestroy(ready);
    if (d) cudaFree(d);
}

double* Op::alloc() const
{
    double* p = nullptr;
    SB_CUDA(cudaMalloc((void**)&p, lay.n * sizeof(double)));
    SB_CUDA(cudaMemsetAsync(p, 0, lay.n * sizeof(double), ctx->st));
    return p;
}

// ---------------------------------------------------------------------------------------------
Op::Op(Context* c, const sb_level_desc& d) : ctx(c)
{
    dim = d.dim;
    if (dim != 2 && dim != 3) SB_FAIL("dim must be 2 or 3");
    for (int i = 0; i < 3; ++i) {
        domain.lo[i] = d.domain_lo[i]; domain.hi[i] = d.domain_hi[i];
        periodic[i] = d.periodic[i]; dXi[i] = d.dXi[i];
        for (int s = 0; s < 2; ++s) { bcAlpha[i][s] = d.bc_alpha[i][s]; bcBeta[i][s] = d.bc_beta[i][s]; }
        map.xmin[i] = d.map_xmin[i]; map.xmax[i] = d.map_xmax[i]; map.ampl[i] = d.map_ampl[i];
    }
    if (dim == 2 && domain.size(1) != 1) SB_FAIL("dim == 2 requires a single layer in slot 1");
    map.kind = d.map_kind; map.fn = d.map_fn; map.user = d.map_user;
    alpha = d.alpha; beta = d.beta; relaxMethod = d.relax_method;
    { const char* gk = getenv("SB_GSRB_KERNEL"); gsrbNatural = gk && std::string(gk) == "natural"; }  // test knob, read at creation
    { const char* tk = getenv("SB_LINE_TMA"); lineTmaAllowed = !(tk && std::string(tk) == "0"); lineTmaForce = tk && std::string(tk) == "force"; }  // "0": keep vertline_fused_k, "force": no size threshold (tests)
    if (d.num_boxes <= 0) SB_FAIL("empty box list");
    boxes.resize(d.num_boxes); boxRank.resize(d.num_boxes);
    for (int b = 0; b < d.num_boxes; ++b) {
        for (int i = 0; i < 3; ++i) { boxes[b].lo[i] = d.box_lo[3 * b + i]; boxes[b].hi[i] = d.box_hi[3 * b + i]; }
        boxRank[b] = d.box_rank ? d.box_rank[b] : 0;
        if (boxRank[b] < 0 || boxRank[b] >= ctx->nranks) SB_FAIL("box_rank out of range");
    }
    // AMR stuff (PoissonOp.cpp:83-93, 118-120)
    if (d.num_crse_boxes > 0) {
        if (!d.crse_box_lo || !d.crse_box_hi) SB_FAIL("num_crse_boxes > 0 without crse_box_lo / crse_box_hi");
        refined = true;
        crseGrids.boxes.resize(d.num_crse_boxes);
        crseGrids.rank.resize(d.num_crse_boxes);
        for (int b = 0; b < d.num_crse_boxes; ++b) {
            for (int i = 0; i < 3; ++i) { crseGrids.boxes[b].lo[i] = d.crse_box_lo[3 * b + i]; crseGrids.boxes[b].hi[i] = d.crse_box_hi[3 * b + i]; }
            crseGrids.rank[b] = d.crse_box_rank ? d.crse_box_rank[b] : 0;
            if (crseGrids.rank[b] < 0 || crseGrids.rank[b] >= ctx->nranks) SB_FAIL("crse_box_rank out of range");
        }
        for (int i = 0; i < 3; ++i) {
            crseGrids.domain.lo[i] = d.crse_domain_lo[i]; crseGrids.domain.hi[i] = d.crse_domain_hi[i];
            // calculateRefinementRatio(crseDomBox, domBox)
            const int nc = crseGrids.domain.size(i), nf = domain.size(i);
            if (nc <= 0 || nf % nc != 0) SB_FAIL("the domain is not a refinement of the coarser AMR level's domain");
            crseRef[i] = nf / nc;
            if (crseGrids.domain.lo[i] * crseRef[i] != domain.lo[i]) SB_FAIL("the domain is not a refinement of the coarser AMR level's domain");
            amrCrseDXi[i] = dXi[i] * (double)crseRef[i];
        }
    }
    setupLayout();
    J = alloc();
    for (int i = 0; i < 3; ++i) Jgup[i] = alloc();
    fillMetricFromMap();
}

// Host-only decomposition logic (no CUDA): the rectangle each rank owns and what each of its six
// sides touches.  Stands in for what DisjointBoxLayout + Copier::exchangeDefine work out per box
// (BoxTools/Copier.cpp:784); here ranks own rectangles of boxes, so it is per tile.
void planDecomposition(const std::vector<Box3>& boxes, const std::vector<int>& boxRank, const Box3& domain,
                       const int periodic[3], int rank, int nr, std::vector<Box3>& tiles, std::vector<int>& local,
                       SideBC side[3][2], bool partial)
{
    tiles.assign(nr, Box3{{0, 0, 0}, {-1, -1, -1}});
    std::vector<long long> pts(nr, 0);
    std::vector<int>       cnt(nr, 0);
    local.clear();
    for (size_t b = 0; b < boxes.size(); ++b) {
        const int r = boxRank[b];
        if (r < 0 || r >= nr) SB_FAIL("box_rank out of range");
        Box3&     t = tiles[r];
        if (cnt[r]++ == 0) t = boxes[b];
        else
            for (int i = 0; i < 3; ++i) { t.lo[i] = std::min(t.lo[i], boxes[b].lo[i]); t.hi[i] = std::max(t.hi[i], boxes[b].hi[i]); }
        pts[r] += boxes[b].numPts();
        if (r == rank) local.push_back((int)b);
    }
    long long total = 0;
    for (int r = 0; r < nr; ++r) {
        if (cnt[r] == 0) SB_FAIL("rank " + std::to_string(r) + " owns no box");
        if (pts[r] != tiles[r].numPts()) SB_FAIL("the boxes of one rank must tile a rectangle (horizontal box decomposition)");
        total += pts[r];
    }
    if (!partial) {
        if (total != domain.numPts()) SB_FAIL("boxes do not cover the domain (single-level operator)");
    } else {
        Box3 patch = tiles[0];
        for (int r = 1; r < nr; ++r)
            for (int i = 0; i < 3; ++i) { patch.lo[i] = std::min(patch.lo[i], tiles[r].lo[i]); patch.hi[i] = std::max(patch.hi[i], tiles[r].hi[i]); }
        if (total != patch.numPts()) SB_FAIL("the boxes of a refined AMR level must form one rectangular patch");
        for (int i = 0; i < 3; ++i)
            if (patch.lo[i] < domain.lo[i] || patch.hi[i] > domain.hi[i]) SB_FAIL("refined patch outside the domain");
    }
    const Box3& tile = tiles[rank];

    // What does each side of the tile touch?
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s) {
            SideBC& sd = side[d][s];
            sd = SideBC{SIDE_PHYS, 1, 0.0, 0.0, -1};
            const bool atDom = s ? tile.hi[d] == domain.hi[d] : tile.lo[d] == domain.lo[d];
            if (atDom && !periodic[d]) continue;
            if (atDom && periodic[d] && tile.lo[d] == domain.lo[d] && tile.hi[d] == domain.hi[d]) { sd.kind = SIDE_PERIODIC_SELF; continue; }
            // neighbour tile: same extents in the other directions, adjacent (or wrapped) in d
            int want = s ? tile.hi[d] + 1 : tile.lo[d] - 1;
            if (atDom) want = s ? domain.lo[d] : domain.hi[d];
            int found = -1;
            for (int r = 0; r < nr && found < 0; ++r) {
                if (r == rank) continue;
                const Box3& t = tiles[r];
                bool ok = s ? t.lo[d] == want : t.hi[d] == want;
                for (int o = 0; o < 3 && ok; ++o)
                    if (o != d && (t.lo[o] != tile.lo[o] || t.hi[o] != tile.hi[o])) ok = false;
                if (ok) found = r;
            }
            if (found < 0) {
                if (!partial) SB_FAIL("rank tiles do not form a process grid (no neighbour across a tile side)");
                // no tile of this level across the side: the coarser AMR level is there (a patch that reaches a
                // periodic boundary without spanning the domain borders coarse cells through the wrap as well)
                sd.kind = SIDE_CF;
                continue;
            }
            sd.kind = SIDE_NEIGHBOR; sd.neighbor = found;
        }
}

void Op::setupLayout()
{
    planDecomposition(boxes, boxRank, domain, periodic, ctx->rank, ctx->nranks, tiles, local, side, refined);
    tile = tiles[ctx->rank];
    lay  = makeLay(tile);
    patch = tiles[0];
    for (const Box3& t : tiles)
        for (int i = 0; i < 3; ++i) { patch.lo[i] = std::min(patch.lo[i], t.lo[i]); patch.hi[i] = std::max(patch.hi[i], t.hi[i]); }
    if (flatZ) side[2][0].kind = side[2][1].kind = -1;  // inactive direction: no BCs, no exchange (activeSides, PoissonOp.cpp:483-487)

    // device-side box list (tile-local indices) and reduction buffers
    const int nl = nlocal();
    std::vector<int> lh(6 * nl);
    for (int b = 0; b < nl; ++b)
        for (int i = 0; i < 3; ++i) {
            lh[3 * b + i]          = boxes[local[b]].lo[i] - tile.lo[i];
            lh[3 * nl + 3 * b + i] = boxes[local[b]].hi[i] - tile.lo[i];
        }
    SB_CUDA(cudaMalloc((void**)&boxLoHi, lh.size() * sizeof(int)));
    SB_CUDA(cudaMemcpy(boxLoHi, lh.data(), lh.size() * sizeof(int), cudaMemcpyHostToDevice));
    SB_CUDA(cudaMalloc((void**)&redPartial, k::reduce_partial_len(nl) * sizeof(double)));
    SB_CUDA(cudaMalloc((void**)&redOut, 2 * nl * sizeof(double)));
    SB_CUDA(cudaMalloc((void**)&shiftBuf, 2 * sizeof(double)));
    SB_CUDA(cudaMalloc((void**)&pivotFlag, sizeof(int)));
    SB_CUDA(cudaMemset(pivotFlag, 0, sizeof(int)));
    if ((size_t)2 * nl + 16 > ctx->hpinLen) SB_FAIL("too many boxes per rank for the pinned scalar buffer");
    SB_CUDA(cudaMalloc((void**)&mtab, 2 * (size_t)(lay.nx + lay.ny + lay.nz) * sizeof(double)));
    SB_CUDA(cudaMalloc((void**)&loBC, (size_t)lay.sz * sizeof(double)));
    SB_CUDA(cudaMalloc((void**)&hiBC, (size_t)lay.sz * sizeof(double)));
    SB_CUDA(cudaMemset(loBC, 0, (size_t)lay.sz * sizeof(double)));
    SB_CUDA(cudaMemset(hiBC, 0, (size_t)lay.sz * sizeof(double)));
    // exchange buffers
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s)
            if (side[d][s].kind == SIDE_NEIGHBOR) {
                const size_t n = k::face_count(lay, d, 1, 1);  // room for the edge-extended exchange
                for (int w = 0; w < 2; ++w) SB_CUDA(cudaMalloc((void**)&xbuf[d][s][w], n * sizeof(double)));
            }
}

Op::~Op()
{
    cudaFree(J); cudaFree(Dinv);
    for (int i = 0; i < 3; ++i) cudaFree(Jgup[i]);
    cudaFree(lineTab); cudaFree(lineTabS); cudaFree(lineTabG); cudaFree(gstart); cudaFree(colTab);
    for (int q = 0; q < 4; ++q) if (q >= 2 || !haloLine) cudaFree(sp[q]);   // sp[0], sp[1] live in haloLine's block
    for (int q = 0; q < 8; ++q) if (q >= 2 || !haloGsrb) cudaFree(sg[q]);
    for (auto& kv : relaxGraphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    cudaFree(mtab); cudaFree(loBC); cudaFree(hiBC); cudaFree(boxLoHi); cudaFree(redPartial); cudaFree(redOut); cudaFree(shiftBuf); cudaFree(pivotFlag);
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s)
            for (int w = 0; w < 2; ++w) cudaFree(xbuf[d][s][w]);
}

Coef Op::coef() const
{
    Coef c;
    c.J = J; c.Dinv = Dinv;
    c.mxl = mtab; c.mxr = mtab + lay.nx;
    c.myl = mtab + 2 * lay.nx; c.myr = c.myl + lay.ny;
    c.mzl = mtab + 2 * (lay.nx + lay.ny); c.mzr = c.mzl + lay.nz;
    c.loBC = loBC; c.hiBC = hiBC; c.beta = beta;
    c.tabJ = coefUniform ? colTab : nullptr;
    c.tabD = coefUniform ? colTab + lay.nz : nullptr;
    return c;
}
BoxList Op::boxlist() const
{
    int w = 1;
    for (int lb : local) w = std::max(w, boxes[lb].size(0));
    return BoxList{nlocal(), boxLoHi, boxLoHi + 3 * nlocal(), w};
}

// LevelGeometry::createMetricCache (LevelGeometry.cpp:238-277): per box, FABs grown by 4 ghosts,
// each filled by GeoSourceInterface::fill_J / fill_Jgup, i.e. from 1-D dx/dXi tables whose xi is
// accumulated from the grown box's small end.
void Op::fillMetricFromMap()
{
    const int G = 4;
    for (int lb = 0; lb < nlocal(); ++lb) {
        const Box3& b = boxes[local[lb]];
        std::vector<double> tab;
        size_t off[6];
        for (int mu = 0; mu < 3; ++mu) {
            const int n = b.size(mu);
            std::vector<double> c, f;
            if (dim == 2 && mu == 1) { c.assign(n, 1.0); f.assign(n + 1, 1.0); }
            else {
                std::vector<double> cg = map.dxdXi(mu, dXi[mu], b.lo[mu] - G, n + 2 * G, 0);
                std::vector<double> fg = map.dxdXi(mu, dXi[mu], b.lo[mu] - G, n + 2 * G + 1, 1);
                c.assign(cg.begin() + G, cg.begin() + G + n);
                f.assign(fg.begin() + G, fg.begin() + G + n + 1);
            }
            off[mu] = tab.size(); tab.insert(tab.end(), c.begin(), c.end());
            off[3 + mu] = tab.size(); tab.insert(tab.end(), f.begin(), f.end());
        }
        double* dt = (double*)ctx->getScratch(tab.size() * sizeof(double));
        SB_CUDA(cudaMemcpyAsync(dt, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
        int blo[3], bhi[3];
        for (int i = 0; i < 3; ++i) { blo[i] = b.lo[i] - tile.lo[i]; bhi[i] = b.hi[i] - tile.lo[i]; }
        k::fill_metric_box(st(), lay, blo, bhi, dt + off[0], dt + off[1], dt + off[2], dt + off[3], dt + off[4], dt + off[5], J,
                           Jgup[0], Jgup[1], Jgup[2]);
        ctx->sync();
    }
}

// Coarsened operator (PoissonOp.cpp:334-405): grids, J (and Jgup) block-averaged from the finer
// depth, matrix elements recomputed from the map at the coarse dXi.
Op::Op(const Op& f, const int ref[3]) : ctx(f.ctx)
{
    dim = f.dim; alpha = f.alpha; beta = f.beta; relaxMethod = f.relaxMethod; map = f.map; depth = f.depth + 1; flatZ = f.flatZ;
    gsrbNatural = f.gsrbNatural; lineTmaAllowed = f.lineTmaAllowed; lineTmaForce = f.lineTmaForce;
    refined = f.refined;  // the coarse-fine sides stay (coarsened CFRegion, PoissonOp.cpp:392-396), with homogeneous ghosts only
    std::memcpy(amrCrseDXi, f.amrCrseDXi, sizeof(amrCrseDXi));  // not scaled (PoissonOp.cpp:345)
    std::memcpy(periodic, f.periodic, sizeof(periodic));
    std::memcpy(bcAlpha, f.bcAlpha, sizeof(bcAlpha));
    std::memcpy(bcBeta, f.bcBeta, sizeof(bcBeta));
    for (int d = 0; d < 3; ++d) dXi[d] = f.dXi[d] * (double)ref[d];
    domain  = coarsen(f.domain, ref);
    boxRank = f.boxRank;
    boxes.resize(f.boxes.size());
    for (size_t b = 0; b < boxes.size(); ++b) {
        if (!coarsenable(f.boxes[b], ref)) SB_FAIL("Grids cannot be coarsened by the requested ref ratio");
        boxes[b] = coarsen(f.boxes[b], ref);
    }
    setupLayout();
    J = alloc();
    k::restrict_avg(st(), f.lay, lay, ref, J, f.J);
    for (int d = 0; d < 3; ++d) {
        Jgup[d] = alloc();
        if (dim == 2 && d == 1) continue;
        k::restrict_face(st(), f.lay, lay, ref, d, Jgup[d], f.Jgup[d]);
    }
    hasNullSpace = f.hasNullSpace;
    cacheMatrixElements();
    finalized = true;
}

// Agglomerated copy of a distributed depth: same geometry and boxes, all owned by the single rank
// of `single`.  The caller gathers J and Jgup from the tiles (they were block-averaged from the
// finer depth, PoissonOp.cpp:368-395, so they cannot be rebuilt from the map) and then calls
// cacheMatrixElements().
Op::Op(Context* single, const Op& f) : ctx(single)
{
    dim = f.dim; alpha = f.alpha; beta = f.beta; relaxMethod = f.relaxMethod; map = f.map; depth = f.depth; flatZ = f.flatZ;
    refined = f.refined; gsrbNatural = f.gsrbNatural; lineTmaAllowed = f.lineTmaAllowed; lineTmaForce = f.lineTmaForce;
    std::memcpy(amrCrseDXi, f.amrCrseDXi, sizeof(amrCrseDXi));
    std::memcpy(periodic, f.periodic, sizeof(periodic));
    std::memcpy(bcAlpha, f.bcAlpha, sizeof(bcAlpha));
    std::memcpy(bcBeta, f.bcBeta, sizeof(bcBeta));
    std::memcpy(dXi, f.dXi, sizeof(dXi));
    domain = f.domain;
    boxes  = f.boxes;
    boxRank.assign(boxes.size(), 0);
    setupLayout();
    J = alloc();
    for (int d = 0; d < 3; ++d) Jgup[d] = alloc();
    hasNullSpace = f.hasNullSpace;
}

// PoissonOp::cacheMatrixElements (PoissonOp.cpp:510-665).
void Op::cacheMatrixElements()
{
    std::vector<double> h(2 * (size_t)(lay.nx + lay.ny + lay.nz), 0.0);
    size_t              off = 0;
    for (int d = 0; d < 3; ++d) {
        const int N = domain.size(d);
        hM[d].assign(2 * (size_t)N, 0.0);
        if (!(dim == 2 && d == 1) && !(flatZ && d == 2)) {  // inactive directions: m_M[d].setVal(0) (PoissonOp.cpp:535-537)
            // fill_dXidx = 1 / fill_dxdXi (GeoSourceInterface.cpp:228-243) over the flattened domain box
            std::vector<double> fc = map.dxdXi(d, dXi[d], domain.lo[d], N + 1, 1);
            std::vector<double> cc = map.dxdXi(d, dXi[d], domain.lo[d], N, 0);
            for (auto& v : fc) v = 1.0 / v;
            for (auto& v : cc) v = 1.0 / v;
            // PoissonOpF.ChF:78-100
            const double dxScale = 1.0 / (dXi[d] * dXi[d]);
            for (int i = 0; i < N; ++i) {
                const double c = dxScale * cc[i];
                hM[d][i]       = c * fc[i];
                hM[d][N + i]   = c * fc[i + 1];
            }
        }
        const int n = d == 0 ? lay.nx : (d == 1 ? lay.ny : lay.nz);
        const int o = tile.lo[d] - domain.lo[d];
        for (int i = 0; i < n; ++i) { h[off + i] = hM[d][o + i]; h[off + n + i] = hM[d][N + o + i]; }
        off += 2 * (size_t)n;
    }
    ctx->sync();
    SB_CUDA(cudaMemcpy(mtab, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));

    if (!Dinv) Dinv = alloc();
    coefUniform = false;
    k::compute_dinv(st(), lay, coef(), alpha, Dinv, dim);
    gsrbCoefSplit = false;
    detectColumnCoefficients();

    // Physical-boundary ghost-fill constants (BCTools.cpp:466-508: dx = dx/dXi * dXi at the
    // boundary face; BCToolsF.ChF:222-337).
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s) {
            SideBC& sd = side[d][s];
            if (sd.kind == SIDE_CF && !(dim == 2 && d == 1) && !(flatZ && d == 2)) {
                // CFInterp::homogInterpAtCFI(phi, m_dXi, m_amrCrseDXi, m_cfiIter) (PoissonOp.cpp:745, CFInterp.cpp:429-515):
                // quadratic through the ghost's two interior neighbours and a zero coarse value, linear for one-cell boxes
                int minSize = std::numeric_limits<int>::max();
                for (int lb : local) {
                    const Box3& b = boxes[lb];
                    if ((s ? b.hi[d] == tile.hi[d] : b.lo[d] == tile.lo[d])) minSize = std::min(minSize, b.size(d));
                }
                const double dxf = dXi[d], dxc = amrCrseDXi[d];
                if (!(dxc > 0.0)) SB_FAIL("coarse-fine side without a coarser AMR level");
                sd.twoCells = minSize > 1;
                if (sd.twoCells) { sd.a = 2.0 * (dxc - dxf) / (dxc + dxf); sd.bb = -(dxc - dxf) / (dxc + 3.0 * dxf); }
                else { sd.a = 1.0 - 2.0 * dxf / (dxf + dxc); sd.bb = 0.0; }
                continue;
            }
            if (sd.kind != SIDE_PHYS || (dim == 2 && d == 1) || (flatZ && d == 2)) continue;
            const int face = s ? domain.hi[d] + 1 : domain.lo[d];
            const double dx = map.dxdXi(d, dXi[d], face, 1, 1, dXi[d])[0];
            int minSize = std::numeric_limits<int>::max();
            for (int lb : local) {
                const Box3& b = boxes[lb];
                if ((s ? b.hi[d] == domain.hi[d] : b.lo[d] == domain.lo[d])) minSize = std::min(minSize, b.size(d));
            }
            sd.twoCells = minSize >= 2;
            sd.a  = sd.twoCells ? bcAlpha[d][s] / 8.0 : bcAlpha[d][s] / 2.0;
            sd.bb = bcBeta[d][s] / dx;
        }

    if (relaxMethod == SB_RELAX_VERTLINE) {
        if (tile.lo[2] != domain.lo[2] || tile.hi[2] != domain.hi[2])
            SB_FAIL("Grids are not suitable for vertical line relaxation. Try setting base.splitDirs = 1 1 0 in 3D or 1 0 in 2D.");
        for (const Box3& b : boxes)
            if (b.lo[2] != domain.lo[2] || b.hi[2] != domain.hi[2])
                SB_FAIL("Grids are not suitable for vertical line relaxation. Try setting base.splitDirs = 1 1 0 in 3D or 1 0 in 2D.");
        // PoissonOpF.ChF:676-688
        const double dz  = dXi[2];
        const double sLo = (bcAlpha[2][0] - 2.0 * bcBeta[2][0] / dz) / (bcAlpha[2][0] + 2.0 * bcBeta[2][0] / dz);
        const double sHi = (bcAlpha[2][1] - 2.0 * bcBeta[2][1] / dz) / (bcAlpha[2][1] + 2.0 * bcBeta[2][1] / dz);
        if (periodic[2]) SB_FAIL("vertical line relaxation with a periodic vertical is not supported by the reference either");
        k::compute_vert_bcs(st(), lay, coef(), sLo, sHi, loBC, hiBC);
        buildLineTables(sLo, sHi);
    }
}

// Are J and Dinv functions of the level only on this tile, BIT FOR BIT?  (True on every grid that is Cartesian in the
// horizontal, vertically stretched ones included.)  Then the kernels that would stream the two arrays -- the residual,
// preCond, the J-weighted sum of removeKernel -- read two [nz] tables holding the same doubles instead: the same
// arithmetic on the same operands, 16 bytes per cell less traffic.
void Op::detectColumnCoefficients()
{
    const int N = lay.nz;
    if (!colTab) SB_CUDA(cudaMalloc((void**)&colTab, 2 * (size_t)N * sizeof(double)));
    SB_CUDA(cudaMemcpy2DAsync(colTab, sizeof(double), J + lay.idx(0, 0, 0), (size_t)lay.sz * sizeof(double), sizeof(double), N,
                              cudaMemcpyDeviceToDevice, ctx->st));
    SB_CUDA(cudaMemcpy2DAsync(colTab + N, sizeof(double), Dinv + lay.idx(0, 0, 0), (size_t)lay.sz * sizeof(double), sizeof(double), N,
                              cudaMemcpyDeviceToDevice, ctx->st));
    double dev[2] = {1.0, 1.0};
    k::j_deviation(st(), lay, J, colTab, redOut);
    SB_CUDA(cudaMemcpyAsync(&dev[0], redOut, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    ctx->sync();
    k::j_deviation(st(), lay, Dinv, colTab + N, redOut);
    SB_CUDA(cudaMemcpyAsync(&dev[1], redOut, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    ctx->sync();
    static const bool allowed = [] { const char* e = getenv("SB_COEF_TABLES"); return !(e && std::string(e) == "0"); }();
    coefUniform = allowed && dev[0] == 0.0 && dev[1] == 0.0;
}

// Fast path of the line relaxation (vertline_smem_k): usable when every column of this depth has
// the same tridiagonal matrix after dividing each row by beta*J, i.e. when the horizontal metric
// is uniform.  Decided from the data, not from the map kind: the 1-D off-diagonal tables must be
// constant in x and y and J must be a function of z only (to 1e-13).  Builds the Thomas
// factorisation of that matrix once (the dgtsv no-interchange recurrence, PoissonOpF.ChF:905-1002).
void Op::buildLineTables(double sLo, double sHi)
{
    // captured pass loops hold pointers to the tables rebuilt below
    for (auto& kv : relaxGraphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    relaxGraphs.clear();
    lineFast  = false;
    lineSplit = false;
    lineTma = lineGeneral = false;
    lineSLo = sLo; lineSHi = sHi;
    const char* force = getenv("SB_LINE_KERNEL");  // "general" keeps vertline_k, "mapped" forces the per-column kernel (tests)
    if (force && std::string(force) == "general") return;
    const int N = lay.nz;
    if (beta == 0.0) return;
    if (force && std::string(force) == "mapped") { buildGeneralLine(sLo, sHi); return; }
    if (k::vertline_smem_bytes(N) > 200 * 1024) { buildGeneralLine(sLo, sHi); return; }
    // h = MxL+MxR+MyL+MyR constant over the whole domain?
    double hmin = 1e300, hmax = -1e300;
    for (int d = 0; d < 2; ++d) {
        if (dim == 2 && d == 1) continue;
        const int Nd = domain.size(d);
        double lo = 1e300, hi = -1e300;
        for (int i = 0; i < Nd; ++i) {
            const double v = hM[d][i] + hM[d][Nd + i];
            lo = std::min(lo, v); hi = std::max(hi, v);
        }
        if (hi - lo > 1e-13 * std::abs(hi)) { buildGeneralLine(sLo, sHi); return; }
        (void)hmin; (void)hmax;
    }
    double h = hM[0][0] + hM[0][domain.size(0)];
    if (dim == 3) h += hM[1][0] + hM[1][domain.size(1)];
    // J(i,j,k) == J(0,0,k)?
    std::vector<double> jcol(N);
    ctx->sync();
    SB_CUDA(cudaMemcpy2D(jcol.data(), sizeof(double), J + lay.idx(0, 0, 0), (size_t)lay.sz * sizeof(double), sizeof(double), N,
                         cudaMemcpyDeviceToHost));
    if (!lineTab) SB_CUDA(cudaMalloc((void**)&lineTab, 4 * (size_t)N * sizeof(double)));
    SB_CUDA(cudaMemcpy(lineTab, jcol.data(), N * sizeof(double), cudaMemcpyHostToDevice));
    k::j_deviation(st(), lay, J, lineTab, redOut);
    double dev = 0.0;
    SB_CUDA(cudaMemcpyAsync(&dev, redOut, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    ctx->sync();
    ctx->allreduceMax(&dev, 1);
    if (!(dev <= 1e-13)) { buildGeneralLine(sLo, sHi); return; }
    if (ctx->nranks > 1) {  // every rank must use the same column of J: take rank 0's via max (values agree to 1e-13)
        ctx->allreduceMax(jcol.data(), std::min(N, 64));
        if (N > 64) for (int o = 64; o < N; o += 64) ctx->allreduceMax(jcol.data() + o, std::min(64, N - o));
    }
    const int           Nz = domain.size(2);
    const double*       mzl = hM[2].data();
    const double*       mzr = hM[2].data() + Nz;
    std::vector<double> t(4 * (size_t)N);
    double              dprev = 0.0;
    for (int k = 0; k < N; ++k) {
        double diag = alpha / beta - h - mzl[k] - mzr[k];
        if (k == 0) diag += -mzl[0] * sLo;
        if (k == N - 1) diag += -mzr[N - 1] * sHi;
        double d = diag;
        if (k > 0) {
            if (!(std::abs(dprev) >= std::abs(mzl[k])) || dprev == 0.0) return;  // dgtsv would interchange rows
            const double f = mzl[k] / dprev;
            t[N + (k - 1)] = -f;
            d              = diag - f * mzr[k - 1];
        }
        if (d == 0.0) return;
        t[k]         = 1.0 / (beta * jcol[k]);
        t[2 * N + k] = 1.0 / d;
        t[3 * N + k] = -(mzr[k] * (1.0 / d));
        dprev        = d;
    }
    t[N + (N - 1)] = 0.0;
    SB_CUDA(cudaMemcpy(lineTab, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
    lineFast = true;

    // Tables of the chunked sweeps on colour-split storage (sb_line.cu): s, a, P, g, c, Q.
    lineSplit = false;
    if (force && std::string(force) == "smem") return;
    if (!k::vertline_split_fits(N)) return;
    const int           CL = k::vertline_split_chunk(N);
    std::vector<double> u;
    if (!k::vertline_split_fused()) {
        u.assign(6 * (size_t)N, 0.0);
        for (int k = 0; k < N; ++k) {
            u[k]         = t[k];
            u[N + k]     = k > 0 ? t[N + (k - 1)] : 0.0;
            u[3 * N + k] = t[2 * N + k];
            u[4 * N + k] = k < N - 1 ? t[3 * N + k] : 0.0;
        }
        for (int k0 = 0; k0 < N; k0 += CL) {
            const int k1 = std::min(N, k0 + CL);
            double    p  = 1.0;
            for (int k = k0; k < k1; ++k) { p *= u[N + k]; u[2 * N + k] = p; }
            p = 1.0;
            for (int k = k1 - 1; k >= k0; --k) { p *= u[4 * N + k]; u[5 * N + k] = p; }
        }
    } else {
        // vertline_fused_k: {a', g}[N] | {P', c}[N] | R[N] | Pend[NW] | T[NW] | Rend[NW]  (sb_line.cu)
        const int NW = k::vertline_split_nw();
        u.assign(5 * (size_t)N + 3 * (size_t)NW, 0.0);
        for (int k = 0; k < N; ++k) {
            const double g = t[2 * N + k];                          // 1 / d'_k
            u[2 * k]     = k > 0 ? -(mzl[k] * g) : 0.0;             // a'_k
            u[2 * k + 1] = g;
            u[2 * N + 2 * k + 1] = k < N - 1 ? t[3 * N + k] : 0.0;  // c_k = -MzR_k g_k
        }
        for (int v = 0; v < NW; ++v) {
            const int k0 = v * CL, k1 = std::min(N, k0 + CL);
            double    P = 1.0, R = 1.0, T = 0.0;
            for (int k = k0; k < k1; ++k) {
                P *= u[2 * k];
                u[2 * N + 2 * k] = P;   // P'_k
                u[4 * N + k]     = R;   // R_k = prod_{m = k0}^{k-1} c_m
                T += R * P;
                R *= u[2 * N + 2 * k + 1];
            }
            if (k0 < k1) { u[5 * N + v] = P; u[5 * N + NW + v] = T; u[5 * N + 2 * NW + v] = R; }
        }
    }
    if (lineTabS) SB_CUDA(cudaFree(lineTabS));
    SB_CUDA(cudaMalloc((void**)&lineTabS, u.size() * sizeof(double)));
    SB_CUDA(cudaMemcpy(lineTabS, u.data(), u.size() * sizeof(double), cudaMemcpyHostToDevice));
    slay      = makeSLay(lay);
    lineSplit = true;
    // The TMA-staged persistent kernel takes over where there is enough work to keep every SM's ring busy; its
    // arithmetic is vertline_fused_k's, operation for operation.
    static int nsm = 0;
    if (!nsm) { int dv = 0; cudaGetDevice(&dv); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dv); }
    const long long ntiles = (long long)(((lay.nx + 1) / 2 + 31) / 32) * lay.ny;
    lineTma = lineTmaAllowed && k::vertline_split_fused() && k::vertline_split_nw() == k::vertline_tma_nw() && k::vertline_tma_fits(N, false) &&
              (ntiles >= 2LL * nsm || lineTmaForce);
}

// Line relaxation for a horizontally varying metric (any map with x / y stretching): the columns no longer share one
// matrix, so the Thomas factorisation is recomputed per column inside vertline_tma_k<GENERAL>; what is prepared here is
// g = 1 / d' below every chunk (line_gstart_k), the vertical tables and the dgtsv-would-pivot check.
void Op::buildGeneralLine(double sLo, double sHi)
{
    const int N = lay.nz;
    if (!lineTmaAllowed || !k::vertline_tma_fits(N, true)) return;  // vertline_k stays (small or odd depths)
    const int NWc = k::vertline_tma_nw(), CL = N / NWc;
    if (!gstart) SB_CUDA(cudaMalloc((void**)&gstart, (size_t)(NWc - 1) * lay.nx * lay.ny * sizeof(double)));
    SB_CUDA(cudaMemsetAsync(pivotFlag, 0, sizeof(int), ctx->st));
    const Coef c = coef();
    k::line_gstart(st(), lay, J, c.mxl, c.myl, c.mzl, alpha / beta, sLo, sHi, CL, gstart, pivotFlag);
    double bad = (double)readPivotFlag(true);
    ctx->allreduceMax(&bad, 1);
    if (bad != 0.0) return;  // dgtsv would interchange rows somewhere: the general kernel reports that at run time
    if (!lineTabG) SB_CUDA(cudaMalloc((void**)&lineTabG, 2 * (size_t)N * sizeof(double)));
    SB_CUDA(cudaMemcpyAsync(lineTabG, c.mzl, 2 * (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, ctx->st));  // MzL | MzR are contiguous in mtab
    slay        = makeSLay(lay);
    lineSplit   = true;
    lineGeneral = true;
    lineTma     = true;
}

// PoissonOp::checkForNullSpace (PoissonOp.cpp:670-696): L[1] == 0 to smallReal?
bool Op::checkForNullSpace()
{
    double* ones = alloc();
    double* L1   = alloc();
    k::fill(st(), ones, lay.n, 1.0);
    applyOp(L1, ones, true);
    const double smallReal = 1.0e4 * std::numeric_limits<double>::epsilon();  // SOMAR_Constants.H:63
    const double mx        = norm(L1, 0);
    ctx->sync();
    cudaFree(ones); cudaFree(L1);
    return !(mx > smallReal);
}

void Op::finalize()
{
    cacheMatrixElements();
    const double smallReal = 1.0e4 * std::numeric_limits<double>::epsilon();
    // RealCmp::neq(alpha, 0) (SOMAR_Constants.H:94-105)
    const bool alphaIsZero = std::abs(alpha - 0.0) <= smallReal * std::max(std::abs(alpha), 0.0);
    hasNullSpace = alphaIsZero ? checkForNullSpace() : false;
    defineCF();
    finalized    = true;
}

// ---------------------------------------------------------------------------------------------
// LevelData::exchange(m_exCopier): faces only (Copier::trimEdges, PoissonOp.cpp:122-123);
// periodic images are part of the copier.
void Op::exchange(double* phi)
{
    SideBC only[3][2];
    bool   any = false;
    for (int d = 0; d < 3; ++d)
        for (int s = 0; s < 2; ++s) {
            only[d][s] = side[d][s];
            if (sideIsBC(only[d][s].kind)) only[d][s].kind = -1;
            if (only[d][s].kind == SIDE_PERIODIC_SELF) any = true;
        }
    if (any) k::fill_ghosts(st(), lay, phi, only, dim);
    if (ctx->nranks > 1) ctx->comm->exchangeFaces(*this, phi);
}

// PoissonOp::applyBCs (PoissonOp.H:139-163, PoissonOp.cpp:726-763): exchange, homogeneous coarse-fine
// ghosts on the sides of a refined patch (CFInterp::homogInterpAtCFI), physical BCs.  The boundary
// conditions of this ABI are Robin pairs (alpha, beta) with no boundary data (the projector's
// HomogNeumBC, AMRNSLevelBC.cpp:48-52), so a_homogPhysBCs = false gives the same ghosts: both values
// are accepted (AMRNSLevel::projectPredict passes false, AMRNSLevelProject.cpp:123,147).
void Op::applyBCs(double* phi, bool homog)
{
    (void)homog;
    k::fill_ghosts(st(), lay, phi, side, dim);
    if (ctx->nranks > 1) ctx->comm->exchangeFaces(*this, phi);
}
void Op::applyBCsWithEdges(double* phi)
{
    if (ctx->nranks == 1) { k::fill_ghosts_with_edges(st(), lay, phi, side, dim); return; }
    // Direction by direction, each over the ghosts of the lower directions, so that an edge ghost
    // receives what the neighbouring tile's face ghost holds (the effect of the reference's
    // CornerCopier exchange, PoissonOp.cpp:1115-1125); then the physical-physical domain edges.
    for (int d = 0; d < 3; ++d) {
        if (dim == 2 && d == 1) continue;
        const int e0 = d == 0 ? 0 : 1, e1 = (d == 2 && dim == 3) ? 1 : 0;
        k::fill_ghosts_dir(st(), lay, phi, d, side[d][0], side[d][1], e0, e1);
        ctx->comm->exchangeDir(*this, phi, d, e0, e1);
    }
    k::extrap_domain_edges(st(), lay, phi, side, dim);
}

void Op::applyOp(double* lhs, double* phi, bool homog)
{
    applyBCs(phi, homog);
    k::apply_op(st(), lay, coef(), lhs, phi);
}
void Op::residual(double* res, double* phi, const double* rhs, bool homog)
{
    applyBCs(phi, homog);
    cudaEvent_t e0;
    ctx->profBegin("residual", depth, &e0);
    k::residual(st(), lay, coef(), res, phi, rhs);
    ctx->profEnd("residual", depth, e0);
}

int Op::readPivotFlag(bool reset)
{
    int flag = 0;
    SB_CUDA(cudaMemcpyAsync(&flag, pivotFlag, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    ctx->sync();
    if (flag && reset) SB_CUDA(cudaMemset(pivotFlag, 0, sizeof(int)));
    return flag;
}
void Op::checkPivot()
{
    if (relaxMethod != SB_RELAX_VERTLINE || lineFast) return;  // only vertline_k raises the flag
    double f = (double)readPivotFlag(true);
    ctx->allreduceMax(&f, 1);
    const int flag = (int)f;
    if (flag) {
        SB_FAIL("vertical line relaxation met a column where LAPACK dgtsv pivots or is singular (flag " + std::to_string(flag) +
                "); the B200 path does not reproduce that branch");
    }
}

// PoissonOp::relax (PoissonOp.cpp:917-955) and the relaxers behind it.
// Line relaxation on colour-split storage: convert once, iterate, convert back.  The call order
// of the reference is kept: physical + exchange ghosts before the first colour, exchange ghosts
// only before the second (PoissonOp.cpp:1957-1965).
void Op::relaxLineSplit(double* cor, const double* res, int iters, bool resUnchanged, int pre)
{
    if (!sp[0]) {
        // with neighbouring tiles the two colour arrays of the correction live in a block the neighbours can write
        if (ctx->nranks > 1 && ctx->peerHalo) {
            haloLine.reset(new PeerHalo);
            haloLine->setup(*this, slay, &sp[0], &sp[1]);
        }
        for (int q = haloLine ? 2 : 0; q < 4; ++q) {
            SB_CUDA(cudaMalloc((void**)&sp[q], slay.n * sizeof(double)));
            SB_CUDA(cudaMemsetAsync(sp[q], 0, slay.n * sizeof(double), ctx->st));
        }
        splitResSrc = nullptr;
        tmaMapsReady = false;
    }
    if (lineTma && !tmaMapsReady) {
        for (int c2 = 0; c2 < 2; ++c2) k::vertline_tma_make_maps(slay, sp[c2], sp[2 + c2], &tmaOth[c2], &tmaRhs[c2]);
        tmaMapsReady = true;
    }
    cudaEvent_t e0;
    ctx->profBegin("linesplit_convert", depth, &e0);
    // the right-hand side is divided by beta J on the way: level by level (shared matrix: J = J(z)) or cell by cell
    const double* scaleK = lineGeneral ? nullptr : lineTab;   // s_k = 1 / (beta J_k) = lineTab[0..nz)
    const double* scaleJ = lineGeneral ? J : nullptr;
    if (pre == RELAX_PRE_PRECOND) {
        k::split_precond(st(), lay, slay, res, Dinv, scaleK, sp[0], sp[1], sp[2], sp[3], scaleJ, beta,
                         coefUniform ? colTab + lay.nz : nullptr);
        splitResSrc = res;
    } else {
        if (!(resUnchanged && splitResSrc == res)) {
            k::split_field(st(), lay, slay, res, sp[2], sp[3], scaleK, nullptr, scaleJ, beta);
            splitResSrc = res;
        }
        k::split_field(st(), lay, slay, cor, sp[0], sp[1], nullptr, pre == RELAX_PRE_SHIFT ? shiftBuf : nullptr);
    }
    ctx->profEnd("linesplit_convert", depth, e0);
    if (ctx->nranks > 1) {
        // Neighbouring tiles: the CTAs that own cells of an exchanged face layer run first; their
        // layers travel on a second stream while the interior of the pass is computed.  Ghost
        // values and their consumers are the same as in the serial order below.
        if (!ctx->commSt) {
            SB_CUDA(cudaStreamCreateWithFlags(&ctx->commSt, cudaStreamNonBlocking));
            SB_CUDA(cudaEventCreateWithFlags(&ctx->evEdge, cudaEventDisableTiming));
            SB_CUDA(cudaEventCreateWithFlags(&ctx->evHalo, cudaEventDisableTiming));
            SB_CUDA(cudaEventCreateWithFlags(&ctx->evPost, cudaEventDisableTiming));
        }
        int nbMask = 0;
        for (int d = 0; d < 2; ++d)
            for (int s = 0; s < 2; ++s)
                if (side[d][s].kind == SIDE_NEIGHBOR) nbMask |= 1 << (2 * d + s);
        const int  last = 2 * iters - 1;
        const bool peer = haloLine && haloLine->ready;
        // peer posts run on the compute stream between the edge and the interior part of a pass (SB_HALO_FORK=1: on
        // the second stream, which relies on the two streams making progress concurrently)
        static const bool haloFork = [] { const char* e = getenv("SB_HALO_FORK"); return e && std::string(e) == "1"; }();
        static const bool haloFuse = [] { const char* e = getenv("SB_HALO_FUSED"); return !(e && std::string(e) == "0"); }();
        const bool fused = peer && haloFuse && lineTma && nbMask && slay.nx >= 2 && slay.ny >= 2;
        auto passes = [&](bool prof) {
            bool forked = false;
            // both sides are done with the ghosts of the previous relaxation on this depth
            if (peer) { haloLine->post(st(), slay, sp[0], sp[1], 0); haloLine->wait(st()); }
            for (int n = 0; n <= last; ++n) {
                const int pass = n & 1;
                k::fill_ghosts_split(st(), slay, sp[0], sp[1], side, dim, pass == 0);
                if (peer) {
                    if (n == 0) haloLine->post(st(), slay, sp[0], sp[1], 3);
                    haloLine->wait(st());
                } else if (n == 0) ctx->comm->exchangeFacesSplit(*this, sp[0], sp[1]);
                else SB_CUDA(cudaStreamWaitEvent(st(), ctx->evHalo, 0));
                if (prof) ctx->profBegin("vertline", depth, &e0);
                if (fused) {
                    // one kernel: the tiles along the exchanged sides first, their results stored into the neighbours' ghost
                    // cells by the sweep that computes them, the arrival published by the last such tile (sb_line_tma.cu)
                    linePass(pass, 0, nbMask, /*fusedHalo*/ n < last);
                } else if (nbMask && n < last) {
                    linePass(pass, 1, nbMask);
                    if (peer && !haloFork) {
                        haloLine->post(st(), slay, sp[0], sp[1], 1 << pass);
                    } else {
                        SB_CUDA(cudaEventRecord(ctx->evEdge, st()));
                        SB_CUDA(cudaStreamWaitEvent(ctx->commSt, ctx->evEdge, 0));
                        if (peer) { haloLine->post(ctx->commSt, slay, sp[0], sp[1], 1 << pass); forked = true; }
                        else {
                            ctx->comm->exchangeFacesSplit(*this, sp[0], sp[1], ctx->commSt);
                            SB_CUDA(cudaEventRecord(ctx->evHalo, ctx->commSt));
                        }
                    }
                    linePass(pass, 2, nbMask);
                } else {
                    linePass(pass);
                    if (!peer && n < last) SB_CUDA(cudaEventRecord(ctx->evHalo, st()));
                }
                if (prof) ctx->profEnd("vertline", depth, e0);
            }
            if (forked) {  // join the second stream
                SB_CUDA(cudaEventRecord(ctx->evPost, ctx->commSt));
                SB_CUDA(cudaStreamWaitEvent(st(), ctx->evPost, 0));
            }
        };
        // The whole pass loop -- kernels on two streams, the NCCL face exchanges between them -- is captured once per
        // iteration count and replayed: at N > 1 every depth is launch-bound otherwise (five launches, a grouped
        // send/recv and two events per colour pass).  SB_MR_GRAPH=0 keeps the plain launches.
        static const bool mrGraph = [] { const char* e = getenv("SB_MR_GRAPH"); return !(e && std::string(e) == "0"); }();
        if (mrGraph && !ctx->isProfiling()) {
            RelaxGraph& g = relaxGraphs[iters];
            if (!g.exec && ++relaxCallsSeen >= 2) {
                cudaGraph_t     graph = nullptr;
                const long long l0    = k::launch_count();
                SB_CUDA(cudaStreamBeginCapture(st(), cudaStreamCaptureModeThreadLocal));
                passes(false);
                SB_CUDA(cudaStreamEndCapture(st(), &graph));
                g.kernels = k::launch_count() - l0;
                k::note_launches(-g.kernels);  // captured, not launched
                SB_CUDA(cudaGraphInstantiate(&g.exec, graph, 0));
                SB_CUDA(cudaGraphDestroy(graph));
            }
            if (g.exec) {
                SB_CUDA(cudaGraphLaunch(g.exec, st()));
                k::note_launches(g.kernels);
            } else passes(false);
        } else passes(ctx->isProfiling());
    } else {
        // Small depths: replay the pass loop as a CUDA graph (captured on the second call with this
        // iteration count, once every kernel has been configured by a plain run).
        static const long long graphCells = [] { const char* e = getenv("SB_GRAPH_CELLS"); return e ? atoll(e) : (1LL << 23); }();
        const bool small = (long long)lay.nx * lay.ny * lay.nz <= graphCells;
        if (small && !ctx->isProfiling()) {
            RelaxGraph& g = relaxGraphs[iters];
            if (!g.exec && ++relaxCallsSeen >= 2) {
                cudaGraph_t     graph = nullptr;
                const long long l0    = k::launch_count();
                SB_CUDA(cudaStreamBeginCapture(st(), cudaStreamCaptureModeThreadLocal));
                linePasses(iters);
                SB_CUDA(cudaStreamEndCapture(st(), &graph));
                g.kernels = k::launch_count() - l0;
                k::note_launches(-g.kernels);  // captured, not launched
                SB_CUDA(cudaGraphInstantiate(&g.exec, graph, 0));
                SB_CUDA(cudaGraphDestroy(graph));
            }
            if (g.exec) {
                SB_CUDA(cudaGraphLaunch(g.exec, st()));
                k::note_launches(g.kernels);
            } else linePasses(iters);
        } else {
            for (int it = 0; it < iters; ++it)
                for (int pass = 0; pass < 2; ++pass) {
                    k::fill_ghosts_split(st(), slay, sp[0], sp[1], side, dim, pass == 0);
                    ctx->profBegin("vertline", depth, &e0);
                    linePass(pass);
                    ctx->profEnd("vertline", depth, e0);
                }
        }
    }
    ctx->profBegin("linesplit_convert", depth, &e0);
    k::unsplit_field(st(), lay, slay, cor, sp[0], sp[1]);
    ctx->profEnd("linesplit_convert", depth, e0);
}

// Point GSRB on colour-split storage (gsrb_split_k): convert once, iterate, convert back.  Call
// order of PoissonOp::gsrb_relax (PoissonOp.cpp:1833-1870): applyBCs (physical + exchange ghosts)
// before the first colour, exchange only before the second.
void Op::relaxGsrbSplit(double* cor, const double* res, int iters, bool resUnchanged, bool shift)
{
    if (!sg[0]) {
        slayG = makeSLay(lay, 1);
        if (ctx->nranks > 1 && ctx->peerHalo) {
            haloGsrb.reset(new PeerHalo);
            haloGsrb->setup(*this, slayG, &sg[0], &sg[1]);
        }
        for (int q = haloGsrb ? 2 : 0; q < 8; ++q) {
            SB_CUDA(cudaMalloc((void**)&sg[q], slayG.n * sizeof(double)));
            SB_CUDA(cudaMemsetAsync(sg[q], 0, slayG.n * sizeof(double), ctx->st));
        }
        splitResSrcG  = nullptr;
        gsrbCoefSplit = false;
    }
    if (!gsrbCoefSplit) {
        k::split_field(st(), lay, slayG, J, sg[4], sg[5], nullptr);
        k::split_field(st(), lay, slayG, Dinv, sg[6], sg[7], nullptr);
        gsrbCoefSplit = true;
    }
    if (!(resUnchanged && splitResSrcG == res)) {
        k::split_field(st(), lay, slayG, res, sg[2], sg[3], nullptr);
        splitResSrcG = res;
    }
    k::split_field(st(), lay, slayG, cor, sg[0], sg[1], nullptr, shift ? shiftBuf : nullptr);
    double* const       p[2] = {sg[0], sg[1]};
    const double* const r[2] = {sg[2], sg[3]};
    const double* const Jc[2] = {sg[4], sg[5]};
    const double* const Dc[2] = {sg[6], sg[7]};
    cudaEvent_t e0;
    // neighbouring tiles: the cells a pass has updated go straight into the neighbour's ghost cells (sb_halo.cu); the
    // bare post / wait pair first says that both sides are done with the ghosts of the previous call
    const bool peer = haloGsrb && haloGsrb->ready;
    if (peer) { haloGsrb->post(st(), slayG, sg[0], sg[1], 0); haloGsrb->wait(st()); }
    for (int it = 0; it < iters; ++it)
        for (int pass = 0; pass < 2; ++pass) {
            const bool physToo = pass == 0;
            k::fill_ghosts_split(st(), slayG, sg[0], sg[1], side, dim, physToo);
            k::fill_ghosts_split_z(st(), slayG, sg[0], sg[1], side[2][0], side[2][1], physToo);
            if (peer) {
                if (it == 0 && pass == 0) haloGsrb->post(st(), slayG, sg[0], sg[1], 3);
                haloGsrb->wait(st());
            } else if (ctx->nranks > 1) ctx->comm->exchangeFacesSplit(*this, sg[0], sg[1], nullptr, &slayG);
            ctx->profBegin("gsrb", depth, &e0);
            k::gsrb_split_pass(st(), slayG, coef(), p, r, Jc, Dc, pass);
            ctx->profEnd("gsrb", depth, e0);
            if (peer && !(it == iters - 1 && pass == 1)) haloGsrb->post(st(), slayG, sg[0], sg[1], 0, pass);
        }
    k::unsplit_field(st(), lay, slayG, cor, sg[0], sg[1]);
}

void Op::linePass(int pass, int region, int nbMask, bool fusedHalo)
{
    if (lineTma) {
        LineTmaArgs a;
        const Coef  c = coef();
        a.mx = c.mxl; a.my = c.myl;
        a.tab = lineGeneral ? lineTabG : lineTabS;
        a.own = sp[pass];
        a.gstart = gstart;
        a.aob = alpha / beta; a.sLo = lineSLo; a.sHi = lineSHi;
        a.pass = pass; a.region = region; a.nbMask = nbMask; a.nbx = a.ntiles = 0;
        a.fault = ctx->parent ? ctx->parent->fault : ctx->fault;
        a.halo  = fusedHalo ? haloLine->devCopy : nullptr;
        k::vertline_tma_pass(st(), slay, tmaOth[1 - pass], tmaRhs[pass], a, lineGeneral);
        return;
    }
    k::vertline_split_pass(st(), slay, coef(), lineTabS, sp[pass], sp[1 - pass], sp[2 + pass], pass, region, nbMask);
}
void Op::linePasses(int iters)
{
    for (int it = 0; it < iters; ++it)
        for (int pass = 0; pass < 2; ++pass) {
            k::fill_ghosts_split(st(), slay, sp[0], sp[1], side, dim, pass == 0);
            linePass(pass);
        }
}

void Op::relax(double* cor, const double* res, int iters, bool resUnchanged, int pre)
{

    const bool gsrbSplit = relaxMethod == SB_RELAX_GSRB && iters >= 2 && !gsrbNatural;
    if (pre != RELAX_PRE_NONE) {
        if (relaxMethod == SB_RELAX_VERTLINE && lineSplit && iters >= 2) { relaxLineSplit(cor, res, iters, resUnchanged, pre); return; }
        if (pre == RELAX_PRE_PRECOND) k::mult_valid(st(), lay, cor, res, Dinv);
        else if (!gsrbSplit) k::add_scalar_valid(st(), lay, cor, shiftBuf);
    }
    if (gsrbSplit) { relaxGsrbSplit(cor, res, iters, resUnchanged, pre == RELAX_PRE_SHIFT); return; }
    switch (relaxMethod) {
        case SB_RELAX_NONE: break;
        case SB_RELAX_GSRB:  // PoissonOp.cpp:1833-1870
            for (int it = 0; it < iters; ++it) {
                cudaEvent_t e0;
                applyBCs(cor, true);
                ctx->profBegin("gsrb", depth, &e0);
                k::gsrb_pass(st(), lay, coef(), cor, res, 0);
                ctx->profEnd("gsrb", depth, e0);
                exchange(cor);
                ctx->profBegin("gsrb", depth, &e0);
                k::gsrb_pass(st(), lay, coef(), cor, res, 1);
                ctx->profEnd("gsrb", depth, e0);
            }
            break;
        case SB_RELAX_VERTLINE: {  // PoissonOp.cpp:1927-2010
            if (iters == 0) return;
            if (lineSplit && iters >= 2) { relaxLineSplit(cor, res, iters, resUnchanged, RELAX_PRE_NONE); return; }
            const size_t n  = (size_t)((lay.nx + 1) / 2) * lay.ny * lay.nz;
            double*      wd = lineFast ? nullptr : (double*)ctx->getScratch(2 * n * sizeof(double));
            double*      wb = wd + (lineFast ? 0 : n);
            for (int it = 0; it < iters; ++it)
                for (int pass = 0; pass < 2; ++pass) {
                    if (pass == 0) applyBCs(cor, true);
                    else exchange(cor);
                    cudaEvent_t e0;
                    ctx->profBegin("vertline", depth, &e0);
                    if (lineFast) k::vertline_smem_pass(st(), lay, coef(), lineTab, cor, res, pass);
                    else k::vertline_pass(st(), lay, coef(), cor, res, pass, wd, wb, pivotFlag);
                    ctx->profEnd("vertline", depth, e0);
                }
            break;
        }
        case SB_RELAX_JACOBI: {  // PoissonOp.cpp:1709-1735
            double* r = (double*)ctx->getScratch(lay.n * sizeof(double));
            for (int it = 0; it < iters; ++it) {
                residual(r, cor, res, true);
                k::jacobi(st(), lay, coef(), cor, r, -1);
            }
            break;
        }
        case SB_RELAX_JACOBIRB: {  // PoissonOp.cpp:1741-1775: ONE residual per iteration, then both colours from it
            double* r = (double*)ctx->getScratch(lay.n * sizeof(double));
            for (int it = 0; it < iters; ++it) {
                residual(r, cor, res, true);
                k::jacobi(st(), lay, coef(), cor, r, 0);
                exchange(cor);
                k::jacobi(st(), lay, coef(), cor, r, 1);
            }
            break;
        }
        default: SB_FAIL("relaxation method " + std::to_string(relaxMethod) + " is not available on the B200 path (sequential GS cannot be parallelised bit-faithfully)");
    }
}

// PoissonOp::preCond (PoissonOp.cpp:893-911)
void Op::preCond(double* phi, const double* rhs, int relaxIters)
{
    k::mult_valid(st(), lay, phi, rhs, Dinv);
    relax(phi, rhs, relaxIters);
}

// PoissonOp::removeKernel (PoissonOp.cpp:821-846) -> Integral::sum (Integral.cpp:249-267)
bool Op::removeKernel(double* phi, bool defer)
{
    if (!hasNullSpace) return false;
    const double dv = dXi[0] * dXi[1] * dXi[2];
    k::reduce_boxes(st(), lay, boxlist(), 4, phi, J, dim == 2 ? dXi[0] * dXi[2] : dv, redPartial, redOut, nullptr,
                    coefUniform ? colTab : nullptr);
    // (sum, vol) over the boxes of this rank, then over the ranks: all on the stream, the host is not involved
    k::sum_boxes(st(), redOut, nlocal(), 2, shiftBuf);
    if (ctx->nranks > 1) {
        if (!ctx->comm) SB_FAIL("nranks > 1 but sb_comm_init was not called");
        ctx->comm->allreduceDevice(shiftBuf, 2, false, st());
    }
    if (defer) return true;
    k::add_scalar_valid(st(), lay, phi, shiftBuf);
    return false;
}

// StateOps::norm (LDFABOps.cpp:134-164) with FArrayBox::norm (FArrayBox.cpp:56-150)
double Op::norm(const double* x, int p, double powScale)
{
    if (p < 0 || p > 2) SB_FAIL("norm type must be 0, 1 or 2");
    const int nl = nlocal();
    k::reduce_boxes(st(), lay, boxlist(), p, x, nullptr, 0.0, redPartial, redOut);
    SB_CUDA(cudaMemcpyAsync(ctx->hpin, redOut, nl * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    ctx->sync();
    double ret = 0.0;
    for (int b = 0; b < nl; ++b) {
        const double numPts = (double)boxes[local[b]].numPts();
        double       boxVal;
        if (p == 0) boxVal = ctx->hpin[b];
        else if (p == 1) boxVal = ctx->hpin[b] / numPts;
        else boxVal = std::sqrt(ctx->hpin[b] / numPts);
        if (p == 0) ret = std::max(ret, boxVal);
        else ret += std::pow(boxVal, p);
    }
    if (p == 0) { ctx->allreduceMax(&ret, 1); }
    else {
        ctx->allreduceSum(&ret, 1);
        ret = std::pow(ret * powScale, 1.0 / (double)p);
    }
    return ret;
}

// StateOps::dotProduct (LDFABOps.cpp:98-121)
double Op::dotProduct(const double* a, const double* b)
{
    const int nl = nlocal();
    k::reduce_boxes(st(), lay, boxlist(), 3, a, b, 0.0, redPartial, redOut);
    SB_CUDA(cudaMemcpyAsync(ctx->hpin, redOut, nl * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    ctx->sync();
    double val = 0.0;
    for (int i = 0; i < nl; ++i) val += ctx->hpin[i];
    ctx->allreduceSum(&val, 1);
    return val;
}

// PoissonOp::MGRestrict (PoissonOp.cpp:993-1018)
void Op::MGRestrict(Op& crse, double* crseRes, const double* fineRes)
{
    int ref[3];
    for (int d = 0; d < 3; ++d) ref[d] = domain.size(d) / crse.domain.size(d);
    k::restrict_avg(st(), lay, crse.lay, ref, crseRes, fineRes);
}

// PoissonOp::MGProlong (PoissonOp.cpp:1032-1152)
bool Op::MGProlong(Op& crse, double* finePhi, double* crseCor, int order, bool deferKernel)
{
    int ref[3];
    for (int d = 0; d < 3; ++d) ref[d] = domain.size(d) / crse.domain.size(d);
    if (order >= 1) {
        crse.applyBCs(crseCor, true);
        k::prolong_linear(st(), lay, crse.lay, ref, finePhi, crseCor);  // constant + slopes
    } else {
        k::prolong_const(st(), lay, crse.lay, ref, finePhi, crseCor);
    }
    if (order >= 2) {
        bool noRoom = false;  // "Is there room for the stencil?" :1095-1104
        for (const Box3& b : crse.boxes)
            for (int d = 0; d < 3; ++d)
                if (!(dim == 2 && d == 1) && b.size(d) < 4) noRoom = true;
        if (!noRoom) {
            k::prolong_quad1(st(), lay, crse.lay, ref, finePhi, crseCor);
            crse.applyBCsWithEdges(crseCor);
            k::prolong_quad2(st(), lay, crse.lay, ref, finePhi, crseCor, dim);
        }
    }
    return removeKernel(finePhi, deferKernel);
}

// PoissonOp::levelDivergence (PoissonOp.cpp:1568-1610)
void Op::levelDivergence(double* div, double* const vel[3])
{
    const double m[3] = {1.0 / dXi[0], 1.0 / dXi[1], 1.0 / dXi[2]};  // mult/dXi with mult = 1
    k::divergence(st(), lay, div, vel[0], vel[1], vel[2], m[0], m[1], m[2], dim);
}

// AMRNSLevel::sendToAdvectingVelocity / sendToCartesianVelocity (AMRNSLevelFill.cpp:194-280).  The
// reference fills dx/dXi per FAB over velFAB.box() -- the box's faces grown by the FluxBox's ghost
// width -- with xi accumulated from that small end (GeoSourceInterface.cpp:166-199), hence `ghost`.
// A face shared by two boxes is scaled once, with the tables of the box it is the low face of.
void Op::scaleVelocity(double* const vel[3], int ghost, bool toAdvecting)
{
    if (ghost < 0) SB_FAIL("ghost width must be >= 0");
    for (int lb = 0; lb < nlocal(); ++lb) {
        const Box3& b = boxes[local[lb]];
        for (int d = 0; d < 3; ++d) {
            if (dim == 2 && d == 1) continue;
            // the other directions in the reference's order (fcDir + offset) % SpaceDim, mapped to slots
            int mus[2], nmu = 0;
            if (dim == 3) { mus[0] = (d + 1) % 3; mus[1] = (d + 2) % 3; nmu = 2; }
            else { mus[0] = d == 0 ? 2 : 0; nmu = 1; }
            std::vector<double> tab;
            size_t              off[2] = {0, 0};
            for (int m = 0; m < nmu; ++m) {
                const int           mu = mus[m], n = b.size(mu);
                std::vector<double> t  = map.dxdXi(mu, dXi[mu], b.lo[mu] - ghost, n + 2 * ghost, 0);
                off[m] = tab.size();
                tab.insert(tab.end(), t.begin() + ghost, t.begin() + ghost + n);
            }
            double* dt = (double*)ctx->getScratch(tab.size() * sizeof(double));
            SB_CUDA(cudaMemcpyAsync(dt, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
            int blo[3], n[3];
            for (int i = 0; i < 3; ++i) { blo[i] = b.lo[i] - tile.lo[i]; n[i] = b.size(i); }
            if (b.hi[d] == tile.hi[d]) n[d] += 1;  // the tile's last face belongs to this box
            k::scale_faces_box(st(), lay, blo, n, mus[0], nmu == 2 ? mus[1] : 0, dt + off[0], nmu == 2 ? dt + off[1] : nullptr, vel[d],
                               !toAdvecting);
            ctx->sync();  // scratch and tab are reused
        }
    }
}

// PoissonOp::levelGradient (PoissonOp.cpp:1486-1545)
void Op::levelGradient(double* const grad[3], double* phi, bool homog)
{
    applyBCs(phi, homog);
    const double smallReal = 1.0e4 * std::numeric_limits<double>::epsilon();
    const bool   scaleBeta = !(std::abs(beta - 1.0) <= smallReal * std::max(std::abs(beta), 1.0));
    for (int d = 0; d < 3; ++d) {
        if (dim == 2 && d == 1) continue;
        k::gradient(st(), lay, grad[d], phi, Jgup[d], d, 1.0 / dXi[d], beta, scaleBeta);
    }
}

}  // namespace sb
