// sb_capi.cpp -- the extern "C" boundary declared in include/somar_b200.h.  No exception and no
// C++ type crosses it; failures set the thread-local message read by sb_last_error().
#include <algorithm>
#include <cmath>
#include <cstring>

#include "sb_comm.h"
#include "sb_host.h"

using namespace sb;

static thread_local std::string g_err;

#define SB_TRY try {
#define SB_END                                            \
    return 0;                                             \
    }                                                     \
    catch (const std::exception& e) { g_err = e.what(); return 1; } \
    catch (...) { g_err = "unknown error"; return 1; }
#define REQ(p) if (!(p)) SB_FAIL("null argument: " #p)

// Every entry point runs on its context's device whatever the caller's current device is (a host that
// drives several GPUs from one thread), and restores the caller's device on the way out.
namespace {
struct DeviceGuard {
    int prev = -1, dev;
    explicit DeviceGuard(int d) : dev(d)
    {
        if (dev < 0) return;
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; return; }
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
int devOf(sb_context* c) { return c ? c->c.device : -1; }
int devOf(sb_op* o) { return o && o->op ? o->op->ctx->device : -1; }
int devOf(sb_field* f) { return f && f->f.op ? f->f.op->ctx->device : -1; }
int devOf(sb_solver* s) { return s && s->s.op ? s->s.op->ctx->device : -1; }
int devOf(sb_amr_solver* s) { return s && !s->s.ops.empty() && s->s.ops[s->s.lmax] ? s->s.ops[s->s.lmax]->ctx->device : -1; }
}  // namespace
#define DEVG(h) DeviceGuard devg__(devOf(h));

extern "C" {

const char* sb_last_error(void) { return g_err.c_str(); }
int         sb_version(void) { return 100; }

int sb_context_create(sb_context** ctx, int device, int rank, int nranks)
{
    SB_TRY REQ(ctx);
    if (nranks < 1 || rank < 0 || rank >= nranks) SB_FAIL("bad rank / nranks");
    *ctx = new sb_context(device, rank, nranks);
    SB_END
}
int sb_context_destroy(sb_context* ctx) { SB_TRY DEVG(ctx) delete ctx; SB_END }
int sb_context_sync(sb_context* ctx) { SB_TRY DEVG(ctx) REQ(ctx); ctx->c.sync(); SB_END }
long long sb_context_launch_count(sb_context* ctx) { return ctx ? k::launch_count() - ctx->c.launches0 : -1; }
int sb_comm_get_unique_id(void* id128) { SB_TRY REQ(id128); Comm::getUniqueId(id128); SB_END }
int sb_comm_init(sb_context* ctx, const void* id128)
{
    SB_TRY DEVG(ctx) REQ(ctx); REQ(id128);
    if (ctx->c.comm) SB_FAIL("communicator already initialised");
    ctx->c.comm = new Comm(&ctx->c, id128);
    SB_END
}

int sb_context_timer_start(sb_context* ctx)
{
    SB_TRY DEVG(ctx) REQ(ctx);
    Context& c = ctx->c;
    if (!c.tm0) { SB_CUDA(cudaEventCreate(&c.tm0)); SB_CUDA(cudaEventCreate(&c.tm1)); }
    SB_CUDA(cudaEventRecord(c.tm0, c.st));
    SB_END
}
int sb_context_timer_stop(sb_context* ctx, double* ms)
{
    SB_TRY DEVG(ctx) REQ(ctx); REQ(ms);
    Context& c = ctx->c;
    if (!c.tm0) SB_FAIL("timer not started");
    SB_CUDA(cudaEventRecord(c.tm1, c.st));
    SB_CUDA(cudaEventSynchronize(c.tm1));
    float f = 0;
    SB_CUDA(cudaEventElapsedTime(&f, c.tm0, c.tm1));
    *ms = f;
    SB_END
}
int sb_context_profile(sb_context* ctx, int enable)
{
    SB_TRY DEVG(ctx) REQ(ctx);
    ctx->c.profResolve();
    if (enable) ctx->c.prof.clear();
    ctx->c.profiling = enable == 1;   // per launch (plain launches, no graph replay)
    ctx->c.phases    = enable == 2;   // per V-cycle phase, production launch path
    SB_END
}
int sb_context_profile_get(sb_context* ctx, const char* key, double* total_ms, long long* count)
{
    SB_TRY DEVG(ctx) REQ(ctx); REQ(key); REQ(total_ms); REQ(count);
    ctx->c.profResolve();
    auto it = ctx->c.prof.find(key);
    *total_ms = it == ctx->c.prof.end() ? 0.0 : it->second.ms;
    *count    = it == ctx->c.prof.end() ? 0 : it->second.count;
    SB_END
}

// Host-only helpers (no CUDA device needed): decomposition plan and MG schedule.
int sb_plan_line_tile_order(int nx, int ny, int nb_mask, int* order, int capacity, int* num_tiles)
{
    SB_TRY REQ(order); REQ(num_tiles);
    if (nx < 2 || ny < 2) SB_FAIL("a tile of at least 2 x 2 columns");
    *num_tiles = k::line_tma_tile_order(nx, ny, nb_mask, order, capacity);
    SB_END
}
int sb_plan_tile(const sb_level_desc* d, int rank, int nranks, int tile_lo[3], int tile_hi[3], int side_kind[6],
                 int side_neighbor[6], int* num_local_boxes)
{
    SB_TRY REQ(d); REQ(tile_lo); REQ(tile_hi); REQ(side_kind); REQ(side_neighbor);
    std::vector<Box3> boxes(d->num_boxes);
    std::vector<int>  br(d->num_boxes);
    for (int b = 0; b < d->num_boxes; ++b) {
        for (int i = 0; i < 3; ++i) { boxes[b].lo[i] = d->box_lo[3 * b + i]; boxes[b].hi[i] = d->box_hi[3 * b + i]; }
        br[b] = d->box_rank ? d->box_rank[b] : 0;
    }
    Box3 dom;
    for (int i = 0; i < 3; ++i) { dom.lo[i] = d->domain_lo[i]; dom.hi[i] = d->domain_hi[i]; }
    std::vector<Box3> tiles;
    std::vector<int>  local;
    SideBC            side[3][2];
    planDecomposition(boxes, br, dom, d->periodic, rank, nranks, tiles, local, side);
    for (int i = 0; i < 3; ++i) { tile_lo[i] = tiles[rank].lo[i]; tile_hi[i] = tiles[rank].hi[i]; }
    for (int i = 0; i < 3; ++i)
        for (int s = 0; s < 2; ++s) { side_kind[2 * i + s] = side[i][s].kind; side_neighbor[2 * i + s] = side[i][s].neighbor; }
    if (num_local_boxes) *num_local_boxes = (int)local.size();
    SB_END
}
int sb_plan_cf_stencils(const int dom_lo[3], const int dom_hi[3], const int periodic[3], const int ref[3], int num_fine_boxes,
                        const int* fine_lo, const int* fine_hi, int box, int dir, int side, int capacity, int* num_cells, int* cells,
                        double* w_first, double* w_second, double* w_mixed)
{
    SB_TRY REQ(dom_lo); REQ(dom_hi); REQ(periodic); REQ(ref); REQ(fine_lo); REQ(fine_hi); REQ(num_cells);
    Box3 dom;
    for (int i = 0; i < 3; ++i) { dom.lo[i] = dom_lo[i]; dom.hi[i] = dom_hi[i]; }
    std::vector<Box3> fb(num_fine_boxes);
    for (int b = 0; b < num_fine_boxes; ++b)
        for (int i = 0; i < 3; ++i) { fb[b].lo[i] = fine_lo[3 * b + i]; fb[b].hi[i] = fine_hi[3 * b + i]; }
    std::vector<int>    c;
    std::vector<double> w1, w2, wm;
    planCFStencils(dom, periodic, ref, fb, box, dir, side, c, w1, w2, wm);
    *num_cells = (int)(c.size() / 3);
    if (cells || w_first || w_second || w_mixed) {
        if (capacity < *num_cells) SB_FAIL("capacity too small");
        if (cells) std::copy(c.begin(), c.end(), cells);
        if (w_first) std::copy(w1.begin(), w1.end(), w_first);
        if (w_second) std::copy(w2.begin(), w2.end(), w_second);
        if (w_mixed) std::copy(wm.begin(), wm.end(), w_mixed);
    }
    SB_END
}
int sb_plan_schedule(const sb_level_desc* d, int max_depth, int* schedule, int capacity, int* num_sched)
{
    SB_TRY REQ(d); REQ(num_sched);
    std::vector<Box3> boxes(d->num_boxes);
    for (int b = 0; b < d->num_boxes; ++b)
        for (int i = 0; i < 3; ++i) { boxes[b].lo[i] = d->box_lo[3 * b + i]; boxes[b].hi[i] = d->box_hi[3 * b + i]; }
    Box3 dom;
    for (int i = 0; i < 3; ++i) { dom.lo[i] = d->domain_lo[i]; dom.hi[i] = d->domain_hi[i]; }
    const bool horiz = d->relax_method == SB_RELAX_VERTLINE;  // MGSolverI.H:154-167
    auto sc = createMGRefScheduleBoxes(d->dim, dom, d->dXi, boxes, max_depth, horiz, horiz);
    *num_sched = (int)sc.size();
    if (schedule) {
        if (capacity < (int)sc.size()) SB_FAIL("capacity too small");
        for (size_t i = 0; i < sc.size(); ++i)
            for (int k = 0; k < 3; ++k) schedule[3 * i + k] = sc[i][k];
    }
    SB_END
}

// ---- PoissonOp ------------------------------------------------------------------------------
int sb_op_create(sb_context* ctx, const sb_level_desc* desc, sb_op** op)
{
    SB_TRY DEVG(ctx) REQ(ctx); REQ(desc); REQ(op);
    *op = new sb_op{new Op(&ctx->c, *desc), true};
    SB_END
}
int sb_op_destroy(sb_op* op)
{
    SB_TRY DEVG(op) if (op) { if (op->owned) delete op->op; delete op; }
    SB_END
}

static void copy3d(Op& o, int centering, double* dev, double* host, const int lo[3], const int hi[3], bool toDevice,
                   cudaStream_t stream = nullptr)
{
    if (!stream) stream = o.ctx->st;
    // allowed region of the device array in global indices
    int alo[3], ahi[3];
    for (int d = 0; d < 3; ++d) {
        alo[d] = o.tile.lo[d] - 1;
        ahi[d] = o.tile.hi[d] + 1;
        if (d == centering) alo[d] = o.tile.lo[d];  // faces: valid faces only in their own direction
    }
    int clo[3], chi[3];
    for (int d = 0; d < 3; ++d) {
        clo[d] = std::max(lo[d], alo[d]);
        chi[d] = std::min(hi[d], ahi[d]);
        if (chi[d] < clo[d]) return;
    }
    const size_t hnx = (size_t)(hi[0] - lo[0] + 1), hny = (size_t)(hi[1] - lo[1] + 1);
    cudaMemcpy3DParms p;
    std::memset(&p, 0, sizeof(p));
    cudaPitchedPtr hp = make_cudaPitchedPtr(host, hnx * sizeof(double), hnx, hny);
    cudaPitchedPtr dp = make_cudaPitchedPtr(dev, (size_t)o.lay.px * sizeof(double), (size_t)o.lay.px, (size_t)o.lay.py);
    cudaPos hpos = make_cudaPos((size_t)(clo[0] - lo[0]) * sizeof(double), (size_t)(clo[1] - lo[1]), (size_t)(clo[2] - lo[2]));
    cudaPos dpos = make_cudaPos((size_t)(OX + clo[0] - o.tile.lo[0]) * sizeof(double), (size_t)(1 + clo[1] - o.tile.lo[1]),
                                (size_t)(1 + clo[2] - o.tile.lo[2]));
    p.extent = make_cudaExtent((size_t)(chi[0] - clo[0] + 1) * sizeof(double), (size_t)(chi[1] - clo[1] + 1),
                               (size_t)(chi[2] - clo[2] + 1));
    if (toDevice) { p.srcPtr = hp; p.srcPos = hpos; p.dstPtr = dp; p.dstPos = dpos; p.kind = cudaMemcpyHostToDevice; }
    else { p.srcPtr = dp; p.srcPos = dpos; p.dstPtr = hp; p.dstPos = hpos; p.kind = cudaMemcpyDeviceToHost; }
    SB_CUDA(cudaMemcpy3DAsync(&p, stream));
}

int sb_op_set_metric(sb_op* op, int centering, int box_id, const double* host, const int lo[3], const int hi[3])
{
    SB_TRY DEVG(op) REQ(op); REQ(host);
    Op& o = *op->op;
    if (box_id < 0 || box_id >= (int)o.boxes.size()) SB_FAIL("box_id out of range");
    if (o.boxRank[box_id] != o.ctx->rank) return 0;
    if (centering < -1 || centering > 2) SB_FAIL("bad centering");
    // copy only the part over this box (cells) / its faces
    const Box3& b = o.boxes[box_id];
    int clo[3], chi[3];
    for (int d = 0; d < 3; ++d) {
        clo[d] = std::max(lo[d], b.lo[d]);
        chi[d] = std::min(hi[d], b.hi[d] + (d == centering ? 1 : 0));
        if (chi[d] < clo[d]) return 0;  // the host array misses this box
    }
    // stage through a contiguous sub-box copy: cudaMemcpy3D handles the strides
    const size_t hnx = (size_t)(hi[0] - lo[0] + 1), hny = (size_t)(hi[1] - lo[1] + 1);
    const double* sub = host + (clo[0] - lo[0]) + hnx * ((size_t)(clo[1] - lo[1]) + hny * (size_t)(clo[2] - lo[2]));
    // describe the sub-box as its own host box with the parent's pitch: emulate by per-plane copies
    double* dev = centering < 0 ? o.J : o.Jgup[centering];
    for (int kk = clo[2]; kk <= chi[2]; ++kk) {
        const double* hp = sub + hnx * hny * (size_t)(kk - clo[2]);
        double*       dp = dev + o.lay.idx(clo[0] - o.tile.lo[0], clo[1] - o.tile.lo[1], kk - o.tile.lo[2]);
        SB_CUDA(cudaMemcpy2DAsync(dp, (size_t)o.lay.px * sizeof(double), hp, hnx * sizeof(double),
                                  (size_t)(chi[0] - clo[0] + 1) * sizeof(double), (size_t)(chi[1] - clo[1] + 1),
                                  cudaMemcpyHostToDevice, o.ctx->st));
    }
    o.ctx->sync();
    SB_END
}
int sb_op_finalize(sb_op* op) { SB_TRY DEVG(op) REQ(op); op->op->finalize(); op->op->ctx->sync(); SB_END }
int sb_op_has_null_space(sb_op* op, int* out) { SB_TRY DEVG(op) REQ(op); REQ(out); *out = op->op->hasNullSpace; SB_END }
int sb_op_halo_mode(sb_op* op, int* out)
{
    SB_TRY DEVG(op) REQ(op); REQ(out);
    const Op& o = *op->op;
    const PeerHalo* h = o.haloLine ? o.haloLine.get() : o.haloGsrb.get();
    *out = o.ctx->nranks <= 1 ? SB_HALO_NONE : (h && h->ready) ? SB_HALO_PEER : SB_HALO_NCCL;
    SB_END
}
int sb_op_new_mg_operator(sb_op* op, const int ref[3], sb_op** crse)
{
    SB_TRY DEVG(op) REQ(op); REQ(crse);
    if (!op->op->finalized) SB_FAIL("sb_op_finalize first");
    if (ref[0] == 1 && ref[1] == 1 && ref[2] == 1) SB_FAIL("newMGOperator(Unit) clones are not needed: reuse the handle");
    *crse = new sb_op{new Op(*op->op, ref), true};
    SB_END
}
int sb_op_get_info(sb_op* op, int domain_lo[3], int domain_hi[3], double dXi[3], int* num_local_boxes)
{
    SB_TRY DEVG(op) REQ(op);
    for (int d = 0; d < 3; ++d) {
        if (domain_lo) domain_lo[d] = op->op->domain.lo[d];
        if (domain_hi) domain_hi[d] = op->op->domain.hi[d];
        if (dXi) dXi[d] = op->op->dXi[d];
    }
    if (num_local_boxes) *num_local_boxes = op->op->nlocal();
    SB_END
}
int sb_op_get_coefficient(sb_op* op, int which, double* host, long long capacity)
{
    SB_TRY DEVG(op) REQ(op); REQ(host);
    Op& o = *op->op;
    if (which >= 2 && which <= 4) {
        const auto& m = o.hM[which - 2];
        if ((long long)m.size() > capacity) SB_FAIL("capacity too small");
        std::memcpy(host, m.data(), m.size() * sizeof(double));
        return 0;
    }
    int centering = -1;
    double* dev = nullptr;
    if (which == 0) dev = o.J;
    else if (which == 1) dev = o.Dinv;
    else if (which >= 5 && which <= 7) { centering = which - 5; dev = o.Jgup[centering]; }
    else SB_FAIL("bad coefficient id");
    int lo[3], hi[3];
    long long n = 1;
    for (int d = 0; d < 3; ++d) { lo[d] = o.domain.lo[d]; hi[d] = o.domain.hi[d] + (d == centering ? 1 : 0); n *= hi[d] - lo[d] + 1; }
    if (n > capacity) SB_FAIL("capacity too small");
    copy3d(o, centering, dev, host, lo, hi, false);
    o.ctx->sync();
    SB_END
}

// ---- fields ---------------------------------------------------------------------------------
int sb_field_create(sb_op* op, int centering, sb_field** f)
{
    SB_TRY DEVG(op) REQ(op); REQ(f);
    if (centering < -1 || centering > 2) SB_FAIL("bad centering");
    *f = new sb_field(op->op, centering);
    SB_END
}
int sb_field_destroy(sb_field* f) { SB_TRY DEVG(f) delete f; SB_END }
int sb_field_upload(sb_field* f, const double* host, const int lo[3], const int hi[3])
{
    SB_TRY DEVG(f) REQ(f); REQ(host);
    copy3d(*f->f.op, f->f.centering, f->f.d, const_cast<double*>(host), lo, hi, true);
    f->f.op->ctx->sync();
    SB_END
}
int sb_field_download(sb_field* f, double* host, const int lo[3], const int hi[3])
{
    SB_TRY DEVG(f) REQ(f); REQ(host);
    copy3d(*f->f.op, f->f.centering, f->f.d, host, lo, hi, false);
    f->f.op->ctx->sync();
    SB_END
}

int sb_field_upload_async(sb_field* f, const double* host, const int lo[3], const int hi[3])
{
    SB_TRY DEVG(f) REQ(f); REQ(host);
    Op& o = *f->f.op;
    // the field's zero fill ran on the compute stream: a late memset must not overwrite this copy
    SB_CUDA(cudaStreamWaitEvent(o.ctx->stream(SB_STREAM_H2D), f->f.ready, 0));
    copy3d(o, f->f.centering, f->f.d, const_cast<double*>(host), lo, hi, true, o.ctx->stream(SB_STREAM_H2D));
    SB_END
}
int sb_field_download_async(sb_field* f, double* host, const int lo[3], const int hi[3])
{
    SB_TRY DEVG(f) REQ(f); REQ(host);
    Op& o = *f->f.op;
    copy3d(o, f->f.centering, f->f.d, host, lo, hi, false, o.ctx->stream(SB_STREAM_D2H));
    SB_END
}
int sb_context_stream_wait(sb_context* ctx, int waiter, int signaller)
{
    SB_TRY DEVG(ctx) REQ(ctx);
    cudaEvent_t e;
    SB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    SB_CUDA(cudaEventRecord(e, ctx->c.stream(signaller)));
    SB_CUDA(cudaStreamWaitEvent(ctx->c.stream(waiter), e, 0));
    SB_CUDA(cudaEventDestroy(e));  // released once the recorded work completes
    SB_END
}
int sb_context_stream_sync(sb_context* ctx, int which)
{
    SB_TRY DEVG(ctx) REQ(ctx); SB_CUDA(cudaStreamSynchronize(ctx->c.stream(which))); SB_END
}

// ---- operator methods ------------------------------------------------------------------------
#define OPF(o) (*(o)->op)
#define D(fld_) ((fld_)->f.d)
static void sameOp(sb_op* op, std::initializer_list<sb_field*> fs)
{
    if (!op) SB_FAIL("null op");
    if (!op->op->finalized) SB_FAIL("sb_op_finalize has not been called");
    for (sb_field* f : fs) {
        if (!f) SB_FAIL("null field");
        if (f->f.op != op->op) SB_FAIL("field belongs to another operator / MG depth");
    }
}
int sb_op_apply_bcs(sb_op* op, sb_field* phi, int homog) { SB_TRY DEVG(op) sameOp(op, {phi}); OPF(op).applyBCs(D(phi), homog); SB_END }
int sb_op_apply_op(sb_op* op, sb_field* lhs, sb_field* phi, int homog)
{
    SB_TRY DEVG(op) sameOp(op, {lhs, phi}); OPF(op).applyOp(D(lhs), D(phi), homog); SB_END
}
int sb_op_residual(sb_op* op, sb_field* res, sb_field* phi, sb_field* rhs, int homog)
{
    SB_TRY DEVG(op) sameOp(op, {res, phi, rhs}); OPF(op).residual(D(res), D(phi), D(rhs), homog); SB_END
}
int sb_op_relax(sb_op* op, sb_field* cor, sb_field* res, int iters)
{
    SB_TRY DEVG(op) sameOp(op, {cor, res}); OPF(op).relax(D(cor), D(res), iters);
    if (OPF(op).relaxMethod == SB_RELAX_VERTLINE) OPF(op).checkPivot();
    SB_END
}
int sb_op_precond(sb_op* op, sb_field* phi, sb_field* rhs, int it) { SB_TRY DEVG(op) sameOp(op, {phi, rhs}); OPF(op).preCond(D(phi), D(rhs), it); SB_END }
int sb_op_remove_kernel(sb_op* op, sb_field* phi) { SB_TRY DEVG(op) sameOp(op, {phi}); OPF(op).removeKernel(D(phi)); SB_END }
int sb_op_norm(sb_op* op, sb_field* x, int p, double pow_scale, double* out)
{
    SB_TRY DEVG(op) sameOp(op, {x}); REQ(out); *out = OPF(op).norm(D(x), p, pow_scale); SB_END
}
int sb_op_dot(sb_op* op, sb_field* a, sb_field* b, double* out) { SB_TRY DEVG(op) sameOp(op, {a, b}); REQ(out); *out = OPF(op).dotProduct(D(a), D(b)); SB_END }
int sb_op_incr(sb_op* op, sb_field* lhs, sb_field* x, double s) { SB_TRY DEVG(op) sameOp(op, {lhs, x}); OPF(op).incr(D(lhs), D(x), s); SB_END }
int sb_op_axby(sb_op* op, sb_field* lhs, sb_field* x, sb_field* y, double a, double b)
{
    SB_TRY DEVG(op) sameOp(op, {lhs, x, y}); OPF(op).axby(D(lhs), D(x), D(y), a, b); SB_END
}
int sb_op_scale(sb_op* op, sb_field* lhs, double s) { SB_TRY DEVG(op) sameOp(op, {lhs}); OPF(op).scale(D(lhs), s); SB_END }
int sb_op_set_to_zero(sb_op* op, sb_field* lhs) { SB_TRY DEVG(op) sameOp(op, {lhs}); OPF(op).setToZero(D(lhs)); SB_END }
int sb_op_assign_local(sb_op* op, sb_field* dst, sb_field* src) { SB_TRY DEVG(op) sameOp(op, {dst, src}); OPF(op).assignLocal(D(dst), D(src)); SB_END }
int sb_op_mg_restrict(sb_op* fine, sb_op* crse, sb_field* crse_res, sb_field* fine_res)
{
    SB_TRY DEVG(fine) sameOp(fine, {fine_res}); sameOp(crse, {crse_res});
    OPF(fine).MGRestrict(OPF(crse), D(crse_res), D(fine_res));
    SB_END
}
int sb_op_mg_prolong(sb_op* fine, sb_op* crse, sb_field* fine_phi, sb_field* crse_cor, int order)
{
    SB_TRY DEVG(fine) sameOp(fine, {fine_phi}); sameOp(crse, {crse_cor});
    OPF(fine).MGProlong(OPF(crse), D(fine_phi), D(crse_cor), order);
    SB_END
}
static void fluxPtrs(sb_op* op, sb_field* const f[3], double* out[3])
{
    for (int d = 0; d < 3; ++d) {
        out[d] = nullptr;
        if (op->op->dim == 2 && d == 1) { if (f[d]) out[d] = D(f[d]); continue; }
        if (!f[d]) SB_FAIL("null flux component");
        if (f[d]->f.op != op->op) SB_FAIL("field belongs to another operator");
        if (f[d]->f.centering != d) SB_FAIL("flux component has the wrong centering");
        out[d] = D(f[d]);
    }
}
int sb_op_level_divergence(sb_op* op, sb_field* div, sb_field* const vel[3])
{
    SB_TRY DEVG(op) sameOp(op, {div});
    double* v[3]; fluxPtrs(op, vel, v);
    OPF(op).levelDivergence(D(div), v);
    SB_END
}
int sb_op_level_gradient(sb_op* op, sb_field* const grad[3], sb_field* phi, int homog)
{
    SB_TRY DEVG(op) sameOp(op, {phi});
    double* g[3]; fluxPtrs(op, grad, g);
    OPF(op).levelGradient(g, D(phi), homog);
    SB_END
}
int sb_op_send_to_advecting_velocity(sb_op* op, sb_field* const vel[3], int ghost)
{
    SB_TRY DEVG(op) sameOp(op, {});
    double* v[3]; fluxPtrs(op, vel, v);
    OPF(op).scaleVelocity(v, ghost, true);
    SB_END
}
int sb_op_send_to_cartesian_velocity(sb_op* op, sb_field* const vel[3], int ghost)
{
    SB_TRY DEVG(op) sameOp(op, {});
    double* v[3]; fluxPtrs(op, vel, v);
    OPF(op).scaleVelocity(v, ghost, false);
    SB_END
}
int sb_op_flux_incr(sb_op* op, sb_field* const vel[3], sb_field* const grad[3], double scale)
{
    SB_TRY DEVG(op) sameOp(op, {});
    double *v[3], *g[3]; fluxPtrs(op, vel, v); fluxPtrs(op, grad, g);
    for (int d = 0; d < 3; ++d) {
        if (op->op->dim == 2 && d == 1) continue;
        k::incr_valid(op->op->st(), op->op->lay, v[d], g[d], -scale, d);  // FArrayBox::plus(grad, -scale)
    }
    SB_END
}

// ---- AMRMGOperator surface ---------------------------------------------------------------------
static sb::Op* opOf(sb_field* f) { return f ? f->f.op : nullptr; }
int sb_op_apply_bcs_amr(sb_op* op, sb_field* phi, sb_field* crse_phi, int homog_phys, int homog_cfi)
{
    SB_TRY DEVG(op) sameOp(op, {phi}); (void)homog_phys;
    OPF(op).applyBCsAMR(D(phi), opOf(crse_phi), crse_phi ? D(crse_phi) : nullptr, homog_cfi != 0);
    SB_END
}
int sb_op_amr_operator(sb_op* op, sb_field* lhs, sb_field* phi_fine, sb_field* phi, sb_field* phi_crse, int homog_phys, sb_op* finer_op)
{
    SB_TRY DEVG(op) sameOp(op, {lhs, phi}); sameOp(finer_op, {phi_fine}); REQ(phi_crse); (void)homog_phys;
    OPF(op).AMROperator(D(lhs), OPF(finer_op), D(phi_fine), D(phi), *opOf(phi_crse), D(phi_crse));
    SB_END
}
int sb_op_amr_operator_nf(sb_op* op, sb_field* lhs, sb_field* phi, sb_field* phi_crse, int homog_phys)
{
    SB_TRY DEVG(op) sameOp(op, {lhs, phi}); REQ(phi_crse); (void)homog_phys;
    OPF(op).AMROperatorNF(D(lhs), D(phi), *opOf(phi_crse), D(phi_crse));
    SB_END
}
int sb_op_amr_operator_nc(sb_op* op, sb_field* lhs, sb_field* phi_fine, sb_field* phi, int homog_phys, sb_op* finer_op)
{
    SB_TRY DEVG(op) sameOp(op, {lhs, phi}); sameOp(finer_op, {phi_fine}); (void)homog_phys;
    OPF(op).AMROperatorNC(D(lhs), OPF(finer_op), D(phi_fine), D(phi));
    SB_END
}
int sb_op_amr_residual(sb_op* op, sb_field* res, sb_field* phi_fine, sb_field* phi, sb_field* phi_crse, sb_field* rhs, int homog_phys,
                       sb_op* finer_op)
{
    SB_TRY DEVG(op) sameOp(op, {res, phi, rhs}); (void)homog_phys;
    if ((finer_op != nullptr) != (phi_fine != nullptr)) SB_FAIL("finer_op and phi_fine go together");
    if (finer_op) sameOp(finer_op, {phi_fine});
    OPF(op).AMRResidual(D(res), finer_op ? finer_op->op : nullptr, phi_fine ? D(phi_fine) : nullptr, D(phi), opOf(phi_crse),
                        phi_crse ? D(phi_crse) : nullptr, D(rhs));
    SB_END
}
int sb_op_amr_norm_level(sb_op* op, sb_field* res, sb_op* finer_op, int p, double* out)
{
    SB_TRY DEVG(op) sameOp(op, {res}); REQ(out);
    *out = OPF(op).AMRNormLevel(D(res), finer_op ? finer_op->op : nullptr, p);
    SB_END
}
int sb_op_get_flux(sb_op* op, sb_field* const flux[3], sb_field* phi)
{
    SB_TRY DEVG(op) sameOp(op, {phi});
    double* g[3]; fluxPtrs(op, flux, g);
    OPF(op).getFlux(g, D(phi));
    SB_END
}
int sb_op_reflux(sb_op* op, sb_field* res, sb_field* fine_phi, sb_field* phi, sb_op* finer_op)
{
    SB_TRY DEVG(op) sameOp(op, {res, phi}); sameOp(finer_op, {fine_phi});
    OPF(op).reflux(D(res), OPF(finer_op), D(fine_phi), D(phi));
    SB_END
}
int sb_op_reflux_flux(sb_op* op, sb_field* div, sb_field* const flux[3], sb_field* const fine_flux[3], sb_op* finer_op)
{
    SB_TRY DEVG(op) sameOp(op, {div}); sameOp(finer_op, {});
    double *f[3], *ff[3]; fluxPtrs(op, flux, f); fluxPtrs(finer_op, fine_flux, ff);
    OPF(op).refluxFlux(D(div), f, OPF(finer_op), ff);
    SB_END
}
int sb_op_comp_divergence(sb_op* op, sb_field* div, sb_field* const flux[3], sb_field* const fine_flux[3], sb_op* finer_op)
{
    SB_TRY DEVG(op) sameOp(op, {div});
    double *f[3], *ff[3] = {nullptr, nullptr, nullptr};
    fluxPtrs(op, flux, f);
    if ((finer_op != nullptr) != (fine_flux != nullptr)) SB_FAIL("finer_op and fine_flux go together");
    if (finer_op) { sameOp(finer_op, {}); fluxPtrs(finer_op, fine_flux, ff); }
    OPF(op).compDivergence(D(div), f, finer_op ? finer_op->op : nullptr, ff);
    SB_END
}
int sb_op_comp_gradient(sb_op* op, sb_field* const grad[3], sb_field* phi, sb_field* crse_phi, int homog_phys, int homog_cfi)
{
    SB_TRY DEVG(op) sameOp(op, {phi}); (void)homog_phys;
    double* g[3]; fluxPtrs(op, grad, g);
    Op& o = OPF(op);
    o.applyBCsAMR(D(phi), opOf(crse_phi), crse_phi ? D(crse_phi) : nullptr, homog_cfi != 0);  // PoissonOp.cpp:1505
    const double smallReal = 1.0e4 * 2.220446049250313e-16;
    const bool   scaleBeta = !(std::abs(o.beta - 1.0) <= smallReal * std::max(std::abs(o.beta), 1.0));
    for (int d = 0; d < 3; ++d) {
        if (o.dim == 2 && d == 1) continue;
        k::gradient(o.st(), o.lay, g[d], D(phi), o.Jgup[d], d, 1.0 / o.dXi[d], o.beta, scaleBeta);
    }
    SB_END
}
int sb_op_average_down(sb_op* fine_op, sb_field* crse, sb_field* fine)
{
    SB_TRY DEVG(fine_op) sameOp(fine_op, {fine}); REQ(crse);
    OPF(fine_op).averageDownTo(*crse->f.op, D(crse), D(fine));
    SB_END
}
int sb_op_get_patch(sb_op* op, int patch_lo[3], int patch_hi[3], int tile_lo[3], int tile_hi[3])
{
    SB_TRY DEVG(op) REQ(op);
    for (int d = 0; d < 3; ++d) {
        if (patch_lo) patch_lo[d] = op->op->patch.lo[d];
        if (patch_hi) patch_hi[d] = op->op->patch.hi[d];
        if (tile_lo) tile_lo[d] = op->op->tile.lo[d];
        if (tile_hi) tile_hi[d] = op->op->tile.hi[d];
    }
    SB_END
}

// ---- AMRHybridSolver ---------------------------------------------------------------------------
int sb_amr_solver_create(sb_op* const* ops, int num_levels, int lmin, int lmax, const sb_mg_options* opt, sb_amr_solver** s)
{
    SB_TRY DEVG(ops && lmax >= 0 && lmax < num_levels ? ops[lmax] : (sb_op*)nullptr) REQ(ops); REQ(opt); REQ(s);
    std::vector<Op*> v(num_levels, nullptr);
    for (int l = 0; l < num_levels; ++l)
        if (ops[l]) {
            if (!ops[l]->op->finalized) SB_FAIL("sb_op_finalize has not been called on every level");
            v[l] = ops[l]->op;
        }
    std::unique_ptr<sb_amr_solver> p(new sb_amr_solver);
    p->s.define(v, lmin, lmax, *opt);
    v[lmax]->ctx->sync();
    *s = p.release();
    SB_END
}
int sb_amr_solver_destroy(sb_amr_solver* s) { SB_TRY DEVG(s) delete s; SB_END }
int sb_amr_solver_solve(sb_amr_solver* s, sb_field* const* phi, sb_field* const* rhs, int homog, int set_phi_to_zero, double metric,
                        sb_solver_status* status)
{
    SB_TRY DEVG(s) REQ(s); REQ(phi); REQ(rhs);
    AMRSolver& a = s->s;
    std::vector<double*>       vphi(a.lmax + 1, nullptr);
    std::vector<const double*> vrhs(a.lmax + 1, nullptr);
    for (int l = a.lbase; l <= a.lmax; ++l) {
        if (!phi[l]) SB_FAIL("phi missing on a level");
        if (phi[l]->f.op != a.ops[l]) SB_FAIL("phi does not live on this level's operator");
        vphi[l] = D(phi[l]);
        if (l >= a.lmin) {
            if (!rhs[l]) SB_FAIL("rhs missing on a level");
            if (rhs[l]->f.op != a.ops[l]) SB_FAIL("rhs does not live on this level's operator");
            vrhs[l] = D(rhs[l]);
        }
    }
    Op& top = *a.ops[a.lmax];
    cudaEvent_t e0, e1;
    SB_CUDA(cudaEventCreate(&e0)); SB_CUDA(cudaEventCreate(&e1));
    SB_CUDA(cudaEventRecord(e0, top.st()));
    SolverStatus st = a.solve(vphi, vrhs, homog != 0, set_phi_to_zero != 0, metric);
    SB_CUDA(cudaEventRecord(e1, top.st()));
    top.ctx->sync();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (status) {
        std::memset(status, 0, sizeof(*status));
        status->status         = st.status;
        status->num_iters      = a.lastIters;
        status->init_res_norm  = st.initResNorm;
        status->final_res_norm = st.finalResNorm;
        status->num_norms      = (int)std::min<size_t>(a.absResNorms.size(), SB_MAX_HISTORY);
        for (int i = 0; i < status->num_norms; ++i) status->res_norms[i] = a.absResNorms[i];
        status->solve_mode = 0;
        status->max_depth  = a.lmax - a.lmin;
        status->device_ms  = ms;
    }
    SB_END
}

// ---- solvers ---------------------------------------------------------------------------------
int sb_mgsolver_create(sb_op* top, const sb_mg_options* opt, const int* schedule, int num_sched, sb_solver** s)
{
    SB_TRY DEVG(top) sameOp(top, {}); REQ(opt); REQ(s);
    std::vector<std::array<int, 3>> sched;
    for (int i = 0; i < num_sched; ++i) sched.push_back({schedule[3 * i], schedule[3 * i + 1], schedule[3 * i + 2]});
    std::unique_ptr<sb_solver> p(new sb_solver);
    p->s.isHybrid = false;
    p->s.op       = top->op;
    p->s.opt      = *opt;
    p->s.mode     = SB_MODE_MG;
    p->s.mg.define(*top->op, *opt, sched, true);
    top->op->ctx->sync();
    *s = p.release();
    SB_END
}
int sb_hybrid_solver_create(sb_op* top, const sb_mg_options* opt, sb_solver** s)
{
    SB_TRY DEVG(top) sameOp(top, {}); REQ(opt); REQ(s);
    std::unique_ptr<sb_solver> p(new sb_solver);
    p->s.isHybrid = true;
    p->s.define(*top->op, *opt);
    top->op->ctx->sync();
    *s = p.release();
    SB_END
}
int sb_solver_destroy(sb_solver* s) { SB_TRY DEVG(s) delete s; SB_END }
int sb_solver_get_schedule(sb_solver* s, int* schedule, int capacity, int* num_sched)
{
    SB_TRY DEVG(s) REQ(s); REQ(num_sched);
    const auto& sc = s->s.mg.refSchedule;
    *num_sched     = (int)sc.size();
    if (schedule) {
        if (capacity < (int)sc.size()) SB_FAIL("capacity too small");
        for (size_t i = 0; i < sc.size(); ++i)
            for (int d = 0; d < 3; ++d) schedule[3 * i + d] = sc[i][d];
    }
    SB_END
}
int sb_solver_set_options(sb_solver* s, const sb_mg_options* opt)
{
    // LevelHybridSolver::modifyOptionsExceptMaxDepth (LevelHybridSolver.cpp:65-81): both sub-solvers follow
    SB_TRY DEVG(s) REQ(s); REQ(opt);
    if (!s->s.mg.ops.empty()) s->s.mg.modifyOptionsExceptMaxDepth(*opt);
    if (s->s.leptic) s->s.leptic->modifyOptionsExceptMaxDepth(*opt);
    const int md = s->s.opt.maxDepth;
    s->s.opt = *opt; s->s.opt.maxDepth = md;
    SB_END
}
static void fillStatus(sb_solver* s, const SolverStatus& st, sb_solver_status* out, float ms)
{
    if (!out) return;
    std::memset(out, 0, sizeof(*out));
    out->status         = st.status;
    out->num_iters      = s->s.mg.lastIters;
    out->init_res_norm  = st.initResNorm;
    out->final_res_norm = st.finalResNorm;
    // MG mode: MGSolver's own history; leptic modes: LevelHybridSolver's m_resNorms (initial norm,
    // then one entry per leptic order / V-cycle, LevelHybridSolver.cpp:312-400)
    const auto& h       = s->s.mode == SB_MODE_MG || !s->s.isHybrid ? s->s.mg.absResNorms : s->s.resNorms;
    out->num_norms      = (int)std::min<size_t>(h.size(), SB_MAX_HISTORY);
    for (int i = 0; i < out->num_norms; ++i) out->res_norms[i] = h[i];
    out->solve_mode = s->s.mode;
    out->max_depth  = s->s.mg.opt.maxDepth;
    out->device_ms  = ms;
}
int sb_solver_solve(sb_solver* s, sb_field* phi, sb_field* rhs, int homog, int set_phi_to_zero, double metric,
                    sb_solver_status* status)
{
    SB_TRY DEVG(s) REQ(s); REQ(phi); REQ(rhs);
    Op& o = *s->s.op;
    if (phi->f.op != &o || rhs->f.op != &o) SB_FAIL("fields do not live on the solver's top operator");
    cudaEvent_t e0, e1;
    SB_CUDA(cudaEventCreate(&e0)); SB_CUDA(cudaEventCreate(&e1));
    SB_CUDA(cudaEventRecord(e0, o.st()));
    SolverStatus st = s->s.solve(D(phi), D(rhs), homog, set_phi_to_zero, metric);
    SB_CUDA(cudaEventRecord(e1, o.st()));
    o.ctx->sync();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    fillStatus(s, st, status, ms);
    SB_END
}
int sb_solver_vcycle(sb_solver* s, sb_field* cor, sb_field* res)
{
    SB_TRY DEVG(s) REQ(s); REQ(cor); REQ(res);
    Op& o = *s->s.op;
    if (cor->f.op != &o || res->f.op != &o) SB_FAIL("fields do not live on the solver's top operator");
    if (s->s.mg.ops.empty()) SB_FAIL("this hybrid solver runs in pure leptic mode: it has no MGSolver to V-cycle with");
    s->s.mg.vCycle_residualEq(D(cor), D(res), 0);
    if (o.relaxMethod == SB_RELAX_VERTLINE) s->s.mg.checkPivotAll();
    SB_END
}

int sb_solver_precond_vcycle(sb_solver* s, sb_field* cor, sb_field* res)
{
    SB_TRY DEVG(s) REQ(s); REQ(cor); REQ(res);
    Op& o = *s->s.op;
    if (cor->f.op != &o || res->f.op != &o) SB_FAIL("fields do not live on the solver's top operator");
    if (s->s.mg.ops.empty()) SB_FAIL("this hybrid solver runs in pure leptic mode: it has no MGSolver to V-cycle with");
    s->s.mg.vCycle_residualEq(D(cor), D(res), 0, true);
    if (o.relaxMethod == SB_RELAX_VERTLINE) s->s.mg.checkPivotAll();
    SB_END
}

// The projection bracket on device-resident fields, any number of ranks (every call inside is collective).
int sb_project_correct(sb_solver* s, sb_field* const vel[3], sb_field* p, double proj_dt, int vel_ghost, sb_field* phi_out,
                       double* init_div_norm, double* final_div_norm, sb_solver_status* status)
{
    SB_TRY DEVG(s) REQ(s); REQ(vel);
    Op& o = *s->s.op;
    if (!s->s.isHybrid && s->s.mg.ops.empty()) SB_FAIL("solver not defined");
    sb_op tmp{&o, false};
    double* v[3]; fluxPtrs(&tmp, vel, v);
    if (p && p->f.op != &o) SB_FAIL("p does not live on the solver's operator");
    if (phi_out && phi_out->f.op != &o) SB_FAIL("phi_out does not live on the solver's operator");
    cudaEvent_t e0, e1;
    SB_CUDA(cudaEventCreate(&e0)); SB_CUDA(cudaEventCreate(&e1));
    SB_CUDA(cudaEventRecord(e0, o.st()));
    SolverStatus st = s->s.projectCorrect(v, p ? D(p) : nullptr, proj_dt, vel_ghost, phi_out ? D(phi_out) : nullptr, init_div_norm, final_div_norm);
    SB_CUDA(cudaEventRecord(e1, o.st()));
    o.ctx->sync();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    fillStatus(s, st, status, ms);
    SB_END
}
int sb_project_predict(sb_solver* s, sb_field* const vel[3], sb_field* p, double proj_dt, int vel_ghost, double norms[3], int* used_fallback,
                       sb_solver_status* status)
{
    SB_TRY DEVG(s) REQ(s); REQ(vel); REQ(p); REQ(norms);
    Op& o = *s->s.op;
    sb_op tmp{&o, false};
    double* v[3]; fluxPtrs(&tmp, vel, v);
    if (p->f.op != &o) SB_FAIL("p does not live on the solver's operator");
    bool fb = false;
    SolverStatus st = s->s.projectPredict(v, D(p), proj_dt, vel_ghost, norms, &fb);
    o.ctx->sync();
    if (used_fallback) *used_fallback = fb ? 1 : 0;
    fillStatus(s, st, status, 0.0f);
    SB_END
}

// AMRNSLevel::projectCorrect, single level (Grade5_SOMAR/AMRNSLevelProject.cpp:247-373).
int sb_project_host(sb_solver* s, double* const vel[3], double* phi, double* p, double proj_dt, double* init_div_norm,
                    double* final_div_norm, sb_solver_status* status)
{
    SB_TRY DEVG(s) REQ(s); REQ(vel);
    Op& o = *s->s.op;
    if (o.ctx->nranks != 1) SB_FAIL("sb_project_host takes whole-domain host arrays: single rank only (use the field API per rank)");
    struct Tmp {
        std::vector<double*> v;
        ~Tmp() { for (double* q : v) cudaFree(q); }
        double* get(Op& o) { v.push_back(o.alloc()); return v.back(); }
    } tmp;
    double *U[3] = {nullptr, nullptr, nullptr}, *G[3] = {nullptr, nullptr, nullptr};
    int     lo[3], hi[3];
    for (int d = 0; d < 3; ++d) {
        if (o.dim == 2 && d == 1) continue;
        REQ(vel[d]);
        U[d] = tmp.get(o); G[d] = tmp.get(o);
        for (int e = 0; e < 3; ++e) { lo[e] = o.domain.lo[e]; hi[e] = o.domain.hi[e] + (e == d ? 1 : 0); }
        copy3d(o, d, U[d], vel[d], lo, hi, true);
    }
    double* div = tmp.get(o);
    double* ph  = tmp.get(o);
    cudaEvent_t e0, e1;
    SB_CUDA(cudaEventCreate(&e0)); SB_CUDA(cudaEventCreate(&e1));
    o.levelDivergence(div, U);                                   // :293
    const double n0 = o.norm(div, s->s.opt.normType);            // :297
    if (init_div_norm) *init_div_norm = n0;
    SB_CUDA(cudaEventRecord(e0, o.st()));
    SolverStatus st = s->s.solve(ph, div, true, true, -1.0);     // :316
    SB_CUDA(cudaEventRecord(e1, o.st()));
    o.levelGradient(G, ph, true);                                // :328
    for (int d = 0; d < 3; ++d)
        if (U[d]) k::incr_valid(o.st(), o.lay, U[d], G[d], -1.0, d);  // :331-336
    if (final_div_norm) {                                        // :354-357
        o.levelDivergence(div, U);
        *final_div_norm = o.norm(div, s->s.opt.normType);
    }
    for (int d = 0; d < 3; ++d) {
        if (!U[d]) continue;
        for (int e = 0; e < 3; ++e) { lo[e] = o.domain.lo[e]; hi[e] = o.domain.hi[e] + (e == d ? 1 : 0); }
        copy3d(o, d, U[d], vel[d], lo, hi, false);
    }
    for (int e = 0; e < 3; ++e) { lo[e] = o.domain.lo[e]; hi[e] = o.domain.hi[e]; }
    if (phi) copy3d(o, -1, ph, phi, lo, hi, false);
    o.ctx->sync();
    if (p && phi) {                                              // :340-349  p += phi / projDt
        const long long n = o.domain.numPts();
        const double    s1 = 1.0 / proj_dt;
        for (long long i = 0; i < n; ++i) p[i] = p[i] + s1 * phi[i];
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    fillStatus(s, st, status, ms);
    SB_END
}

}  // extern "C"
