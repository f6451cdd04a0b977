"""somar_b200 -- B200-native pressure projection for SOMAR (host-side Python mirror).

The compute lives in lib/libsomar_b200.so (CUDA, sm_100a) behind the C ABI of
include/somar_b200.h.  This package mirrors the reference's interface for the path, so tests and
benchmarks read like SOMAR code:

    PoissonOp            <- Grade3_Calculus/Elliptic/PoissonOp.H
    MGSolver             <- Grade3_Calculus/Elliptic/MGSolver.H
    LevelHybridSolver    <- Grade3_Calculus/Elliptic/LevelHybridSolver.H
    make_base_grids      <- Grade2_AnisotropicChombo/AnisotropicAMR.cpp:1461-1580 (makeBaseLevelMesh)

Arrays crossing this layer are numpy float64 in Fortran order over global index boxes.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import (CELL, FACE_X, FACE_Y, FACE_Z, MAP_CARTESIAN, MAP_STRETCHED, RELAX_GSRB, RELAX_VERTLINE, STATUS_NAMES,
                   MGOptions, SolverStatus, SomarB200Error)

__all__ = ["Context", "PoissonOp", "Field", "MGSolver", "LevelHybridSolver", "AMRHybridSolver", "make_base_grids",
           "assign_boxes_to_ranks", "default_options", "SomarB200Error"]


def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def make_base_grids(domain_lo, domain_hi, max_base_grid_size, split_dirs, block_factor):
    """Box list of AnisotropicAMR::makeBaseLevelMesh (AnisotropicAMR.cpp:1461-1580): the domain cut
    into near-equal boxes of at most max_base_grid_size (0 = unsplit) in the split directions.
    Returns (lo[n,3], hi[n,3]) in the reference's box order (x fastest)."""
    lo = np.asarray(domain_lo, dtype=np.int64)
    hi = np.asarray(domain_hi, dtype=np.int64)
    size = hi - lo + 1
    split = np.asarray(split_dirs, dtype=np.int64)
    mbgs = np.asarray(max_base_grid_size, dtype=np.int64).copy()
    for d in range(3):
        if mbgs[d] == 0 or mbgs[d] > size[d] or split[d] == 0:
            mbgs[d] = size[d]
    bf = np.where(split == 1, block_factor, 1)
    if np.any(lo % bf) or np.any(size % bf):
        raise ValueError("domain not coarsenable by blockFactor")
    blk_lo = lo // bf
    blk_sz = size // bf
    num = np.ones(3, dtype=np.int64)
    base = blk_sz.copy()
    for d in range(3):
        if split[d] == 0:
            continue
        bmax = mbgs[d] // block_factor
        if bmax <= 0:
            raise ValueError("maxBaseGridSize < blockFactor")
        nd = 1
        while nd * bmax < blk_sz[d]:
            nd += 1
        num[d] = nd
        base[d] = (blk_sz[d] + nd - 1) // nd
    los, his = [], []
    for k in range(num[2]):
        for j in range(num[1]):
            for i in range(num[0]):
                idx = np.array([i, j, k])
                b_lo = np.where(split == 1, (blk_lo + idx * base) * bf, lo)
                b_hi = np.where(split == 1, (np.minimum(blk_lo + idx * base + base - 1, blk_lo + blk_sz - 1) + 1) * bf - 1, hi)
                los.append(b_lo)
                his.append(b_hi)
    return np.array(los, dtype=np.int32), np.array(his, dtype=np.int32)


def process_grid(nranks, nbx, nby):
    """Px x Py with Px*Py = nranks dividing the box grid, as square as possible (x gets the larger)."""
    best = None
    for px in range(1, nranks + 1):
        if nranks % px:
            continue
        py = nranks // px
        if nbx % px or nby % py:
            continue
        # as square as possible; ties go to splitting y (x rows stay long and contiguous)
        score = abs(np.log((nbx / px) / max(nby / py, 1e-9))) + (1e-6 if px > py else 0.0)
        if best is None or score < best[0]:
            best = (score, px, py)
    if best is None:
        raise ValueError(f"cannot lay {nranks} ranks over a {nbx} x {nby} box grid")
    return best[1], best[2]


def assign_boxes_to_ranks(box_lo, box_hi, nranks):
    """Horizontal tile decomposition: ranks form a Px x Py grid over the (regular) box grid; each
    rank owns a rectangle of boxes.  Stands in for Chombo's LoadBalance (BoxTools/LoadBalance.cpp)."""
    xs = np.unique(box_lo[:, 0])
    ys = np.unique(box_lo[:, 1])
    px, py = process_grid(nranks, len(xs), len(ys))
    ix = np.searchsorted(xs, box_lo[:, 0]) // (len(xs) // px)
    iy = np.searchsorted(ys, box_lo[:, 1]) // (len(ys) // py)
    return (ix + px * iy).astype(np.int32)


def _level_desc(domain_lo, domain_hi, dXi, box_lo, box_hi, box_rank, periodic, dim, relax_method):
    box_lo = np.ascontiguousarray(box_lo, dtype=np.int32)
    box_hi = np.ascontiguousarray(box_hi, dtype=np.int32)
    box_rank = np.ascontiguousarray(box_rank if box_rank is not None else np.zeros(len(box_lo)), dtype=np.int32)
    d = capi.LevelDesc()
    d.dim = dim
    d.domain_lo, d.domain_hi, d.periodic = _i3(domain_lo), _i3(domain_hi), _i3(periodic)
    d.dXi = (C.c_double * 3)(*[float(v) for v in dXi])
    d.num_boxes = len(box_lo)
    d.box_lo = box_lo.ctypes.data_as(capi.IP)
    d.box_hi = box_hi.ctypes.data_as(capi.IP)
    d.box_rank = box_rank.ctypes.data_as(capi.IP)
    d.relax_method = relax_method
    d._keep = (box_lo, box_hi, box_rank)
    return d


def plan_tile(domain_lo, domain_hi, box_lo, box_hi, box_rank, rank, nranks, periodic=(0, 0, 0), dim=3):
    """Host-only (no GPU): the rectangle `rank` owns and what its six sides touch.
    Returns (tile_lo, tile_hi, side_kind[6], side_neighbor[6], num_local_boxes); side index = 2*dir + side,
    kind 0 physical, 1 periodic onto itself, 2 neighbour rank."""
    d = _level_desc(domain_lo, domain_hi, (1, 1, 1), box_lo, box_hi, box_rank, periodic, dim, RELAX_VERTLINE)
    lo, hi, kind, nb, n = (C.c_int * 3)(), (C.c_int * 3)(), (C.c_int * 6)(), (C.c_int * 6)(), C.c_int()
    capi.check(capi.load().sb_plan_tile(C.byref(d), rank, nranks, lo, hi, kind, nb, C.byref(n)))
    return list(lo), list(hi), list(kind), list(nb), n.value


def plan_cf_stencils(dom_lo, dom_hi, periodic, ref, fine_lo, fine_hi, box, direction, side):
    """Host-only: stencil records of the quadratic coarse-fine ghost interpolation for one side of one fine box
    (include/somar_b200.h: sb_plan_cf_stencils).  Returns (cells[n,3], w_first[n,2,5], w_second[n,2,5], w_mixed[n,3,3])."""
    fl = np.ascontiguousarray(fine_lo, dtype=np.int32)
    fh = np.ascontiguousarray(fine_hi, dtype=np.int32)
    lib = capi.load()
    n = C.c_int()
    args = (_i3(dom_lo), _i3(dom_hi), _i3(periodic), _i3(ref), len(fl), fl.ctypes.data_as(capi.IP), fh.ctypes.data_as(capi.IP), box,
            direction, side)
    capi.check(lib.sb_plan_cf_stencils(*args, 0, C.byref(n), None, None, None, None))
    cells = np.zeros((n.value, 3), dtype=np.int32)
    w1, w2, wm = np.zeros((n.value, 2, 5)), np.zeros((n.value, 2, 5)), np.zeros((n.value, 3, 3))
    if n.value:
        capi.check(lib.sb_plan_cf_stencils(*args, n.value, C.byref(n), cells.ctypes.data_as(capi.IP), w1.ctypes.data_as(capi.DP),
                                           w2.ctypes.data_as(capi.DP), wm.ctypes.data_as(capi.DP)))
    return cells, w1, w2, wm


def plan_schedule(domain_lo, domain_hi, dXi, box_lo, box_hi, relax_method=RELAX_VERTLINE, max_depth=-1, dim=3):
    """Host-only: the MG refinement schedule MGSolver::define would build (MGCoarseningStrategy.cpp)."""
    d = _level_desc(domain_lo, domain_hi, dXi, box_lo, box_hi, None, (0, 0, 0), dim, relax_method)
    n = C.c_int()
    buf = (C.c_int * (3 * 64))()
    capi.check(capi.load().sb_plan_schedule(C.byref(d), max_depth, buf, 64, C.byref(n)))
    return [tuple(buf[3 * i:3 * i + 3]) for i in range(n.value)]


def default_options(**kw):
    """MGSolver<T>::getDefaultOptions with the reference's proj.* defaults (ProjectorParameters.cpp:124-222)."""
    o = MGOptions()
    capi.load().sb_mg_default_options(C.byref(o))
    for k, v in kw.items():
        if k.startswith("bottom_"):
            setattr(o.bottom, k[len("bottom_"):], v)
        else:
            if not hasattr(o, k):
                raise AttributeError(k)
            setattr(o, k, v)
    return o


class Context:
    """One per process / GPU."""

    def __init__(self, device=0, rank=0, nranks=1):
        self.lib = capi.load()
        self.h = C.c_void_p()
        capi.check(self.lib.sb_context_create(C.byref(self.h), device, rank, nranks))
        self.rank, self.nranks = rank, nranks

    def init_comm(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        capi.check(self.lib.sb_comm_init(self.h, buf))

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        capi.check(capi.load().sb_comm_get_unique_id(buf))
        return buf.raw

    def sync(self):
        capi.check(self.lib.sb_context_sync(self.h))

    def stream_wait(self, waiter, signaller):
        """Work enqueued later on stream `waiter` starts after everything enqueued so far on `signaller`."""
        capi.check(self.lib.sb_context_stream_wait(self.h, waiter, signaller))

    def stream_sync(self, which):
        capi.check(self.lib.sb_context_stream_sync(self.h, which))

    def timer_start(self):
        capi.check(self.lib.sb_context_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        capi.check(self.lib.sb_context_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile(self, enable=True):
        capi.check(self.lib.sb_context_profile(self.h, int(enable)))

    def profile_get(self, key):
        ms, n = C.c_double(), C.c_longlong()
        capi.check(self.lib.sb_context_profile_get(self.h, key.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self):
        return int(self.lib.sb_context_launch_count(self.h))

    def close(self):
        if self.h:
            self.lib.sb_context_destroy(self.h)
            self.h = C.c_void_p()


class Field:
    """LevelData<FArrayBox> (centering CELL) or one direction of a LevelData<FluxBox>."""

    def __init__(self, op, centering=CELL):
        self.op, self.centering = op, centering
        self.lib = op.lib
        self.h = C.c_void_p()
        capi.check(self.lib.sb_field_create(op.h, centering, C.byref(self.h)))

    def box(self, ghost=0):
        """The box this level's data spans: the domain on a base level, the refined patch otherwise."""
        lo = np.array(self.op.patch_lo) - ghost
        hi = np.array(self.op.patch_hi) + ghost
        if self.centering >= 0:
            hi[self.centering] += 1
        return lo, hi

    def upload_ptr(self, ptr, lo=None, hi=None):
        """Upload from a raw host pointer (e.g. a pinned torch tensor) laid out as the box [lo, hi]."""
        if lo is None:
            lo, hi = self.box()
        capi.check(self.lib.sb_field_upload(self.h, C.cast(ptr, capi.DP), _i3(lo), _i3(hi)))

    def download_ptr(self, ptr, lo=None, hi=None):
        if lo is None:
            lo, hi = self.box()
        capi.check(self.lib.sb_field_download(self.h, C.cast(ptr, capi.DP), _i3(lo), _i3(hi)))

    def upload_ptr_async(self, ptr, lo=None, hi=None):
        """Asynchronous upload from pinned host memory on the context's H2D stream."""
        if lo is None:
            lo, hi = self.box()
        capi.check(self.lib.sb_field_upload_async(self.h, C.cast(ptr, capi.DP), _i3(lo), _i3(hi)))

    def download_ptr_async(self, ptr, lo=None, hi=None):
        """Asynchronous download into pinned host memory on the context's D2H stream."""
        if lo is None:
            lo, hi = self.box()
        capi.check(self.lib.sb_field_download_async(self.h, C.cast(ptr, capi.DP), _i3(lo), _i3(hi)))

    def upload(self, arr, lo=None, hi=None):
        if lo is None:
            lo, hi = self.box()
        shape = tuple(int(h - l + 1) for l, h in zip(lo, hi))
        a = np.asfortranarray(arr, dtype=np.float64).reshape(shape, order="F")
        capi.check(self.lib.sb_field_upload(self.h, a.ctypes.data_as(capi.DP), _i3(lo), _i3(hi)))
        return self

    def download(self, lo=None, hi=None):
        if lo is None:
            lo, hi = self.box()
        shape = tuple(int(h - l + 1) for l, h in zip(lo, hi))
        a = np.zeros(shape, dtype=np.float64, order="F")
        capi.check(self.lib.sb_field_download(self.h, a.ctypes.data_as(capi.DP), _i3(lo), _i3(hi)))
        return a

    def free(self):
        if self.h:
            self.lib.sb_field_destroy(self.h)
            self.h = C.c_void_p()


class PoissonOp:
    """Elliptic::PoissonOp for the projection: J(alpha + beta Lap) with HomogNeumBC by default."""

    def __init__(self, ctx, domain_lo, domain_hi, dXi, box_lo, box_hi, box_rank=None, periodic=(0, 0, 0), dim=3,
                 map_kind=MAP_CARTESIAN, map_xmin=(0, 0, 0), map_xmax=(1, 1, 1), map_ampl=(0, 0, 0), bc_alpha=None, bc_beta=None,
                 alpha=0.0, beta=1.0, relax_method=RELAX_VERTLINE, finalize=True, crse_op=None, _handle=None):
        """crse_op: the PoissonOp of the next coarser AMR level (a_crseGrids of PoissonOp.cpp:33-43); the boxes then
        describe a refined patch of a domain that is a refinement of crse_op's."""
        self.ctx, self.lib = ctx, ctx.lib
        self.dim = dim
        self.domain_lo, self.domain_hi = tuple(int(v) for v in domain_lo), tuple(int(v) for v in domain_hi)
        self.patch_lo, self.patch_hi = self.domain_lo, self.domain_hi
        self.dXi = tuple(float(v) for v in dXi)
        self.crse_op = crse_op
        if _handle is not None:
            self.h = _handle
            return
        self.box_lo = np.ascontiguousarray(box_lo, dtype=np.int32)
        self.box_hi = np.ascontiguousarray(box_hi, dtype=np.int32)
        n = self.box_lo.shape[0]
        self.box_rank = np.zeros(n, dtype=np.int32) if box_rank is None else np.ascontiguousarray(box_rank, dtype=np.int32)
        d = capi.LevelDesc()
        d.dim = dim
        d.domain_lo, d.domain_hi, d.periodic = _i3(domain_lo), _i3(domain_hi), _i3(periodic)
        d.dXi = (C.c_double * 3)(*self.dXi)
        d.num_boxes = n
        d.box_lo = self.box_lo.ctypes.data_as(capi.IP)
        d.box_hi = self.box_hi.ctypes.data_as(capi.IP)
        d.box_rank = self.box_rank.ctypes.data_as(capi.IP)
        d.map_kind = map_kind
        d.map_xmin = (C.c_double * 3)(*map_xmin)
        d.map_xmax = (C.c_double * 3)(*map_xmax)
        d.map_ampl = (C.c_double * 3)(*map_ampl)
        for i in range(3):
            for s in range(2):
                d.bc_alpha[i][s] = 0.0 if bc_alpha is None else bc_alpha[i][s]
                d.bc_beta[i][s] = 1.0 if bc_beta is None else bc_beta[i][s]
        d.alpha, d.beta, d.relax_method = alpha, beta, relax_method
        if crse_op is not None:
            d.num_crse_boxes = len(crse_op.box_lo)
            d.crse_box_lo = crse_op.box_lo.ctypes.data_as(capi.IP)
            d.crse_box_hi = crse_op.box_hi.ctypes.data_as(capi.IP)
            d.crse_box_rank = crse_op.box_rank.ctypes.data_as(capi.IP)
            d.crse_domain_lo, d.crse_domain_hi = _i3(crse_op.domain_lo), _i3(crse_op.domain_hi)
            self.patch_lo = tuple(int(v) for v in self.box_lo.min(axis=0))
            self.patch_hi = tuple(int(v) for v in self.box_hi.max(axis=0))
        self.h = C.c_void_p()
        capi.check(self.lib.sb_op_create(ctx.h, C.byref(d), C.byref(self.h)))
        if finalize:
            self.finalize()

    # -- construction helpers
    def finalize(self):
        capi.check(self.lib.sb_op_finalize(self.h))

    def set_metric(self, centering, box_id, arr, lo, hi):
        a = np.asfortranarray(arr, dtype=np.float64)
        capi.check(self.lib.sb_op_set_metric(self.h, centering, box_id, a.ctypes.data_as(capi.DP), _i3(lo), _i3(hi)))

    def new_mg_operator(self, ref):
        h = C.c_void_p()
        capi.check(self.lib.sb_op_new_mg_operator(self.h, _i3(ref), C.byref(h)))
        lo, hi, dxi = (C.c_int * 3)(), (C.c_int * 3)(), (C.c_double * 3)()
        capi.check(self.lib.sb_op_get_info(h, lo, hi, dxi, None))
        return PoissonOp(self.ctx, list(lo), list(hi), list(dxi), None, None, dim=self.dim, _handle=h)

    @property
    def has_null_space(self):
        v = C.c_int()
        capi.check(self.lib.sb_op_has_null_space(self.h, C.byref(v)))
        return bool(v.value)

    def halo_mode(self):
        """How relaxations exchange face ghosts with neighbouring tiles: 'none', 'nccl' or 'peer' (sb_op_halo_mode)."""
        v = C.c_int()
        capi.check(self.lib.sb_op_halo_mode(self.h, C.byref(v)))
        return ("none", "nccl", "peer")[v.value]

    def coefficient(self, which):
        """0 J, 1 Dinv, 2..4 M_d (2*N_d), 5..7 Jgup_d over the (face) domain box."""
        n = np.array(self.domain_hi) - np.array(self.domain_lo) + 1
        if which in (2, 3, 4):
            out = np.zeros(2 * n[which - 2])
        else:
            shape = n.copy()
            if which >= 5:
                shape[which - 5] += 1
            out = np.zeros(tuple(shape), order="F")
        capi.check(self.lib.sb_op_get_coefficient(self.h, which, out.ctypes.data_as(capi.DP), out.size))
        return out

    def field(self, centering=CELL, data=None):
        f = Field(self, centering)
        if data is not None:
            f.upload(data)
        return f

    def flux(self, data=None):
        fs = [Field(self, d) if not (self.dim == 2 and d == 1) else None for d in range(3)]
        if data is not None:
            for f, a in zip(fs, data):
                if f is not None:
                    f.upload(a)
        return fs

    @staticmethod
    def _f3(fs):
        return (C.c_void_p * 3)(*[f.h if f is not None else None for f in fs])

    # -- LevelOperator / MGOperator / StateOps surface
    def applyBCs(self, phi, homog=True):
        capi.check(self.lib.sb_op_apply_bcs(self.h, phi.h, int(homog)))

    def applyOp(self, lhs, phi, homog=True):
        capi.check(self.lib.sb_op_apply_op(self.h, lhs.h, phi.h, int(homog)))

    def residual(self, res, phi, rhs, homog=True):
        capi.check(self.lib.sb_op_residual(self.h, res.h, phi.h, rhs.h, int(homog)))

    def relax(self, cor, res, iters):
        capi.check(self.lib.sb_op_relax(self.h, cor.h, res.h, iters))

    def preCond(self, phi, rhs, relax_iters=0):
        capi.check(self.lib.sb_op_precond(self.h, phi.h, rhs.h, relax_iters))

    def removeKernel(self, phi):
        capi.check(self.lib.sb_op_remove_kernel(self.h, phi.h))

    def norm(self, x, p=2, pow_scale=1.0):
        v = C.c_double()
        capi.check(self.lib.sb_op_norm(self.h, x.h, p, pow_scale, C.byref(v)))
        return v.value

    def dotProduct(self, a, b):
        v = C.c_double()
        capi.check(self.lib.sb_op_dot(self.h, a.h, b.h, C.byref(v)))
        return v.value

    def incr(self, lhs, x, scale):
        capi.check(self.lib.sb_op_incr(self.h, lhs.h, x.h, scale))

    def axby(self, lhs, x, y, a, b):
        capi.check(self.lib.sb_op_axby(self.h, lhs.h, x.h, y.h, a, b))

    def scale(self, lhs, s):
        capi.check(self.lib.sb_op_scale(self.h, lhs.h, s))

    def setToZero(self, lhs):
        capi.check(self.lib.sb_op_set_to_zero(self.h, lhs.h))

    def assignLocal(self, dst, src):
        capi.check(self.lib.sb_op_assign_local(self.h, dst.h, src.h))

    def MGRestrict(self, crse_res, fine_res, crse_op):
        capi.check(self.lib.sb_op_mg_restrict(self.h, crse_op.h, crse_res.h, fine_res.h))

    def MGProlong(self, fine_phi, crse_cor, crse_op, order):
        capi.check(self.lib.sb_op_mg_prolong(self.h, crse_op.h, fine_phi.h, crse_cor.h, order))

    def levelDivergence(self, div, vel):
        capi.check(self.lib.sb_op_level_divergence(self.h, div.h, self._f3(vel)))

    def levelGradient(self, grad, phi, homog=True):
        capi.check(self.lib.sb_op_level_gradient(self.h, self._f3(grad), phi.h, int(homog)))

    # -- AMRMGOperator surface (Elliptic/AMRMGOperator.H:43-218, PoissonOp.cpp:1156-1478)
    def applyBCsAMR(self, phi, crse_phi=None, homog_phys=True, homog_cfi=True):
        capi.check(self.lib.sb_op_apply_bcs_amr(self.h, phi.h, crse_phi.h if crse_phi else None, int(homog_phys), int(homog_cfi)))

    def AMROperatorNF(self, lhs, phi, phi_crse, homog_phys=True):
        capi.check(self.lib.sb_op_amr_operator_nf(self.h, lhs.h, phi.h, phi_crse.h, int(homog_phys)))

    def AMROperatorNC(self, lhs, phi_fine, phi, finer_op, homog_phys=True):
        capi.check(self.lib.sb_op_amr_operator_nc(self.h, lhs.h, phi_fine.h, phi.h, int(homog_phys), finer_op.h))

    def AMROperator(self, lhs, phi_fine, phi, phi_crse, finer_op, homog_phys=True):
        capi.check(self.lib.sb_op_amr_operator(self.h, lhs.h, phi_fine.h, phi.h, phi_crse.h, int(homog_phys), finer_op.h))

    def AMRResidual(self, res, phi, rhs, phi_fine=None, finer_op=None, phi_crse=None, homog_phys=True):
        """AMRResidual / AMRResidualNF (no finer level) / AMRResidualNC (no coarser level)."""
        capi.check(self.lib.sb_op_amr_residual(self.h, res.h, phi_fine.h if phi_fine else None, phi.h, phi_crse.h if phi_crse else None,
                                               rhs.h, int(homog_phys), finer_op.h if finer_op else None))

    def AMRNormLevel(self, res, finer_op=None, p=2):
        v = C.c_double()
        capi.check(self.lib.sb_op_amr_norm_level(self.h, res.h, finer_op.h if finer_op else None, p, C.byref(v)))
        return v.value

    def getFlux(self, flux, phi):
        capi.check(self.lib.sb_op_get_flux(self.h, self._f3(flux), phi.h))

    def reflux(self, res, fine_phi, phi, finer_op):
        capi.check(self.lib.sb_op_reflux(self.h, res.h, fine_phi.h, phi.h, finer_op.h))

    def compDivergence(self, div, flux, fine_flux=None, finer_op=None):
        ff = self._f3(fine_flux) if fine_flux is not None else None
        capi.check(self.lib.sb_op_comp_divergence(self.h, div.h, self._f3(flux), ff, finer_op.h if finer_op else None))

    def averageDown(self, crse, fine):
        """CFInterp::coarsen: this (fine) level's data averaged onto the coarser level's covered cells."""
        capi.check(self.lib.sb_op_average_down(self.h, crse.h, fine.h))

    def compGradient(self, grad, phi, crse_phi=None, homog_phys=True, homog_cfi=True):
        capi.check(self.lib.sb_op_comp_gradient(self.h, self._f3(grad), phi.h, crse_phi.h if crse_phi else None, int(homog_phys),
                                                int(homog_cfi)))

    def sendToAdvectingVelocity(self, vel, ghost=1):
        """AMRNSLevel::sendToAdvectingVelocity (AMRNSLevelFill.cpp:194-232), in place on device face fields."""
        capi.check(self.lib.sb_op_send_to_advecting_velocity(self.h, self._f3(vel), ghost))

    def sendToCartesianVelocity(self, vel, ghost=1):
        """AMRNSLevel::sendToCartesianVelocity (AMRNSLevelFill.cpp:238-280)."""
        capi.check(self.lib.sb_op_send_to_cartesian_velocity(self.h, self._f3(vel), ghost))

    def fluxIncr(self, vel, grad, scale=1.0):
        capi.check(self.lib.sb_op_flux_incr(self.h, self._f3(vel), self._f3(grad), scale))

    def free(self):
        if self.h:
            self.lib.sb_op_destroy(self.h)
            self.h = C.c_void_p()


STREAM_COMPUTE, STREAM_H2D, STREAM_D2H = 0, 1, 2  # include/somar_b200.h SB_STREAM_*


class _SolverBase:
    def __init__(self, op):
        self.op, self.lib = op, op.lib
        self.h = C.c_void_p()

    @property
    def schedule(self):
        n = C.c_int()
        capi.check(self.lib.sb_solver_get_schedule(self.h, None, 0, C.byref(n)))
        buf = (C.c_int * (3 * n.value))()
        capi.check(self.lib.sb_solver_get_schedule(self.h, buf, n.value, C.byref(n)))
        return [tuple(buf[3 * i:3 * i + 3]) for i in range(n.value)]

    def solve(self, phi, rhs, homog=True, set_phi_to_zero=True, convergence_metric=-1.0):
        st = SolverStatus()
        capi.check(self.lib.sb_solver_solve(self.h, phi.h, rhs.h, int(homog), int(set_phi_to_zero), convergence_metric, C.byref(st)))
        return st

    def vcycle(self, cor, res):
        capi.check(self.lib.sb_solver_vcycle(self.h, cor.h, res.h))

    def precond_vcycle(self, cor, res):
        """op.preCond(cor, res, 0); vCycle_residualEq(cor, res, 0) -- MGSolverI.H:342-347."""
        capi.check(self.lib.sb_solver_precond_vcycle(self.h, cor.h, res.h))

    def set_options(self, opt):
        capi.check(self.lib.sb_solver_set_options(self.h, C.byref(opt)))

    def project_correct(self, vel, p=None, proj_dt=1.0, vel_ghost=-1, phi_out=None):
        """AMRNSLevel::projectCorrect on device-resident fields: returns (initDivNorm, finalDivNorm, status)."""
        n0, n1, st = C.c_double(), C.c_double(), SolverStatus()
        capi.check(self.lib.sb_project_correct(self.h, PoissonOp._f3(vel), p.h if p else None, proj_dt, vel_ghost,
                                               phi_out.h if phi_out else None, C.byref(n0), C.byref(n1), C.byref(st)))
        return n0.value, n1.value, st

    def project_predict(self, vel, p, proj_dt=1.0, vel_ghost=-1):
        """AMRNSLevel::projectPredict on device-resident fields: returns ((initial, lagged, corrected) norms, used_fallback, status)."""
        norms, fb, st = (C.c_double * 3)(), C.c_int(), SolverStatus()
        capi.check(self.lib.sb_project_predict(self.h, PoissonOp._f3(vel), p.h, proj_dt, vel_ghost, norms, C.byref(fb), C.byref(st)))
        return tuple(norms), bool(fb.value), st

    def project_host(self, vel, proj_dt=1.0, p=None):
        """AMRNSLevel::projectCorrect with host arrays: returns (vel_out, phi, initDivNorm, finalDivNorm, status)."""
        op = self.op
        n = np.array(op.domain_hi) - np.array(op.domain_lo) + 1
        vs, ptrs = [], []
        for d in range(3):
            if op.dim == 2 and d == 1:
                vs.append(None)
                ptrs.append(None)
                continue
            shape = n.copy()
            shape[d] += 1
            a = np.array(vel[d], dtype=np.float64, order="F").reshape(tuple(shape), order="F")
            vs.append(a)
            ptrs.append(a.ctypes.data_as(capi.DP))
        phi = np.zeros(tuple(n), order="F")
        arr = (capi.DP * 3)(*[q if q is not None else capi.DP() for q in ptrs])
        n0, n1, st = C.c_double(), C.c_double(), SolverStatus()
        pp = p.ctypes.data_as(capi.DP) if p is not None else None
        capi.check(self.lib.sb_project_host(self.h, arr, phi.ctypes.data_as(capi.DP), pp, proj_dt, C.byref(n0), C.byref(n1), C.byref(st)))
        return vs, phi, n0.value, n1.value, st

    def free(self):
        if self.h:
            self.lib.sb_solver_destroy(self.h)
            self.h = C.c_void_p()


class AMRHybridSolver:
    """Elliptic::AMRHybridSolver::define(vAMRMGOps, lmin, lmax, opts) / solve(vphi, vrhs, ...)."""

    def __init__(self, ops, lmin=0, lmax=None, opt=None):
        self.ops, self.lib = list(ops), ops[-1].lib
        self.lmin, self.lmax = lmin, len(ops) - 1 if lmax is None else lmax
        opt = opt or default_options()
        arr = (C.c_void_p * len(ops))(*[o.h if o is not None else None for o in ops])
        self.h = C.c_void_p()
        capi.check(self.lib.sb_amr_solver_create(arr, len(ops), self.lmin, self.lmax, C.byref(opt), C.byref(self.h)))

    def solve(self, phi, rhs, homog=True, set_phi_to_zero=True, convergence_metric=-1.0):
        n = len(self.ops)
        pa = (C.c_void_p * n)(*[f.h if f is not None else None for f in phi])
        ra = (C.c_void_p * n)(*[f.h if f is not None else None for f in rhs])
        st = SolverStatus()
        capi.check(self.lib.sb_amr_solver_solve(self.h, pa, ra, int(homog), int(set_phi_to_zero), convergence_metric, C.byref(st)))
        return st

    def free(self):
        if self.h:
            self.lib.sb_amr_solver_destroy(self.h)
            self.h = C.c_void_p()


class MGSolver(_SolverBase):
    """Elliptic::MGSolver<LevelData<FArrayBox>>::define(topOp, opt, refSchedule)."""

    def __init__(self, op, opt=None, schedule=None):
        super().__init__(op)
        opt = opt or default_options()
        if schedule:
            flat = (C.c_int * (3 * len(schedule)))(*[int(v) for r in schedule for v in r])
            capi.check(self.lib.sb_mgsolver_create(op.h, C.byref(opt), flat, len(schedule), C.byref(self.h)))
        else:
            capi.check(self.lib.sb_mgsolver_create(op.h, C.byref(opt), None, 0, C.byref(self.h)))


class LevelHybridSolver(_SolverBase):
    """Elliptic::LevelHybridSolver::define(mgOp, opts): picks MG, Leptic or Leptic_MG by lepticity
    (LevelHybridSolver.cpp:457-498); status.solve_mode says which."""

    def __init__(self, op, opt=None):
        super().__init__(op)
        opt = opt or default_options()
        capi.check(self.lib.sb_hybrid_solver_create(op.h, C.byref(opt), C.byref(self.h)))
