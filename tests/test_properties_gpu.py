"""Size-independent properties at sizes the CPU oracle cannot reach in seconds, up to
BASELINE.json's full S5 grid (1024 x 1024 x 256): linearity, null space, restriction/prolongation
identities, a manufactured solution, and the projection's divergence drop."""
import numpy as np
import pytest

import somar_b200 as sb

pytestmark = pytest.mark.gpu


def _op(ctx, nx, L=(16.0, 16.0, 1.0), box=(128, 128, 0), bf=16, relax=sb.RELAX_VERTLINE, ampl=None):
    nx = np.array(nx)
    dXi = np.array(L) / nx
    lo = np.array([0, 0, -nx[2]])
    hi = lo + nx - 1
    blo, bhi = sb.make_base_grids(lo, hi, box, (1, 1, 0), bf)
    kw = {}
    if ampl is not None:
        xmin = lo * dXi
        kw = dict(map_kind=sb.MAP_STRETCHED, map_xmin=xmin, map_xmax=xmin + np.array(L), map_ampl=ampl)
    return sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, relax_method=relax, **kw)


def test_linearity_and_null_space(ctx):
    nx = (256, 256, 64)
    op = _op(ctx, nx, L=(4.0, 4.0, 1.0), ampl=(0.0, 0.0, -0.1))
    rng = np.random.default_rng(3)
    x, y = rng.standard_normal(nx), rng.standard_normal(nx)
    fx, fy, fz, out = op.field(data=x), op.field(data=y), op.field(data=2.5 * x - 0.75 * y), op.field()
    op.applyOp(out, fx); Lx = out.download()
    op.applyOp(out, fy); Ly = out.download()
    op.applyOp(out, fz); Lz = out.download()
    scale = np.max(np.abs(Lx))
    assert np.max(np.abs(Lz - (2.5 * Lx - 0.75 * Ly))) <= 1e-12 * scale
    ones = op.field(data=np.ones(nx))
    op.applyOp(out, ones)
    assert np.max(np.abs(out.download())) <= 1e-9 * scale       # L[1] = 0 up to rounding of 1/(1/x)
    op.free()


def test_restrict_of_constant_prolong_is_identity(ctx):
    nx = (128, 128, 64)
    op = _op(ctx, nx, L=(2.0, 2.0, 1.0), box=(64, 64, 0))
    crse = op.new_mg_operator((2, 2, 2))
    rng = np.random.default_rng(5)
    c0 = rng.standard_normal((64, 64, 32))
    c, f, c2 = crse.field(data=c0), op.field(), crse.field()
    op.setToZero(f)
    # order-0 prolongation (+ removeKernel, which only shifts by the mean) then block averaging
    op.MGProlong(f, c, crse, 0)
    op.MGRestrict(c2, f, crse)
    got = c2.download()
    assert np.max(np.abs((got - got.mean()) - (c0 - c0.mean()))) <= 1e-13
    crse.free()
    op.free()


def test_manufactured_solution_256(ctx):
    nx = (256, 256, 64)
    L = (4.0, 4.0, 1.0)
    op = _op(ctx, nx, L=L)
    x = (np.arange(nx[0]) + 0.5) * L[0] / nx[0]
    y = (np.arange(nx[1]) + 0.5) * L[1] / nx[1]
    z = (np.arange(-nx[2], 0) + 0.5) / nx[2]
    k = 4 * 2 * np.pi / L[0]
    sol = np.cos(k * x)[:, None, None] * np.cos(k * y)[None, :, None] * (16 * z**2 * (z + 1) ** 2)[None, None, :]
    sol -= sol.mean()
    phi, rhs = op.field(data=sol), op.field()
    op.applyOp(rhs, phi)
    solver = sb.LevelHybridSolver(op, sb.default_options(relTol=1e-12, absTol=1e-14, maxIters=20))
    out = op.field()
    st = solver.solve(out, rhs)
    assert st.status in (1, 4)
    got = out.download()
    got -= got.mean()
    assert np.max(np.abs(got - sol)) <= 1e-8 * np.max(np.abs(sol))
    solver.free()
    op.free()


@pytest.mark.parametrize("nx", [(512, 512, 128), (1024, 1024, 256)])
def test_projection_reduces_divergence_full_size(ctx, nx):
    import bench
    nxa = np.array(nx)
    L = np.array([16.0, 16.0, 1.0]) * nxa / np.array([1024, 1024, 256])
    op = _op(ctx, nx, L=tuple(L))
    dXi = L / nxa
    lo = np.array([0, 0, -nx[2]])
    bench.NX = tuple(nx)
    res = bench.synthetic_residual(lo, lo + nxa - 1, dXi, 7)
    rhs, phi, chk = op.field(data=res), op.field(), op.field()
    del res
    solver = sb.LevelHybridSolver(op, sb.default_options())
    st = solver.solve(phi, rhs)
    assert st.status == 1                                  # CONVERGED within maxIters = 10 FMG cycles
    assert st.final_res_norm <= 1e-6 * st.norms[0]
    assert all(b < a for a, b in zip(st.norms[:-1], st.norms[1:]))   # monotone, like the reference's hang/diverge tests demand
    op.residual(chk, phi, rhs)
    assert abs(op.norm(chk, 2) - st.final_res_norm) <= 1e-9 * st.norms[0]
    solver.free()
    op.free()


@pytest.mark.parametrize("nx,periodic", [((34, 18, 13), (0, 0, 0)), ((31, 17, 9), (0, 0, 0)), ((64, 32, 256), (1, 0, 0)),
                                          ((48, 24, 100), (1, 1, 0)), ((16, 16, 3), (0, 0, 0)), ((256, 128, 64), (0, 0, 0))])
def test_line_relaxation_kernels_agree(ctx, nx, periodic, monkeypatch):
    """The three implementations of one line-relaxation sweep -- vertline_k (dgtsv's operation order),
    vertline_smem_k (shared-matrix Thomas) and the chunked sweeps on colour-split storage
    (sb_line.cu) -- must agree to rounding, including odd extents, ragged chunks and periodic sides."""
    nxa = np.array(nx)
    dXi = np.array([4.0, 2.0, 1.0]) / nxa
    lo = np.array([0, 0, -nx[2]])
    hi = lo + nxa - 1
    rng = np.random.default_rng(11)
    phi0, rhs0 = rng.standard_normal(nx), rng.standard_normal(nx)
    out = {}
    for kind in ("general", "smem", "split"):
        monkeypatch.setenv("SB_LINE_KERNEL", kind)
        xmin = lo * dXi
        op = sb.PoissonOp(ctx, lo, hi, dXi, lo[None, :], hi[None, :], periodic=periodic, relax_method=sb.RELAX_VERTLINE,
                          map_kind=sb.MAP_STRETCHED, map_xmin=xmin, map_xmax=xmin + np.array([4.0, 2.0, 1.0]),
                          map_ampl=(0.0, 0.0, -0.1))
        phi, rhs = op.field(data=phi0), op.field(data=rhs0)
        op.relax(phi, rhs, 3)
        out[kind] = phi.download()
        op.free()
    scale = np.max(np.abs(out["general"]))
    assert np.max(np.abs(out["smem"] - out["general"])) <= 1e-13 * scale
    assert np.max(np.abs(out["split"] - out["general"])) <= 1e-13 * scale


@pytest.mark.parametrize("relax,periodic", [(sb.RELAX_VERTLINE, (0, 0, 0)), (sb.RELAX_VERTLINE, (1, 1, 0)), (sb.RELAX_GSRB, (0, 0, 0))])
def test_fused_precond_vcycle_is_bitwise_the_two_calls(ctx, relax, periodic):
    """sb_solver_precond_vcycle fuses preCond(cor, res, 0) (and, in an all-Neumann / periodic
    problem, the null-space shift after each prolongation) into the layout conversion that starts
    a line relaxation; the arithmetic per cell is unchanged, so the result must be bit-identical
    to sb_op_precond followed by sb_solver_vcycle."""
    nx = np.array((64, 32, 32))
    dXi = np.array([4.0, 2.0, 1.0]) / nx
    lo = np.array([0, 0, -nx[2]])
    hi = lo + nx - 1
    blo, bhi = sb.make_base_grids(lo, hi, (32, 16, 0), (1, 1, 0), 8)
    op = sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, periodic=periodic, relax_method=relax)
    solver = sb.MGSolver(op, sb.default_options())
    r0 = np.random.default_rng(2).standard_normal(tuple(nx))
    r0 -= r0.mean()
    res, a, b = op.field(data=r0), op.field(), op.field()
    op.preCond(a, res, 0)
    solver.vcycle(a, res)
    solver.precond_vcycle(b, res)
    assert np.array_equal(a.download(), b.download())
    solver.free()
    op.free()


def test_async_copies_are_ordered_by_stream_wait(ctx):
    """sb_field_upload_async / sb_field_download_async on the copy streams, ordered against the compute
    stream with sb_context_stream_wait, give the same bytes as the synchronous calls."""
    import torch
    nx = (64, 32, 16)
    op = _op(ctx, nx, L=(4.0, 2.0, 1.0), box=(32, 16, 0), bf=8)
    n = int(np.prod(nx))
    src = torch.empty(n, dtype=torch.float64, pin_memory=True)
    dst = torch.empty(n, dtype=torch.float64, pin_memory=True)
    src.numpy()[:] = np.random.default_rng(9).standard_normal(n)
    a, b, ref = op.field(), op.field(), op.field()
    ref.upload_ptr(src.data_ptr())
    op.preCond(b, ref, 0)
    want = b.download()
    C_, H, D = sb.STREAM_COMPUTE, sb.STREAM_H2D, sb.STREAM_D2H
    for _ in range(3):
        ctx.stream_wait(H, C_)
        a.upload_ptr_async(src.data_ptr())
        ctx.stream_wait(C_, H)
        op.preCond(b, a, 0)
        ctx.stream_wait(D, C_)
        b.download_ptr_async(dst.data_ptr())
        ctx.stream_sync(D)
        assert np.array_equal(dst.numpy().reshape(nx, order="F"), want)
    op.free()


@pytest.mark.parametrize("nx,periodic,box", [((34, 18, 13), (0, 0, 0), (0, 0, 0)), ((64, 32, 32), (1, 1, 1), (16, 16, 16)),
                                              ((48, 24, 10), (1, 0, 0), (16, 8, 0)), ((31, 17, 9), (0, 1, 0), (0, 0, 0))])
def test_gsrb_split_storage_is_bitwise_the_natural_kernel(ctx, nx, periodic, box, monkeypatch):
    """Point GSRB on colour-split storage (gsrb_split_k) evaluates the expression of gsrb_k on the same
    operands in the same order: identical bits, including odd extents, z-split boxes and periodic sides."""
    nxa = np.array(nx)
    L = np.array([4.0, 2.0, 3.0])
    dXi = L / nxa
    lo = np.array([-3, 5, -nx[2]])
    hi = lo + nxa - 1
    blo, bhi = (sb.make_base_grids(lo, hi, box, (1, 1, 1), 1) if any(box) else (lo[None, :], hi[None, :]))
    rng = np.random.default_rng(12)
    phi0, rhs0 = rng.standard_normal(nx), rng.standard_normal(nx)
    out = {}
    for kind in ("natural", "split"):
        monkeypatch.setenv("SB_GSRB_KERNEL", kind)
        xmin = lo * dXi
        op = sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, periodic=periodic, relax_method=sb.RELAX_GSRB, map_kind=sb.MAP_STRETCHED,
                          map_xmin=xmin, map_xmax=xmin + L, map_ampl=(0.05, 0.02, -0.1))
        phi, rhs = op.field(data=phi0), op.field(data=rhs0)
        op.relax(phi, rhs, 3)
        out[kind] = phi.download()
        op.free()
    assert np.array_equal(out["natural"], out["split"])


@pytest.mark.parametrize("nx,periodic,box", [((512, 256, 64), (0, 0, 0), (128, 128, 0)), ((640, 128, 128), (1, 0, 0), (0, 0, 0)),
                                              ((500, 200, 32), (0, 1, 0), (0, 0, 0))])
def test_tma_line_kernel_is_bitwise_the_fused_kernel(ctx, nx, periodic, box, monkeypatch):
    """vertline_tma_k (persistent, warp-specialised, operands by TMA) runs vertline_fused_k's arithmetic operation for
    operation: identical bits, including a ragged last tile in x (TMA zero fill), periodic sides and several boxes."""
    nxa = np.array(nx)
    L = np.array([8.0, 4.0, 1.0])
    dXi = L / nxa
    lo = np.array([0, 0, -nx[2]])
    hi = lo + nxa - 1
    blo, bhi = (sb.make_base_grids(lo, hi, box, (1, 1, 0), 4) if any(box) else (lo[None, :], hi[None, :]))
    rng = np.random.default_rng(21)
    phi0, rhs0 = rng.standard_normal(nx), rng.standard_normal(nx)
    out = {}
    for tma in ("0", "1"):
        monkeypatch.setenv("SB_LINE_TMA", tma)
        xmin = lo * dXi
        op = sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, periodic=periodic, relax_method=sb.RELAX_VERTLINE, map_kind=sb.MAP_STRETCHED,
                          map_xmin=xmin, map_xmax=xmin + L, map_ampl=(0.0, 0.0, -0.1))
        phi, rhs = op.field(data=phi0), op.field(data=rhs0)
        op.relax(phi, rhs, 3)
        out[tma] = phi.download()
        op.free()
    assert np.array_equal(out["0"], out["1"])


@pytest.mark.parametrize("nx,periodic,ampl", [((128, 64, 64), (0, 0, 0), (0.05, 0.03, -0.1)), ((96, 40, 32), (1, 0, 0), (0.0, 0.04, 0.0)),
                                               ((130, 34, 128), (0, 0, 0), (0.06, 0.0, -0.05)), ((64, 64, 256), (0, 1, 0), (0.05, 0.0, -0.1))])
def test_mapped_grid_line_kernel_agrees_with_the_dgtsv_order_kernel(ctx, nx, periodic, ampl, monkeypatch):
    """Horizontally stretched maps: vertline_tma_k<GENERAL> (per-column factorisation recomputed in the kernel, chunked
    sweeps) against vertline_k (dgtsv's operation order, one thread per column) -- agreement to rounding."""
    nxa = np.array(nx)
    L = np.array([4.0, 2.0, 1.0])
    dXi = L / nxa
    lo = np.array([0, 0, -nx[2]])
    hi = lo + nxa - 1
    rng = np.random.default_rng(22)
    phi0, rhs0 = rng.standard_normal(nx), rng.standard_normal(nx)
    out = {}
    for kind in ("general", "auto"):
        if kind == "auto":
            monkeypatch.delenv("SB_LINE_KERNEL", raising=False)
        else:
            monkeypatch.setenv("SB_LINE_KERNEL", kind)
        xmin = lo * dXi
        if periodic[0] or periodic[1]:
            amp = tuple(0.0 if periodic[d] else ampl[d] for d in range(3))   # a periodic direction keeps its uniform map
        else:
            amp = ampl
        op = sb.PoissonOp(ctx, lo, hi, dXi, lo[None, :], hi[None, :], periodic=periodic, relax_method=sb.RELAX_VERTLINE,
                          map_kind=sb.MAP_STRETCHED, map_xmin=xmin, map_xmax=xmin + L, map_ampl=amp)
        phi, rhs = op.field(data=phi0), op.field(data=rhs0)
        op.relax(phi, rhs, 3)
        out[kind] = phi.download()
        op.free()
    scale = np.max(np.abs(out["general"]))
    assert np.max(np.abs(out["auto"] - out["general"])) <= 1e-12 * scale


def test_mapped_kernel_on_a_uniform_grid_agrees_with_the_shared_matrix_kernel(ctx, monkeypatch):
    """SB_LINE_KERNEL=mapped forces the per-column kernel where all columns do share one matrix: same mathematics."""
    nx = (256, 128, 64)
    nxa = np.array(nx)
    L = np.array([4.0, 2.0, 1.0])
    dXi = L / nxa
    lo = np.array([0, 0, -nx[2]])
    hi = lo + nxa - 1
    rng = np.random.default_rng(23)
    phi0, rhs0 = rng.standard_normal(nx), rng.standard_normal(nx)
    out = {}
    for kind in ("split", "mapped"):
        monkeypatch.setenv("SB_LINE_KERNEL", kind)
        xmin = lo * dXi
        op = sb.PoissonOp(ctx, lo, hi, dXi, lo[None, :], hi[None, :], relax_method=sb.RELAX_VERTLINE, map_kind=sb.MAP_STRETCHED,
                          map_xmin=xmin, map_xmax=xmin + L, map_ampl=(0.0, 0.0, -0.1))
        phi, rhs = op.field(data=phi0), op.field(data=rhs0)
        op.relax(phi, rhs, 4)
        out[kind] = phi.download()
        op.free()
    assert np.max(np.abs(out["mapped"] - out["split"])) <= 1e-13 * np.max(np.abs(out["split"]))
