"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle (the reference's
own C++ solver stack, oracle/_ref) run live on the same seeded inputs.

Tolerances (BASELINE.json north_star): per-cycle residual norms 1e-10 relative, final pressure and
projected velocity 1e-9 relative in max-norm, identical cycle counts and solver status.  The
per-cycle norms are compared the way the solver itself monitors them, as relative residuals
|r_i|/|r_0| (MGSolverI.H:361): |norm_i - ref_i| <= 1e-10 |r_0|, and additionally to 1e-6 of their
own value (a residual that has dropped 8 orders of magnitude is itself only defined to ~1e-8
relative in fp64, whatever the summation order; and rhs - L[phi] has a rounding floor of about
eps * |L| * |phi| -- 5e-11 |r_0| on the strongly anisotropic DJL grid -- below which a norm is
noise, hence the absolute term 1e-11 |r_0|, ten times tighter than the north-star bound).  The
element-wise kernels are written to agree far more tightly; those bounds are stated per test."""


def assert_norms(got, ref):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape
    assert np.all(np.abs(got - ref) <= 1e-10 * ref[0]), (got, ref)
    assert np.all(np.abs(got - ref) <= 1e-6 * ref + 1e-11 * ref[0]), (got, ref)
import numpy as np
import pytest

import somar_b200 as sb
from _oracle import have_ref, run_ref
from cases import CASES, make_op, rand_field, rand_velocity, ref_kwargs, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_ref(3), reason="oracle/_ref/d3/somar_ref not built")]

ALL = sorted(CASES)


@pytest.mark.parametrize("name", ALL)
def test_coefficients(ctx, name):
    c = CASES[name]
    op = make_op(ctx, c)
    ref = run_ref("applyop", inp=[rand_field(c, 1)], **ref_kwargs(c))
    n = np.array(c["nx"])
    assert op.has_null_space == bool(ref.kv["hasNullSpace"])
    assert rel_err(op.coefficient(0), ref["J"]) <= 4e-16
    assert rel_err(op.coefficient(1), ref["Dinv"]) <= 1e-15
    for d in range(3):
        assert rel_err(op.coefficient(2 + d), ref[f"M{d}"]) == 0.0
        assert rel_err(op.coefficient(5 + d), ref[f"Jgup{d}"]) <= 4e-16
    op.free()


@pytest.mark.parametrize("name", ALL)
def test_apply_op_and_norms(ctx, name):
    c = CASES[name]
    op = make_op(ctx, c)
    phi0 = rand_field(c, 1)
    ref = run_ref("applyop", inp=[phi0], **ref_kwargs(c))
    phi, lhs = op.field(data=phi0), op.field()
    op.applyOp(lhs, phi)
    got = lhs.download()
    assert rel_err(got, ref["lhs"]) <= 1e-14
    assert abs(op.norm(lhs, 2) - ref.kv["norm2"]) <= 1e-13 * ref.kv["norm2"]
    assert abs(op.norm(lhs, 0) - ref.kv["norm0"]) <= 1e-14 * ref.kv["norm0"]
    op.free()


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("iters", [1, 3])
def test_relax(ctx, name, iters):
    c = CASES[name]
    op = make_op(ctx, c)
    phi0, rhs0 = rand_field(c, 2), rand_field(c, 3, zero_mean=True)
    ref = run_ref("relax", inp=[phi0, rhs0], extra={"drv.relaxIters": iters}, **ref_kwargs(c))
    phi, rhs = op.field(data=phi0), op.field(data=rhs0)
    op.relax(phi, rhs, iters)
    assert rel_err(phi.download(), ref["phi"]) <= 1e-12
    op.free()


V_OPTS = dict(numCycles=1, numSmoothDown=2, numSmoothUp=2, numSmoothBottom=2, prolongOrder=1, maxIters=20, relTol=1e-10)


def _proj_overrides(o):
    return {f"proj.{k}": v for k, v in o.items()}


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("optset", ["defaults", "vcycle"])
def test_solve(ctx, name, optset):
    c = CASES[name]
    op = make_op(ctx, c)
    rhs0 = rand_field(c, 4, zero_mean=True)
    over = {} if optset == "defaults" else V_OPTS
    ref = run_ref("solve", inp=[rhs0], extra=_proj_overrides(over), **ref_kwargs(c))
    assert int(ref.kv["solveMode"]) == 1
    solver = sb.LevelHybridSolver(op, sb.default_options(**over))
    phi, rhs = op.field(), op.field(data=rhs0)
    st = solver.solve(phi, rhs)
    assert st.max_depth == int(ref.kv["maxDepth"])
    assert st.status == int(ref.kv["status"])
    # reference norm() calls at depth 0: hybrid's |res|, then MGSolver's initial and per-cycle norms
    ref_norms = ref["norms"][1:]
    assert st.num_norms == len(ref_norms)
    assert_norms(st.norms, ref_norms)
    assert rel_err(phi.download(), ref["phi"]) <= 1e-9
    solver.free()
    op.free()


@pytest.mark.parametrize("name", ["line_cart", "line_stretch", "line_aniso", "gsrb_stretch", "line_perx", "s_line64", "s_line128", "s_line256", "s_line64_xystretch"])
def test_project(ctx, name):
    c = CASES[name]
    op = make_op(ctx, c)
    vel0 = rand_velocity(c, 5)
    ref = run_ref("project", inp=vel0, **ref_kwargs(c))
    solver = sb.LevelHybridSolver(op, sb.default_options())
    vel, phi, n0, n1, st = solver.project_host(vel0)
    assert abs(n0 - ref.kv["initDivNorm"]) <= 1e-13 * ref.kv["initDivNorm"]
    assert st.status == int(ref.kv["status"])
    assert_norms(st.norms, ref["norms"][1:])
    assert rel_err(phi, ref["phi"]) <= 1e-9
    for d in range(3):
        assert rel_err(vel[d], ref[f"vel{d}"]) <= 1e-9
    assert abs(n1 - ref.kv["finalDivNorm"]) <= 1e-6 * max(ref.kv["finalDivNorm"], 1e-30) + 1e-12 * n0
    solver.free()
    op.free()


@pytest.mark.parametrize("name", ["line_stretch", "line_aniso", "gsrb_stretch", "line_cart"])
@pytest.mark.parametrize("ghost", [0, 1])
def test_velocity_transforms(ctx, name, ghost):
    """AMRNSLevel::sendToAdvectingVelocity / sendToCartesianVelocity (AMRNSLevelFill.cpp:194-280) on
    device-resident face fields, against the reference's GeoSourceInterface::fill_dxdXi +
    FArrayBox::mult / divide (oracle driver, mode transform): same tables, same two roundings per
    face, so the match is exact."""
    c = CASES[name]
    op = make_op(ctx, c)
    vel0 = rand_velocity(c, 6)
    ref = run_ref("transform", inp=vel0, extra={"drv.velGhost": ghost}, **ref_kwargs(c))
    vel = [op.field(centering=d, data=vel0[d]) for d in range(3)]
    op.sendToAdvectingVelocity(vel, ghost)
    for d in range(3):
        assert rel_err(vel[d].download(), ref[f"adv{d}"]) == 0.0
    op.sendToCartesianVelocity(vel, ghost)
    for d in range(3):
        assert rel_err(vel[d].download(), ref[f"cart{d}"]) == 0.0
    op.free()


def test_device_resident_projection_chain(ctx):
    """toAdvecting -> levelDivergence -> solve -> levelGradient -> vel -= grad -> toCartesian with every
    field on the device gives bit for bit what sb_project_host computes from the
    advecting velocity, followed by the back-transform."""
    c = CASES["line_stretch"]
    op = make_op(ctx, c)
    cart0 = rand_velocity(c, 8)
    solver = sb.LevelHybridSolver(op, sb.default_options())
    vel = [op.field(centering=d, data=cart0[d]) for d in range(3)]
    grad = [op.field(centering=d) for d in range(3)]
    div, phi = op.field(), op.field()
    op.sendToAdvectingVelocity(vel, 1)
    adv0 = [v.download() for v in vel]
    op.levelDivergence(div, vel)
    solver.solve(phi, div)
    op.levelGradient(grad, phi)
    op.fluxIncr(vel, grad, 1.0)
    adv1 = [v.download() for v in vel]
    op.sendToCartesianVelocity(vel, 1)
    velh, phih, n0, n1, st = solver.project_host(adv0)
    for d in range(3):
        assert np.array_equal(adv1[d], velh[d])
    assert np.array_equal(phi.download(), phih)
    chk = [op.field(centering=d, data=velh[d]) for d in range(3)]
    op.sendToCartesianVelocity(chk, 1)
    for d in range(3):
        assert np.array_equal(vel[d].download(), chk[d].download())
    solver.free()
    op.free()


@pytest.mark.parametrize("method", [2, 3])
@pytest.mark.parametrize("name", ["gsrb_stretch", "gsrb_perxy"])
def test_relax_jacobi(ctx, name, method):
    """ProjectorParameters::RelaxMethod JACOBI (2) and JACOBIRB (3): PoissonOp::jacobi_relax /
    jacobiRB_relax (PoissonOp.cpp:1709-1775, PoissonOpF.ChF:264-310)."""
    c = dict(CASES[name], relax=method)
    op = make_op(ctx, c)
    phi0, rhs0 = rand_field(c, 2), rand_field(c, 3, zero_mean=True)
    ref = run_ref("relax", inp=[phi0, rhs0], extra={"drv.relaxIters": 3}, **ref_kwargs(c))
    phi, rhs = op.field(data=phi0), op.field(data=rhs0)
    op.relax(phi, rhs, 3)
    assert rel_err(phi.download(), ref["phi"]) <= 1e-12
    op.free()


@pytest.mark.parametrize("order", [0, 2, 3])
@pytest.mark.parametrize("name", ["line_stretch", "gsrb_perxy"])
def test_solve_vcycle_prolong_orders(ctx, name, order):
    """MGSolver V-cycles with proj.prolongOrder 0 (injection), 2 and 3 (quadratic upgrades, PoissonOp.cpp:1032-1152);
    the default test sets use 1 in V-cycles and 3 in FMG."""
    c = CASES[name]
    op = make_op(ctx, c)
    rhs0 = rand_field(c, 4, zero_mean=True)
    over = dict(V_OPTS, prolongOrder=order)
    ref = run_ref("solve", inp=[rhs0], extra=_proj_overrides(over), **ref_kwargs(c))
    solver = sb.LevelHybridSolver(op, sb.default_options(**over))
    phi, rhs = op.field(), op.field(data=rhs0)
    st = solver.solve(phi, rhs)
    assert st.status == int(ref.kv["status"])
    assert st.max_depth == int(ref.kv["maxDepth"])
    ref_norms = ref["norms"][1:]
    if st.status == 0:
        # DIVERGED: the solver withdraws the last correction and drops its norm from the history
        # (MGSolverI.H:383-392); the oracle's tap still saw that norm() call
        ref_norms = ref_norms[:-1]
    assert_norms(st.norms, ref_norms)
    assert rel_err(phi.download(), ref["phi"]) <= 1e-9
    solver.free()
    op.free()
