"""GPU parity tests for the 2-D build of the reference (CH_SPACEDIM == 2: BASELINE.json configs
C1 LockExchange and C2 DJL are 2-D).  The oracle is oracle/_ref/d2/somar_ref -- the reference's
own C++ compiled with -DCH_SPACEDIM=2.  The library runs 2-D problems as dim = 2 with the
directions (x, z) in slots 0 and 2 (ny = 1), so every 2-D array [nx, nz] is viewed as [nx, 1, nz].

`c1_lockexchange` is BASELINE.json configs[0] at its full size: exec/LockExchange/
inputs.ThesisTestCase2D.research1 (1152 x 128, L = 36 x 2, Cartesian, GSRB, no AMR); the deck
lacks base.maxBaseGridSize / base.blockFactor (SURVEY.md 8, C1 caveats), chosen here as 128 / 16.
Tolerances as in test_parity_gpu.py (north_star: norms 1e-10 relative, fields 1e-9 max-norm)."""
import numpy as np
import pytest

import somar_b200 as sb
from _oracle import have_ref, run_ref
from cases import rel_err
from test_parity_gpu import V_OPTS, _proj_overrides, assert_norms

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_ref(2), reason="oracle/_ref/d2/somar_ref not built")]

# lepticity = max(dXi_h) / Lz stays below the MG-mode limit for all of these (LevelHybridSolver.cpp:457-498)
CASES2D = {
    "c1_lockexchange": dict(nx=(1152, 128), L=(36.0, 2.0), max_box=(128, 0), bf=16, periodic=(0, 0), relax=5, map="cartesian", ampl=(0, 0)),
    "gsrb2d_stretch": dict(nx=(64, 32), L=(2.0, 4.0), max_box=(16, 0), bf=4, periodic=(0, 0), relax=5, map="stretched", ampl=(0.04, 0.2)),
    "line2d_cart": dict(nx=(64, 32), L=(4.0, 1.0), max_box=(16, 0), bf=4, periodic=(0, 0), relax=6, map="cartesian", ampl=(0, 0)),
    "line2d_zstretch_perx": dict(nx=(128, 32), L=(8.0, 1.0), max_box=(32, 0), bf=8, periodic=(1, 0), relax=6, map="stretched", ampl=(0.0, -0.1)),
    "line2d_stretch": dict(nx=(64, 16), L=(2.0, 1.0), max_box=(32, 0), bf=4, periodic=(0, 0), relax=6, map="stretched", ampl=(0.05, -0.1)),
}
ALL = sorted(CASES2D)


def geometry(c):
    nx = np.array([c["nx"][0], 1, c["nx"][1]])
    L = np.array([c["L"][0], 1.0, c["L"][1]])
    dXi = L / nx
    lo = np.array([0, 0, -nx[2]])
    hi = lo + nx - 1
    return nx, L, dXi, lo, hi


def make_op(ctx, c):
    nx, L, dXi, lo, hi = geometry(c)
    blo, bhi = sb.make_base_grids(lo, hi, (c["max_box"][0], 0, 0), (1, 0, 0), c["bf"])
    xmin = lo * dXi
    kind = sb.MAP_CARTESIAN if c["map"] == "cartesian" else sb.MAP_STRETCHED
    return sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, periodic=(c["periodic"][0], 0, c["periodic"][1]), dim=2, map_kind=kind,
                        map_xmin=xmin, map_xmax=xmin + L, map_ampl=(c["ampl"][0], 0.0, c["ampl"][1]), relax_method=c["relax"])


def ref_kwargs(c):
    return dict(nx=c["nx"], L=c["L"], max_box=c["max_box"], block_factor=c["bf"], offset=(0, -c["nx"][1]), periodic=c["periodic"],
                relax=c["relax"], mapname=c["map"], ampl=c["ampl"], dim=2)


def rand_field(c, seed, zero_mean=False):
    a = np.random.default_rng(seed).standard_normal(c["nx"])
    if zero_mean:
        a -= a.mean()
    return np.asfortranarray(a)


def rand_velocity(c, seed):
    rng = np.random.default_rng(seed)
    out = []
    for d in range(2):
        shape = list(c["nx"])
        shape[d] += 1
        u = rng.standard_normal(tuple(shape))
        lo, hi = [slice(None)] * 2, [slice(None)] * 2
        lo[d], hi[d] = 0, -1
        if c["periodic"][d]:
            u[tuple(hi)] = u[tuple(lo)]
        else:
            u[tuple(lo)] = 0.0
            u[tuple(hi)] = 0.0
        out.append(np.asfortranarray(u))
    return out


def up(a):
    """[nx, nz] -> [nx, 1, nz] (Fortran order keeps the memory image)."""
    a = np.asfortranarray(a)
    return a.reshape((a.shape[0], 1, a.shape[1]), order="F")


@pytest.mark.parametrize("name", ALL)
def test_coefficients_2d(ctx, name):
    c = CASES2D[name]
    op = make_op(ctx, c)
    ref = run_ref("applyop", inp=[rand_field(c, 1)], **ref_kwargs(c))
    assert op.has_null_space == bool(ref.kv["hasNullSpace"])
    assert rel_err(op.coefficient(0), ref["J"]) <= 4e-16
    assert rel_err(op.coefficient(1), ref["Dinv"]) <= 1e-15
    for d2, d3 in ((0, 0), (1, 2)):
        assert rel_err(op.coefficient(2 + d3), ref[f"M{d2}"]) == 0.0
        assert rel_err(op.coefficient(5 + d3), ref[f"Jgup{d2}"]) <= 4e-16
    op.free()


@pytest.mark.parametrize("name", ALL)
def test_apply_op_2d(ctx, name):
    c = CASES2D[name]
    op = make_op(ctx, c)
    phi0 = rand_field(c, 1)
    ref = run_ref("applyop", inp=[phi0], **ref_kwargs(c))
    phi, lhs = op.field(data=up(phi0)), op.field()
    op.applyOp(lhs, phi)
    assert rel_err(lhs.download(), ref["lhs"]) <= 1e-14
    assert abs(op.norm(lhs, 2) - ref.kv["norm2"]) <= 1e-13 * ref.kv["norm2"]
    assert abs(op.norm(lhs, 0) - ref.kv["norm0"]) <= 1e-14 * ref.kv["norm0"]
    op.free()


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("iters", [1, 3])
def test_relax_2d(ctx, name, iters):
    c = CASES2D[name]
    op = make_op(ctx, c)
    phi0, rhs0 = rand_field(c, 2), rand_field(c, 3, zero_mean=True)
    ref = run_ref("relax", inp=[phi0, rhs0], extra={"drv.relaxIters": iters}, **ref_kwargs(c))
    phi, rhs = op.field(data=up(phi0)), op.field(data=up(rhs0))
    op.relax(phi, rhs, iters)
    assert rel_err(phi.download(), ref["phi"]) <= 1e-12
    op.free()


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("optset", ["defaults", "vcycle"])
def test_solve_2d(ctx, name, optset):
    c = CASES2D[name]
    op = make_op(ctx, c)
    rhs0 = rand_field(c, 4, zero_mean=True)
    over = {} if optset == "defaults" else V_OPTS
    ref = run_ref("solve", inp=[rhs0], extra=_proj_overrides(over), **ref_kwargs(c))
    assert int(ref.kv["solveMode"]) == 1
    solver = sb.LevelHybridSolver(op, sb.default_options(**over))
    phi, rhs = op.field(), op.field(data=up(rhs0))
    st = solver.solve(phi, rhs)
    assert st.max_depth == int(ref.kv["maxDepth"])
    assert st.status == int(ref.kv["status"])
    ref_norms = ref["norms"][1:]
    assert st.num_norms == len(ref_norms)
    assert_norms(st.norms, ref_norms)
    assert rel_err(phi.download(), ref["phi"]) <= 1e-9
    solver.free()
    op.free()


@pytest.mark.parametrize("name", ALL)
def test_project_2d(ctx, name):
    c = CASES2D[name]
    op = make_op(ctx, c)
    vel0 = rand_velocity(c, 5)
    ref = run_ref("project", inp=vel0, **ref_kwargs(c))
    solver = sb.LevelHybridSolver(op, sb.default_options())
    vel, phi, n0, n1, st = solver.project_host([up(vel0[0]), None, up(vel0[1])])
    assert abs(n0 - ref.kv["initDivNorm"]) <= 1e-13 * ref.kv["initDivNorm"]
    assert st.status == int(ref.kv["status"])
    assert_norms(st.norms, ref["norms"][1:])
    assert rel_err(phi, ref["phi"]) <= 1e-9
    assert rel_err(vel[0], ref["vel0"]) <= 1e-9
    assert rel_err(vel[2], ref["vel1"]) <= 1e-9
    assert abs(n1 - ref.kv["finalDivNorm"]) <= 1e-6 * max(ref.kv["finalDivNorm"], 1e-30) + 1e-12 * n0
    solver.free()
    op.free()
