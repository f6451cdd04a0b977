"""GPU parity tests of the AMR half of the projection: coarse-fine ghost interpolation,
the refluxed composite operator, AMRNormLevel and the composite solve (AMRHybridSolver), called through the C ABI and
compared with the oracle -- the reference's own AMRHybridSolver / PoissonOp / MappedQuadCFInterp /
AnisotropicFluxRegister run live on the same seeded inputs -- and with the committed fixtures tests/golden/amr/*.npz.

Tolerances (BASELINE.json north_star): composite residual norms per iteration within 1e-10 of the initial norm,
pressure on every level within 1e-9 relative in max-norm, identical solver status and iteration count.  The
element-wise pieces are held far tighter (stated per test)."""
import glob
import os

import numpy as np
import pytest

import somar_b200 as sb
from _oracle import have_ref, run_ref
from amr_cases import (AMR_CASES, C3_DECK, SPEC_CASES, composite_rhs_levels, level_shapes, level_specs, make_amr_ops, ndim, num_levels,
                       rand_velocity_levels, ref_kwargs_amr3)
from cases import rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (have_ref(3) and have_ref(2)), reason="oracle/_ref not built")]
HERE = os.path.dirname(os.path.abspath(__file__))
ALL = dict(AMR_CASES)
INTERP = dict(SPEC_CASES, **{n: c for n, c in AMR_CASES.items() if n not in SPEC_CASES})


def up(a, D):
    """[nx, nz] -> [nx, 1, nz] for the 2-D build (Fortran order keeps the memory image)."""
    a = np.asfortranarray(a)
    return a if D == 3 else a.reshape((a.shape[0], 1, a.shape[1]), order="F")


def down(a, D):
    return a if D == 3 else a.reshape((a.shape[0], a.shape[2]), order="F")


def rand_levels(c, seed):
    rng = np.random.default_rng(seed)
    return [np.asfortranarray(rng.standard_normal(sh)) for sh in level_shapes(c)]


def uncovered_masks(c):
    D = ndim(c)
    specs, shapes = level_specs(c), level_shapes(c)
    pick = (lambda v: np.array([v[0], v[2]])) if D == 2 else (lambda v: np.array(v))
    masks = [np.ones(sh, bool) for sh in shapes]
    for l in range(1, len(specs)):
        ref = pick(specs[l]["ref"])
        lo_c = pick(specs[l]["reg_lo"]) // ref - pick(specs[l - 1]["reg_lo"])
        n_c = np.array(shapes[l]) // ref
        masks[l - 1][tuple(slice(lo_c[d], lo_c[d] + n_c[d]) for d in range(D))] = False
    return masks


@pytest.mark.parametrize("name", sorted(INTERP))
def test_cf_ghost_interpolation(ctx, name):
    """PoissonOp::applyBCs(phi, &crsePhi, t, true, false) -> CFInterp::interpAtCFI -> MappedQuadCFInterp: every
    coarse-fine face ghost of every refined level (centred, one-sided and order-dropped coarse stencils included)."""
    c = INTERP[name]
    D, nl = ndim(c), num_levels(c)
    data = rand_levels(c, 5)
    r = run_ref("amr", inp=data, **ref_kwargs_amr3(c, **{"drv.cfInterpOnly": 1}))
    ops = make_amr_ops(ctx, c)
    fields = [ops[l].field(data=up(data[l], D)) for l in range(nl)]
    for l in range(1, nl):
        ops[l].applyBCsAMR(fields[l], crse_phi=fields[l - 1], homog_phys=True, homog_cfi=False)
        lo, hi = fields[l].box(ghost=1)
        if D == 2:
            lo[1], hi[1] = 0, 0
        got = down(fields[l].download(lo, hi), D)
        shape = tuple(n + 2 for n in level_shapes(c)[l])
        want = r["fineWithGhosts" if l == nl - 1 else "midWithGhosts"].reshape(shape, order="F")
        # faces only (coarse-fine ghosts, and the physical-boundary ghosts where the patch touches a wall): the oracle
        # dump leaves edge / corner ghosts at 0
        inner = tuple(slice(1, -1) for _ in range(D))
        assert np.array_equal(got[inner], want[inner])
        scale = np.max(np.abs(want))
        for d in range(D):
            for side in (0, -1):
                sl = list(inner)
                sl[d] = side
                assert np.max(np.abs(got[tuple(sl)] - want[tuple(sl)])) <= 1e-13 * scale, (l, d, side)
    for o in ops:
        o.free()


@pytest.mark.parametrize("name", sorted(ALL))
def test_composite_operator_and_norms(ctx, name):
    """rhs - L[phi] with rhs = 0 through AMRResidualNF / AMRResidual / AMRResidualNC (quadratic coarse-fine ghosts,
    flux-register reflux) and AMRNormLevel (covered cells masked), on random data."""
    c = ALL[name]
    D, nl = ndim(c), num_levels(c)
    data = rand_levels(c, 6)
    r = run_ref("amr", inp=data, **ref_kwargs_amr3(c, **{"drv.applyOnly": 1}))
    ops = make_amr_ops(ctx, c)
    phi = [ops[l].field(data=up(data[l], D)) for l in range(nl)]
    res = [ops[l].field() for l in range(nl)]
    zero = [ops[l].field() for l in range(nl)]
    for l in range(nl - 1, -1, -1):   # the order of the oracle hook (a finer level's ghosts are reset by the coarser one's reflux)
        ops[l].AMRResidual(res[l], phi[l], zero[l], phi_fine=phi[l + 1] if l + 1 < nl else None, finer_op=ops[l + 1] if l + 1 < nl else None,
                           phi_crse=phi[l - 1] if l > 0 else None)
    masks = uncovered_masks(c)
    for l in range(nl):
        want = r[f"minusL{l}"].reshape(level_shapes(c)[l], order="F")
        got = down(res[l].download(), D)
        m = masks[l]
        assert np.max(np.abs(got - want)[m]) <= 1e-13 * np.max(np.abs(want)), l
        n = ops[l].AMRNormLevel(res[l], finer_op=ops[l + 1] if l + 1 < nl else None, p=2)
        assert abs(n - r.kv[f"minusLNorm{l}"]) <= 1e-12 * r.kv[f"minusLNorm{l}"], l
    for o in ops:
        o.free()


def _solve(ctx, c, rhs):
    D, nl = ndim(c), num_levels(c)
    ops = make_amr_ops(ctx, c)
    solver = sb.AMRHybridSolver(ops, 0, nl - 1, sb.default_options())
    phi = [ops[l].field() for l in range(nl)]
    rr = [ops[l].field(data=up(rhs[l], D)) for l in range(nl)]
    st = solver.solve(phi, rr)
    out = [down(phi[l].download(), D) for l in range(nl)]
    solver.free()
    for o in ops:
        o.free()
    return st, out


def _check_against(st, phis, r, nl):
    assert st.status == int(r.kv["status"])
    lev = r["amrLevelNorms"].reshape(-1, nl)
    comp = np.sqrt((lev ** 2).sum(axis=1))       # AMRHybridSolver::computeAMRResidual, normType 2 (:545-554)
    assert st.num_norms == len(comp), (st.norms, comp)
    got = np.array(st.norms)
    assert np.all(np.abs(got - comp) <= 1e-10 * comp[0]), (got, comp)
    assert np.all(np.abs(got - comp) <= 1e-6 * comp + 1e-11 * comp[0]), (got, comp)
    for l in range(nl):
        assert rel_err(phis[l], r[f"phi{l}"]) <= 1e-9, l


@pytest.mark.parametrize("name", sorted(ALL))
def test_amr_solve_matches_the_reference(ctx, name):
    """AMRHybridSolver::solve on a solvable composite right-hand side: status, number of iterations, composite residual
    norm per iteration, pressure on every level."""
    c = ALL[name]
    rhs, _ = composite_rhs_levels(c, 3)
    r = run_ref("amr", inp=rhs, **ref_kwargs_amr3(c))
    st, phis = _solve(ctx, c, rhs)
    _check_against(st, phis, r, num_levels(c))


FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "amr", "*.npz")))


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_amr_solve_matches_golden(ctx, path):
    """The committed outputs of the reference's AMRHybridSolver (tests/golden/make_golden_amr.py)."""
    z = np.load(path)
    c = AMR_CASES[str(z["name"])]
    st, phis = _solve(ctx, c, [z["rhs0"], z["rhs1"]])
    assert st.status == int(z["status"])
    assert rel_err(phis[0], z["phi0"]) <= 1e-9
    assert rel_err(phis[1], z["phi1"]) <= 1e-9


def test_c3_buoyant_vortex_ring_hierarchy_at_deck_size(ctx):
    """BASELINE.json configs[2]: exec/BuoyantVortexRing/inputs -- 64^3 base level (triply periodic, boxes of 16^3, GSRB) and
    two refined levels at ratio (4, 4, 4); rectangular patches stand in for the deck's vorticity-tagged grids."""
    c = C3_DECK
    rhs, _ = composite_rhs_levels(c, 3)
    r = run_ref("amr", inp=rhs, timeout=3000, **ref_kwargs_amr3(c))
    st, phis = _solve(ctx, c, rhs)
    _check_against(st, phis, r, 3)


def composite_project(ops, c, vel0):
    """AMRNSLevel::projectDownToThis (AMRNSLevelProject.cpp:740-860) on device-resident fields, through the C ABI."""
    D, nl = ndim(c), num_levels(c)
    upv = lambda lev: [up(lev[0], D), None, up(lev[1], D)] if D == 2 else [up(a, D) for a in lev]
    vel = [ops[l].flux(upv(vel0[l])) for l in range(nl)]
    grad = [ops[l].flux() for l in range(nl)]
    rhs = [ops[l].field() for l in range(nl)]
    phi = [ops[l].field() for l in range(nl)]
    out = {}

    def comp_div(tag):
        for l in range(nl):
            ops[l].compDivergence(rhs[l], vel[l], vel[l + 1] if l + 1 < nl else None, ops[l + 1] if l + 1 < nl else None)
        for l in range(nl):
            out[f"{tag}{l}"] = down(rhs[l].download(), D)
    comp_div("div_init")
    for l in range(nl - 1, 0, -1):
        ops[l].averageDown(rhs[l - 1], rhs[l])
    for l in range(nl):
        out[f"rhs{l}"] = down(rhs[l].download(), D)
    solver = sb.AMRHybridSolver(ops, 0, nl - 1, sb.default_options())
    st = solver.solve(phi, rhs)
    for l in range(nl):
        ops[l].compGradient(grad[l], phi[l], crse_phi=phi[l - 1] if l else None, homog_phys=True, homog_cfi=False)
        ops[l].fluxIncr(vel[l], grad[l], 1.0)
    for l in range(nl):
        out[f"phi{l}"] = down(phi[l].download(), D)
        k = 0
        for d in range(3):
            if vel[l][d] is not None:
                out[f"vel{l}_{k}"] = down(vel[l][d].download(), D)
                k += 1
    comp_div("div_final")
    solver.free()
    return st, out


@pytest.mark.parametrize("name", ["amr_r2_centre", "amr_aniso_wall", "amr_r4_periodic", "amr2d_r2_centre", "amr2d_djl_r22", "amr3_c3_mini"])
def test_composite_projection(ctx, name):
    """The sync projection over the AMR hierarchy: compDivergence with flux-register refluxing, averageDown of the
    right-hand side, AMRHybridSolver, compGradient with interpolated coarse-fine ghosts, vel -= grad."""
    c = ALL[name]
    D, nl = ndim(c), num_levels(c)
    vel0 = rand_velocity_levels(c, 7)
    r = run_ref("amr", inp=[a for lev in vel0 for a in lev], **ref_kwargs_amr3(c, **{"drv.compProject": 1}))
    ops = make_amr_ops(ctx, c)
    st, out = composite_project(ops, c, vel0)
    masks = uncovered_masks(c)
    shapes = level_shapes(c)
    assert st.status == int(r.kv["status"])
    lev = r["amrLevelNorms"].reshape(-1, nl)
    comp = np.sqrt((lev ** 2).sum(axis=1))
    assert st.num_norms == len(comp)
    assert np.all(np.abs(np.array(st.norms) - comp) <= 1e-10 * comp[0])
    for l in range(nl):
        m = masks[l]
        for tag in ("div_init", "div_final"):   # covered cells hold the reference's unmasked register leftovers: not compared
            want = r[f"{tag}{l}"].reshape(shapes[l], order="F")
            scale = np.max(np.abs(r[f"div_init{l}"]))
            assert np.max(np.abs(out[f"{tag}{l}"] - want)[m]) <= (1e-13 if tag == "div_init" else 1e-9) * scale, (tag, l)
        assert rel_err(out[f"rhs{l}"], r[f"rhs{l}"]) <= 1e-13, l
        assert rel_err(out[f"phi{l}"], r[f"phi{l}"]) <= 1e-9, l
        for d in range(D):
            assert rel_err(out[f"vel{l}_{d}"], r[f"vel{l}_{d}"]) <= 1e-9, (l, d)
    for o in ops:
        o.free()
