"""CPU tests of the two-level AMR oracle (AMRHybridSolver, quadratic
coarse-fine ghost interpolation, flux-register reflux).  TEST INFRASTRUCTURE for the next round's
CUDA path: the reference's own C++ (AMRHybridSolver, MappedQuadCFInterp, AnisotropicFluxRegister,
PoissonOp's AMR interface) linked against oracle/fort_leaves.cpp.  The reference ships no vectors for
this path either, so the restated coarse-fine leaves are pinned by known answers and invariants:

* the quadratic CF interpolation reproduces any quadratic polynomial at the fine ghost cells to rounding
  (pins MAPPEDPHISTAR / MAPPEDQUADINTERP, MappedQuadCFInterpF.ChF:9-127, with mixed terms and refinement
  ratios 2, 4 and (2, 2, 1));
* the composite operator with reflux is conservative: the composite integral of rhs - L[phi] is zero to
  rounding for ANY phi (pins ANISOTROPICINCREMENTFINE and the flux bookkeeping);
* AMRHybridSolver converges on a solvable composite right-hand side, and the composite residual evaluated
  afterwards through the operators' public AMR interface confirms the drop;
* committed fixtures (tests/golden/amr/*.npz, made by tests/golden/make_golden_amr.py from this oracle)
  are reproduced bit for bit."""
import glob
import os

import numpy as np
import pytest

from _oracle import have_ref, run_ref
from amr_cases import AMR_CASES, composite_integral, composite_rhs, fine_shape, ref_kwargs_amr

pytestmark = pytest.mark.skipif(not (have_ref(3) and have_ref(2)), reason="oracle/_ref/d{2,3}/somar_ref not built")
CASES3D = sorted(n for n, c in AMR_CASES.items() if len(c["nx"]) == 3 and "region2" not in c)
TWO_LEVEL = sorted(n for n, c in AMR_CASES.items() if "region2" not in c)
HERE = os.path.dirname(os.path.abspath(__file__))


def _poly(x, y, z):
    return 1.0 + 0.3 * x - 0.7 * y + 1.1 * z + 0.5 * x * x - 0.25 * y * y + 0.8 * z * z + 0.6 * x * y - 0.9 * y * z + 0.4 * x * z


def _centres(lo, n, dx):
    return [(np.arange(lo[d], lo[d] + n[d]) + 0.5) * dx[d] for d in range(3)]


@pytest.mark.parametrize("name", CASES3D)
def test_cf_interpolation_reproduces_quadratics(name):
    c = AMR_CASES[name]
    nx, ref, reg = c["nx"], c["ref"], c["region"]
    dxc = np.array(c["L"]) / np.array(nx)
    dxf = dxc / np.array(ref)
    xc = _centres(c["offset"], nx, dxc)
    p0 = _poly(xc[0][:, None, None], xc[1][None, :, None], xc[2][None, None, :])
    flo = tuple(reg[d] * ref[d] for d in range(3))
    nf = fine_shape(c)
    xf = _centres(flo, nf, dxf)
    p1 = _poly(xf[0][:, None, None], xf[1][None, :, None], xf[2][None, None, :])
    r = run_ref("amr", inp=[np.asfortranarray(p0), np.asfortranarray(p1)], **ref_kwargs_amr(c, **{"drv.cfInterpOnly": 1}))
    g = r["fineWithGhosts"].reshape(tuple(n + 2 for n in nf), order="F")
    xg = _centres(tuple(v - 1 for v in flo), tuple(n + 2 for n in nf), dxf)
    want = _poly(xg[0][:, None, None], xg[1][None, :, None], xg[2][None, None, :])
    scale = np.max(np.abs(want))
    assert np.array_equal(g[1:-1, 1:-1, 1:-1], p1)
    checked = 0
    for d in range(3):
        for side, at_wall in ((0, reg[d] == c["offset"][d]), (-1, reg[3 + d] == c["offset"][d] + nx[d] - 1)):
            if at_wall and not c["periodic"][d]:
                continue                      # physical boundary, not a coarse-fine interface
            sl = [slice(1, -1)] * 3
            sl[d] = side
            # where the patch touches a wall in a tangential direction, the coarse slopes of the first coarse
            # cell use physical ghost values (not part of this known answer): leave that strip out
            for t in range(3):
                if t == d or c["periodic"][t]:
                    continue
                lo_cut = 1 + (ref[t] if reg[t] == c["offset"][t] else 0)
                hi_cut = -1 - (ref[t] if reg[3 + t] == c["offset"][t] + nx[t] - 1 else 0)
                sl[t] = slice(lo_cut, hi_cut)
            assert np.max(np.abs(g[tuple(sl)] - want[tuple(sl)])) <= 1e-13 * scale
            checked += 1
    assert checked >= 5


@pytest.mark.parametrize("name", TWO_LEVEL)
def test_two_level_solve_converges_and_is_conservative(name):
    c = AMR_CASES[name]
    r0, r1 = composite_rhs(c, 3)
    assert abs(composite_integral(c, r0.ravel(order="F"), r1)) <= 1e-10 * np.abs(r0).sum()
    r = run_ref("amr", inp=[r0, r1], **ref_kwargs_amr(c))
    assert int(r.kv["status"]) == 1                                   # SolverStatus::CONVERGED
    # the residual evaluated independently of the solver's bookkeeping dropped by the solver's relTol
    assert r.kv["res_finalNorm0"] <= 2e-6 * r.kv["res_initNorm0"]
    assert abs(r.kv["res_initNorm0"] - r.kv["initResNorm"]) <= 0.2 * r.kv["initResNorm"]
    # conservation of the refluxed composite operator: integral of (rhs - L[phi]) = integral of rhs = 0
    nf = fine_shape(c)
    assert r.kv["res_finalNorm1"] <= 2e-6 * r.kv["res_initNorm0"]
    tot = composite_integral(c, r["res_final0"], r["res_final1"].reshape(nf, order="F"))
    assert abs(tot) <= 1e-9 * (np.abs(r["res_init0"]).sum() + np.abs(r["res_init1"]).sum())


FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "amr", "*.npz")))


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_amr_oracle_reproduces_golden(path):
    z = np.load(path)
    c = AMR_CASES[str(z["name"])]
    r = run_ref("amr", inp=[z["rhs0"], z["rhs1"]], **ref_kwargs_amr(c))
    assert int(r.kv["status"]) == int(z["status"])
    assert np.array_equal(r["phi0"], z["phi0"].ravel(order="F"))
    assert np.array_equal(r["phi1"], z["phi1"].ravel(order="F"))


from amr_cases import SPEC_CASES  # noqa: E402


@pytest.mark.parametrize("name", sorted(SPEC_CASES))
def test_cf_interpolation_numpy_spec_matches_the_reference(name):
    """tests/amr_cfinterp_spec.py restates the reference's coarse-fine ghost interpolation (which coarse cells
    get centred, one-sided or dropped derivative stencils, phistar, the normal quadratic) in numpy -- the
    blueprint of next round's CUDA kernel.  On random data it must give the reference's ghost values bit for bit."""
    from amr_cfinterp_spec import CFInterpSpec
    c = SPEC_CASES[name]
    nx, ref, reg, off = c["nx"], c["ref"], c["region"], np.array(c["offset"])
    rng = np.random.default_rng(5)
    nf = fine_shape(c)
    p0, p1 = rng.standard_normal(nx), rng.standard_normal(nf)
    r = run_ref("amr", inp=[np.asfortranarray(p0), np.asfortranarray(p1)], **ref_kwargs_amr(c, **{"drv.cfInterpOnly": 1}))
    g = r["fineWithGhosts"].reshape(tuple(n + 2 for n in nf), order="F")
    flo = np.array([reg[d] * ref[d] for d in range(3)])
    fmb = c["fine_max_box"]                       # fine boxes as oracle/ref_driver.cpp cuts them
    nb = [(nf[d] + fmb - 1) // fmb for d in range(3)]
    sz = np.array([nf[d] // nb[d] for d in range(3)])
    boxes = [(flo + np.array(i) * sz, flo + np.array(i) * sz + sz - 1) for i in np.ndindex(*nb)]
    dxf = np.array(c["L"]) / np.array(nx) / np.array(ref)
    spec = CFInterpSpec(off, off + np.array(nx) - 1, c["periodic"], ref, boxes, dxf)
    out = spec.ghosts(lambda cc: p0[tuple(np.array(cc) - off)], lambda ff: p1[tuple(np.array(ff) - flo)])
    assert len(out) > 0
    bad = [(f, g[tuple(np.array(f) - flo + 1)], v) for f, v in out.items() if g[tuple(np.array(f) - flo + 1)] != v]
    assert not bad, (len(bad), bad[:3])


@pytest.mark.parametrize("name", CASES3D)
def test_composite_operator_numpy_spec_matches_the_reference(name):
    """tests/amr_operator_spec.py restates the two-level composite operator (quadratic coarse-fine ghosts on the
    fine level, flux-register reflux on the coarse level) in numpy.  On random data it must reproduce the
    reference on every fine cell and every uncovered coarse cell.  (Covered coarse cells are not compared: next to
    fine-fine box interfaces the reference also adds fine-register increments there; AMRNormLevel masks them and
    the V-cycle overwrites them with the restricted fine residual.)"""
    from amr_cases import region_slices
    from amr_operator_spec import composite_minus_L
    c = AMR_CASES[name]
    nx, ref, reg = c["nx"], c["ref"], c["region"]
    rng = np.random.default_rng(6)
    nf = fine_shape(c)
    p0, p1 = rng.standard_normal(nx), rng.standard_normal(nf)
    r = run_ref("amr", inp=[np.asfortranarray(p0), np.asfortranarray(p1)], **ref_kwargs_amr(c, **{"drv.applyOnly": 1}))
    flo = np.array([reg[d] * ref[d] for d in range(3)])
    fmb = c["fine_max_box"]
    nb = [(nf[d] + fmb - 1) // fmb for d in range(3)]
    sz = np.array([nf[d] // nb[d] for d in range(3)])
    boxes = [(flo + np.array(i) * sz, flo + np.array(i) * sz + sz - 1) for i in np.ndindex(*nb)]
    m0, m1 = composite_minus_L(c, p0, p1, boxes)
    R0, R1 = r["minusL0"].reshape(nx, order="F"), r["minusL1"].reshape(nf, order="F")
    uncovered = np.ones(nx, bool)
    uncovered[region_slices(c)] = False
    assert np.max(np.abs(m1 - R1)) <= 1e-13 * np.max(np.abs(R1))
    assert np.max(np.abs(m0 - R0)[uncovered]) <= 1e-13 * np.max(np.abs(R0))


THREE_LEVEL = sorted(n for n, c in AMR_CASES.items() if "region2" in c)


@pytest.mark.parametrize("name", THREE_LEVEL)
def test_three_level_solve_converges(name):
    """Base level + two nested refined patches (the BuoyantVortexRing deck's hierarchy in miniature): the middle level
    goes through AMRResidual with both neighbours (AMRHybridSolver.cpp:603-626)."""
    from amr_cases import composite_rhs_levels, ref_kwargs_amr3
    c = AMR_CASES[name]
    rhs, _ = composite_rhs_levels(c, 3)
    r = run_ref("amr", inp=rhs, **ref_kwargs_amr3(c))
    assert int(r.kv["numLevels"]) == 3 and int(r.kv["status"]) == 1
    norms = r["amrLevelNorms"].reshape(-1, 3)
    comp = np.sqrt((norms ** 2).sum(axis=1))
    assert comp[-1] <= 1e-6 * comp[0]
    for l in range(3):
        assert r.kv[f"res_finalNorm{l}"] <= 2e-6 * comp[0]
