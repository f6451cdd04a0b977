"""The drop-in proof (SURVEY.md 8 row b): the reference's own driver code with the two `using` lines of
Grade5_SOMAR/AMRNSLevel.H:1448-1450 swapped for the B200 shim classes of integration/ (B200PoissonOp : PoissonOp,
B200LevelHybridSolver), compiled against the reference's headers by oracle/build_ref.sh and linked to
libsomar_b200.so -- oracle/_ref/d*/somar_ref_b200.  Every call goes reference C++ API -> C ABI -> CUDA; the result is
compared with the unmodified reference binary on the same inputs:

* the operator through its virtual LevelOperator / MGOperator interface, driven by the REFERENCE'S OWN
  MGSolver<LevelData<FArrayBox>> and BiCGStabSolver (drv.useMGSolver=1): applyOp, relax, preCond, MGRestrict, MGProlong,
  norm, dotProduct on the device, the solver's control flow on the host;
* B200LevelHybridSolver behind LevelHybridSolver's define / solve surface (MG and leptic modes);
* the projection bracket of AMRNSLevel::projectCorrect: levelDivergence, norm, solve, levelGradient."""
import numpy as np
import pytest

import test_parity2d_gpu as t2
from _oracle import have_ref, run_ref
from cases import CASES, rand_field, rand_velocity, ref_kwargs, rel_err
from test_leptic_gpu import LEPTIC3D
from test_parity_gpu import V_OPTS, _proj_overrides, assert_norms

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (have_ref(3) and have_ref(3, shim=True)), reason="oracle/_ref/d3/somar_ref_b200 not built")]


@pytest.mark.parametrize("name", ["line_stretch", "gsrb_perxy", "line_perx"])
def test_operator_calls_through_the_reference_api(name):
    c = CASES[name]
    phi0, rhs0 = rand_field(c, 2), rand_field(c, 3, zero_mean=True)
    a = run_ref("applyop", inp=[phi0], **ref_kwargs(c))
    b = run_ref("applyop", inp=[phi0], shim=True, **ref_kwargs(c))
    assert b.kv["hasNullSpace"] == a.kv["hasNullSpace"]
    assert rel_err(b["lhs"], a["lhs"]) <= 1e-14
    assert abs(b.kv["norm2"] - a.kv["norm2"]) <= 1e-13 * a.kv["norm2"]
    a = run_ref("relax", inp=[phi0, rhs0], extra={"drv.relaxIters": 3}, **ref_kwargs(c))
    b = run_ref("relax", inp=[phi0, rhs0], extra={"drv.relaxIters": 3}, shim=True, **ref_kwargs(c))
    assert rel_err(b["phi"], a["phi"]) <= 1e-12


@pytest.mark.parametrize("name,optset", [("line_cart", "defaults"), ("gsrb_stretch", "vcycle"), ("line_aniso", "vcycle")])
def test_reference_mgsolver_drives_the_device_operator(name, optset):
    """MGSolver<LevelData<FArrayBox>>::solve of the reference, unmodified, with a B200PoissonOp as its MGOperator."""
    c = CASES[name]
    rhs0 = rand_field(c, 4, zero_mean=True)
    extra = dict(_proj_overrides({} if optset == "defaults" else V_OPTS), **{"drv.useMGSolver": 1})
    a = run_ref("solve", inp=[rhs0], extra=extra, **ref_kwargs(c))
    b = run_ref("solve", inp=[rhs0], extra=extra, shim=True, **ref_kwargs(c))
    assert b.kv["impl"] == "b200-shim"
    assert int(b.kv["status"]) == int(a.kv["status"]) and int(b.kv["maxDepth"]) == int(a.kv["maxDepth"])
    assert len(b["norms"]) == len(a["norms"])
    assert_norms(b["norms"], a["norms"])
    assert rel_err(b["phi"], a["phi"]) <= 1e-9


@pytest.mark.parametrize("name", ["line_stretch", "gsrb_cart", "lepmg3d_cart", "lep3d_perx"])
def test_level_hybrid_solver_shim(name):
    c = CASES[name] if name in CASES else LEPTIC3D[name][0]
    rhs0 = rand_field(c, 4, zero_mean=True)
    a = run_ref("solve", inp=[rhs0], **ref_kwargs(c))
    b = run_ref("solve", inp=[rhs0], shim=True, **ref_kwargs(c))
    assert int(b.kv["status"]) == int(a.kv["status"]) and int(b.kv["maxDepth"]) == int(a.kv["maxDepth"])
    assert int(b.kv["solveMode"]) == int(a.kv["solveMode"])
    assert len(b["hybridNorms"]) == len(a["hybridNorms"])
    assert_norms(b["hybridNorms"], a["hybridNorms"])
    assert rel_err(b["phi"], a["phi"]) <= 1e-9


@pytest.mark.parametrize("name", ["line_stretch", "gsrb_stretch"])
def test_projection_bracket_through_the_shim(name):
    c = CASES[name]
    vel0 = rand_velocity(c, 5)
    a = run_ref("project", inp=vel0, **ref_kwargs(c))
    b = run_ref("project", inp=vel0, shim=True, **ref_kwargs(c))
    assert rel_err(b["div"], a["div"]) <= 1e-14
    assert abs(b.kv["initDivNorm"] - a.kv["initDivNorm"]) <= 1e-13 * a.kv["initDivNorm"]
    assert int(b.kv["status"]) == int(a.kv["status"])
    assert rel_err(b["phi"], a["phi"]) <= 1e-9
    for d in range(3):
        assert rel_err(b[f"grad{d}"], a[f"grad{d}"]) <= 1e-9
        assert rel_err(b[f"vel{d}"], a[f"vel{d}"]) <= 1e-9


@pytest.mark.skipif(not (have_ref(2) and have_ref(2, shim=True)), reason="oracle/_ref/d2/somar_ref_b200 not built")
def test_shim_2d_build_c1_lockexchange():
    """CH_SPACEDIM = 2 build of the shim on BASELINE.json configs[0] (LockExchange 2-D, full size)."""
    c = t2.CASES2D["c1_lockexchange"]
    vel0 = t2.rand_velocity(c, 5)
    a = run_ref("project", inp=vel0, **t2.ref_kwargs(c))
    b = run_ref("project", inp=vel0, shim=True, **t2.ref_kwargs(c))
    assert int(b.kv["status"]) == int(a.kv["status"])
    assert_norms(b["hybridNorms"], a["hybridNorms"])
    assert rel_err(b["phi"], a["phi"]) <= 1e-9
    for d in range(2):
        assert rel_err(b[f"vel{d}"], a[f"vel{d}"]) <= 1e-9


@pytest.mark.parametrize("name", ["amr_r2_centre"])
def test_reference_amr_solver_drives_the_device_operators(name):
    """The AMR half of the boundary through the drop-in class: in somar_ref_b200 the operators of every AMR level are
    B200PoissonOps (built with their coarser level's grids, PoissonOp.cpp:83-120), and the reference's OWN AMRHybridSolver --
    with the LevelHybridSolvers / MGSolvers it builds for its level solves (AMRHybridSolver.cpp:660-760) -- drives them through
    the virtual interface: relaxations, residuals, restriction / prolongation and norms of every level solve run on the device
    (refined levels with their homogeneous coarse-fine ghosts), the inhomogeneous coarse-fine operators stay with the base
    class.  Same status, same composite residual history, same pressure on both levels as the all-CPU reference."""
    from amr_cases import AMR_CASES, composite_rhs_levels, num_levels, ref_kwargs_amr3
    c = AMR_CASES[name]
    nl = num_levels(c)
    rhs, _ = composite_rhs_levels(c, 3)
    a = run_ref("amr", inp=rhs, **ref_kwargs_amr3(c))
    b = run_ref("amr", inp=rhs, shim=True, **ref_kwargs_amr3(c))
    assert int(b.kv["status"]) == int(a.kv["status"])
    na, nb_ = a["amrLevelNorms"].reshape(-1, nl), b["amrLevelNorms"].reshape(-1, nl)
    ca, cb = np.sqrt((na ** 2).sum(axis=1)), np.sqrt((nb_ ** 2).sum(axis=1))
    assert len(ca) == len(cb)
    # the level solves inside an AMR iteration stop on a tolerance, so rounding-level differences of the smoothers come back
    # as ~1e-7 of each composite residual (measured: 2.8e-10 of the initial one)
    assert np.all(np.abs(ca - cb) <= 1e-9 * ca[0])
    assert np.all(np.abs(ca - cb) <= 1e-6 * ca + 1e-11 * ca[0])
    # ... and as ~3e-8 of the pressure (measured), well inside the solver's own tolerance (relTol 1e-6).  The library's own
    # AMRHybridSolver (tests/test_amr_gpu.py), whose smoothers follow the reference's operation order, meets 1e-9.
    errs = [rel_err(b[f"phi{l}"], a[f"phi{l}"]) for l in range(nl)]
    assert max(errs) <= 2e-7, errs
