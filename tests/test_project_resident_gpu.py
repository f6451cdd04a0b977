"""The projection bracket on device-resident fields: sb_project_correct and sb_project_predict
-- AMRNSLevel::projectCorrect (AMRNSLevelProject.cpp:247-373) and ::projectPredict (:63-243) between the level's own BC
fills -- against the oracle (modes `project` and `predict` of oracle/ref_driver.cpp, which restate those two Grade5
functions around the reference's own operator and LevelHybridSolver calls).  Velocity and pressure never leave the
device between the calls; only divergence norms and the solver status come back."""
import numpy as np
import pytest

import somar_b200 as sb
from _oracle import have_ref, run_ref
from cases import CASES, make_op, rand_field, rand_velocity, ref_kwargs, rel_err
from test_parity_gpu import assert_norms

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_ref(3), reason="oracle/_ref/d3/somar_ref not built")]


@pytest.mark.parametrize("name", ["line_stretch", "gsrb_perxy", "s_line64"])
def test_project_correct_resident(ctx, name):
    c = CASES[name]
    op = make_op(ctx, c)
    vel0 = rand_velocity(c, 5)
    p0 = rand_field(c, 6)
    dt = 0.25
    ref = run_ref("project", inp=vel0, **ref_kwargs(c))
    solver = sb.LevelHybridSolver(op, sb.default_options())
    vel, p, phi = op.flux(vel0), op.field(data=p0), op.field()
    n0, n1, st = solver.project_correct(vel, p=p, proj_dt=dt, vel_ghost=-1, phi_out=phi)
    assert abs(n0 - ref.kv["initDivNorm"]) <= 1e-13 * ref.kv["initDivNorm"]
    assert st.status == int(ref.kv["status"])
    assert_norms(st.norms, ref["norms"][1:])
    assert rel_err(phi.download(), ref["phi"]) <= 1e-9
    for d in range(3):
        assert rel_err(vel[d].download(), ref[f"vel{d}"]) <= 1e-9
    want_p = p0 + ref["phi"].reshape(c["nx"], order="F") * (1.0 / dt)           # p.plus(phi, 1 / projDt)
    assert rel_err(p.download(), want_p) <= 1e-9
    assert abs(n1 - ref.kv["finalDivNorm"]) <= 1e-6 * max(ref.kv["finalDivNorm"], 1e-30) + 1e-12 * n0
    solver.free()
    op.free()


@pytest.mark.parametrize("name,kind", [("line_stretch", "random_p"), ("gsrb_perxy", "random_p"), ("line_stretch", "zero_p"),
                                       ("line_cart", "consistent_p")])
def test_project_predict_resident(ctx, name, kind):
    """random_p: the lagged pressure makes the divergence worse -> the quick-and-dirty projectCorrect fallback runs;
    zero_p: nothing changes, no fallback; consistent_p: p = phi / dt of the exact projection -> the lagged correction alone
    removes the divergence."""
    c = CASES[name]
    dt = 0.5
    vel0 = rand_velocity(c, 5)
    if kind == "random_p":
        p0 = rand_field(c, 6)
    elif kind == "zero_p":
        p0 = np.zeros(c["nx"], order="F")
    else:
        p0 = np.asfortranarray(run_ref("project", inp=vel0, **ref_kwargs(c))["phi"].reshape(c["nx"], order="F") / dt)
    ref = run_ref("predict", inp=vel0 + [p0], extra={"drv.projDt": dt}, **ref_kwargs(c))
    op = make_op(ctx, c)
    solver = sb.LevelHybridSolver(op, sb.default_options())
    vel, p = op.flux(vel0), op.field(data=p0)
    norms, fallback, st = solver.project_predict(vel, p, proj_dt=dt, vel_ghost=-1)
    assert fallback == bool(ref.kv["usedFallback"])
    assert fallback == (kind == "random_p")
    assert abs(norms[0] - ref.kv["initDivNorm"]) <= 1e-13 * ref.kv["initDivNorm"]
    assert abs(norms[1] - ref.kv["laggedDivNorm"]) <= 1e-9 * ref.kv["initDivNorm"] + 1e-12 * ref.kv["laggedDivNorm"]
    if fallback:
        assert st.status == int(ref.kv["status"])
        assert abs(norms[2] - ref.kv["correctedDivNorm"]) <= 1e-9 * ref.kv["laggedDivNorm"]
    else:
        assert norms[2] == -1.0
    scale = max(np.max(np.abs(ref[f"vel{d}"])) for d in range(3))
    for d in range(3):
        assert np.max(np.abs(vel[d].download().ravel(order="F") - ref[f"vel{d}"])) <= 1e-9 * scale
    assert rel_err(p.download(), ref["p"]) <= 1e-9
    # the solver's own options are back after the fallback: a full solve converges as before
    phi, rhs = op.field(), op.field(data=rand_field(c, 4, zero_mean=True))
    st2 = solver.solve(phi, rhs)
    ref2 = run_ref("solve", inp=[rand_field(c, 4, zero_mean=True)], **ref_kwargs(c))
    assert st2.status == int(ref2.kv["status"]) and st2.num_norms == len(ref2["norms"][1:])
    solver.free()
    op.free()
