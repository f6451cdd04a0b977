"""GPU parity for the Helmholtz form of the operator, L = J (alpha + beta Lap) with alpha != 0
(PoissonOp::setAlphaAndBeta, PoissonOp.cpp:707-718): the form the reference's implicit viscous /
diffusive solves take (SURVEY.md 8 row f4 builds on it).  No null space here, so removeKernel is a
no-op and the coefficient tables carry alpha.  Oracle: the reference's own PoissonOp constructed with
the same alpha, beta (oracle driver keys drv.alpha / drv.beta); tolerances as in test_parity_gpu.py."""
import numpy as np
import pytest

import somar_b200 as sb
from _oracle import have_ref, run_ref
from cases import CASES, geometry, rand_field, ref_kwargs, rel_err
from test_parity_gpu import assert_norms

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_ref(3), reason="oracle/_ref/d3/somar_ref not built")]

ALPHA, BETA = 1.0, -0.05
EXTRA = {"drv.alpha": ALPHA, "drv.beta": BETA}


def _op(ctx, c):
    nx, L, dXi, lo, hi = geometry(c)
    blo, bhi = sb.make_base_grids(lo, hi, c["max_box"], (1, 1, 0), c["bf"])
    xmin = lo * dXi
    kind = sb.MAP_CARTESIAN if c["map"] == "cartesian" else sb.MAP_STRETCHED
    return sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, periodic=c["periodic"], map_kind=kind, map_xmin=xmin, map_xmax=xmin + L,
                        map_ampl=c["ampl"], relax_method=c["relax"], alpha=ALPHA, beta=BETA)


@pytest.mark.parametrize("name", ["line_stretch", "line_perx", "gsrb_cart", "gsrb_stretch"])
def test_helmholtz_apply_relax_solve(ctx, name):
    c = CASES[name]
    op = _op(ctx, c)
    phi0, rhs0 = rand_field(c, 1), rand_field(c, 4)
    ref = run_ref("applyop", inp=[phi0], extra=EXTRA, **ref_kwargs(c))
    assert op.has_null_space == bool(ref.kv["hasNullSpace"]) is False
    assert rel_err(op.coefficient(1), ref["Dinv"]) <= 1e-15
    phi, lhs = op.field(data=phi0), op.field()
    op.applyOp(lhs, phi)
    assert rel_err(lhs.download(), ref["lhs"]) <= 1e-14

    ref = run_ref("relax", inp=[phi0, rhs0], extra=dict(EXTRA, **{"drv.relaxIters": 3}), **ref_kwargs(c))
    phi, rhs = op.field(data=phi0), op.field(data=rhs0)
    op.relax(phi, rhs, 3)
    assert rel_err(phi.download(), ref["phi"]) <= 1e-12

    ref = run_ref("solve", inp=[rhs0], extra=EXTRA, **ref_kwargs(c))
    solver = sb.LevelHybridSolver(op, sb.default_options())
    out = op.field()
    st = solver.solve(out, rhs)
    assert st.status == int(ref.kv["status"])
    assert st.max_depth == int(ref.kv["maxDepth"])
    assert_norms(st.norms, ref["norms"][1:])
    assert rel_err(out.download(), ref["phi"]) <= 1e-9
    solver.free()
    op.free()
