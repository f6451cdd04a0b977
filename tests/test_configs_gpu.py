"""GPU parity on the grids of BASELINE.json's 3-D exec decks (SURVEY.md 8, config matrix):

* `c4_doublediffusion`: exec/DoubleDiffusion/inputs.3D.SaltFingers in full -- 128^3, L = 0.5^3,
  single level, base.maxBaseGridSize 64^3 (boxes split in z too), GSRB, all-Neumann; the deck lacks
  base.blockFactor (SURVEY C4 caveat), 16 is used.
* `c3_bvr_base`: the base level of exec/BuoyantVortexRing/inputs -- 64^3, L = 10^3, nxOffset -32,
  triply periodic (null space), 64 boxes of 16^3, GSRB.  The deck's two refined levels (AMR
  composite solve) are SURVEY row f2.

Both take LevelHybridSolver's MG branch with SemicoarseningStrategy.  Oracle: the reference's own
C++ (oracle/_ref/d3) on the same seeded inputs; tolerances as in test_parity_gpu.py."""
import numpy as np
import pytest

import somar_b200 as sb
from _oracle import have_ref, run_ref
from cases import rel_err
from test_parity_gpu import assert_norms

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_ref(3), reason="oracle/_ref/d3/somar_ref not built")]

CONFIGS = {
    "c4_doublediffusion": dict(nx=(128, 128, 128), L=(0.5, 0.5, 0.5), offset=(0, 0, 0), max_box=(64, 64, 64), bf=16,
                               periodic=(0, 0, 0)),
    "c3_bvr_base": dict(nx=(64, 64, 64), L=(10.0, 10.0, 10.0), offset=(-32, -32, -32), max_box=(16, 16, 16), bf=16,
                        periodic=(1, 1, 1)),
}


def _setup(ctx, c):
    nx = np.array(c["nx"])
    dXi = np.array(c["L"]) / nx
    lo = np.array(c["offset"])
    hi = lo + nx - 1
    blo, bhi = sb.make_base_grids(lo, hi, c["max_box"], (1, 1, 1), c["bf"])
    op = sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, periodic=c["periodic"], relax_method=sb.RELAX_GSRB)
    kw = dict(nx=c["nx"], L=c["L"], max_box=c["max_box"], block_factor=c["bf"], offset=c["offset"], periodic=c["periodic"],
              relax=5, split_dirs=(1, 1, 1))
    return op, kw


def _wall_velocity(c, seed):
    rng = np.random.default_rng(seed)
    out = []
    for d in range(3):
        shape = list(c["nx"])
        shape[d] += 1
        u = rng.standard_normal(tuple(shape))
        lo, hi = [slice(None)] * 3, [slice(None)] * 3
        lo[d], hi[d] = 0, -1
        if c["periodic"][d]:
            u[tuple(hi)] = u[tuple(lo)]
        else:
            u[tuple(lo)] = 0.0
            u[tuple(hi)] = 0.0
        out.append(np.asfortranarray(u))
    return out


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_config_projection(ctx, name):
    """AMRNSLevel::projectCorrect's bracket (divergence, level solve, gradient, correction) with the
    deck's solver defaults (FMG, 16 + 16 smooths, BiCGStab bottom)."""
    c = CONFIGS[name]
    op, kw = _setup(ctx, c)
    vel0 = _wall_velocity(c, 5)
    ref = run_ref("project", inp=vel0, timeout=1500, **kw)
    assert int(ref.kv["solveMode"]) == 1
    assert op.has_null_space == bool(ref.kv["hasNullSpace"]) if "hasNullSpace" in ref.kv else True
    solver = sb.LevelHybridSolver(op, sb.default_options())
    vel, phi, n0, n1, st = solver.project_host(vel0)
    assert abs(n0 - ref.kv["initDivNorm"]) <= 1e-13 * ref.kv["initDivNorm"]
    assert st.status == int(ref.kv["status"])
    assert st.max_depth == int(ref.kv["maxDepth"])
    assert_norms(st.norms, ref["norms"][1:])
    assert rel_err(phi, ref["phi"]) <= 1e-9
    for d in range(3):
        assert rel_err(vel[d], ref[f"vel{d}"]) <= 1e-9
    assert abs(n1 - ref.kv["finalDivNorm"]) <= 1e-6 * max(ref.kv["finalDivNorm"], 1e-30) + 1e-12 * n0
    solver.free()
    op.free()
