"""GPU parity against the committed golden vectors (tests/golden/*.npz, produced from the reference
itself by tests/golden/make_golden.py).  Needs no oracle binary and no /root/reference."""
import ast
import glob
import os

import numpy as np
import pytest

import somar_b200 as sb
from cases import make_op, rel_err
from test_parity_gpu import assert_norms

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _load(path):
    z = np.load(path, allow_pickle=False)
    return ast.literal_eval(str(z["case"])), z


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_golden(ctx, path):
    c, z = _load(path)
    op = make_op(ctx, c)
    assert op.has_null_space == bool(z["hasNullSpace"])
    assert rel_err(op.coefficient(0), z["J"]) <= 4e-16
    assert rel_err(op.coefficient(1), z["Dinv"]) <= 1e-15
    for d in range(3):
        assert rel_err(op.coefficient(2 + d), z[f"M{d}"]) == 0.0
    # operator and norms
    phi, lhs = op.field(data=z["apply_in"]), op.field()
    op.applyOp(lhs, phi)
    assert rel_err(lhs.download(), z["apply_out"]) <= 1e-14
    assert abs(op.norm(lhs, 2) - float(z["apply_norm2"])) <= 1e-13 * float(z["apply_norm2"])
    assert abs(op.norm(lhs, 0) - float(z["apply_norm0"])) <= 1e-14 * float(z["apply_norm0"])
    # two relaxation sweeps
    cor, rhs = op.field(data=z["relax_phi_in"]), op.field(data=z["relax_rhs"])
    op.relax(cor, rhs, 2)
    assert rel_err(cor.download(), z["relax_out"]) <= 1e-12
    # full solve with the reference's default options (FMG, 16/16 smooths, BiCGStab bottom)
    solver = sb.LevelHybridSolver(op, sb.default_options())
    p, r = op.field(), op.field(data=z["solve_rhs"])
    st = solver.solve(p, r)
    assert st.status == int(z["solve_status"]) and st.max_depth == int(z["solve_maxDepth"])
    assert_norms(st.norms, z["solve_norms"][1:])
    assert rel_err(p.download(), z["solve_phi"]) <= 1e-9
    # projection bracket
    vel, ph, n0, n1, st = solver.project_host([z["proj_u0"], z["proj_u1"], z["proj_u2"]])
    assert abs(n0 - float(z["proj_initDivNorm"])) <= 1e-13 * float(z["proj_initDivNorm"])
    assert st.status == int(z["proj_status"])
    assert_norms(st.norms, z["proj_norms"][1:])
    assert rel_err(ph, z["proj_phi"]) <= 1e-9
    for d in range(3):
        assert rel_err(vel[d], z[f"proj_v{d}"]) <= 1e-9
    solver.free()
    op.free()


@pytest.mark.parametrize("name", ["g_line_cart", "g_line_zstretch_perx"])
def test_golden_general_line_kernel(name, monkeypatch):
    """The same golden solve through the general (per-column dgtsv-order) line kernel."""
    monkeypatch.setenv("SB_LINE_KERNEL", "general")
    c, z = _load(os.path.join(HERE, "golden", name + ".npz"))
    ctx = sb.Context(0, 0, 1)
    op = make_op(ctx, c)
    solver = sb.LevelHybridSolver(op, sb.default_options())
    p, r = op.field(), op.field(data=z["solve_rhs"])
    st = solver.solve(p, r)
    assert st.status == int(z["solve_status"])
    assert_norms(st.norms, z["solve_norms"][1:])
    assert rel_err(p.download(), z["solve_phi"]) <= 1e-9
    ctx.close()


@pytest.mark.parametrize("name", ["cube", "flat"])
def test_apply_op_matches_the_references_own_python_kit(ctx, name):
    """The CUDA operator against code the reference's authors wrote independently of their Fortran: PythonScripts/ElliKit.py's
    sparse Div . Grad Laplacian with mirror (Neumann) ghosts, imported by tests/golden/make_golden_ellikit.py where
    /root/reference exists.  Cartesian map (J = 1), HomogNeumBC on every side, one box and 2 x 2 boxes."""
    z = np.load(os.path.join(HERE, "golden", "independent", "ellikit_laplacian.npz"))
    nx, L = np.array(z[f"{name}_nx"]), np.array(z[f"{name}_L"], dtype=float)
    phi0, want = z[f"{name}_phi"], z[f"{name}_lap"]
    dXi = L / nx
    lo = np.array([0, 0, -nx[2]])
    hi = lo + nx - 1
    for max_box in ((0, 0, 0), (int(nx[0]) // 2, int(nx[1]) // 2, 0)):
        blo, bhi = sb.make_base_grids(lo, hi, max_box, (1, 1, 0), 2)
        op = sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, relax_method=sb.RELAX_GSRB)
        phi, lhs = op.field(data=np.asfortranarray(phi0)), op.field()
        op.applyOp(lhs, phi)
        assert rel_err(lhs.download(), want) <= 1e-12
        op.free()
