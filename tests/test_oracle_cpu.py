"""CPU tests of the oracle itself (run where oracle/_ref is built): it must reproduce the committed
golden vectors bit for bit (guards against drift of the restated Fortran leaves), solve the
authors' own manufactured-solution test (AMRNSLevelProject.cpp:385-412, disabled in the reference)
and satisfy the run-time invariants the reference prints (L[1] = 0, |Div U| drop)."""
import ast
import glob
import os

import numpy as np
import pytest

from _oracle import have_ref, run_ref
from cases import ref_kwargs

pytestmark = pytest.mark.skipif(not have_ref(3), reason="oracle/_ref/d3/somar_ref not built")
HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_oracle_reproduces_golden(path):
    z = np.load(path)
    c = ast.literal_eval(str(z["case"]))
    r = run_ref("solve", inp=[z["solve_rhs"]], **ref_kwargs(c))
    assert np.array_equal(r["phi"], z["solve_phi"].ravel(order="F"))
    assert np.array_equal(r["norms"], z["solve_norms"])
    r = run_ref("applyop", inp=[z["apply_in"]], **ref_kwargs(c))
    assert np.array_equal(r["lhs"], z["apply_out"].ravel(order="F"))


@pytest.mark.parametrize("relax", [5, 6])
def test_manufactured_solution(relax):
    # sol = cos(kx x) cos(ky y) 16 z^2 (z+1)^2 on [0,L]^2 x [-1,0]; rhs = L[sol]; solve; compare up to a constant
    nx, L = (32, 32, 32), (1.0, 1.0, 1.0)
    c = dict(nx=nx, L=L, max_box=(16, 16, 0), bf=4, periodic=(0, 0, 0), relax=relax, map="cartesian", ampl=(0, 0, 0))
    x = (np.arange(nx[0]) + 0.5) / nx[0]
    y = (np.arange(nx[1]) + 0.5) / nx[1]
    z = (np.arange(-nx[2], 0) + 0.5) / nx[2]
    k = 2 * 2 * np.pi
    sol = np.cos(k * x)[:, None, None] * np.cos(k * y)[None, :, None] * (16 * z**2 * (z + 1) ** 2)[None, None, :]
    sol -= sol.mean()
    rhs = run_ref("applyop", inp=[sol], **ref_kwargs(c))["lhs"]
    assert abs(rhs.sum()) <= 1e-9 * np.abs(rhs).sum()          # solvability: Sum[rhs] = 0 (PoissonOp.cpp:853-885)
    r = run_ref("solve", inp=[rhs], extra={"proj.relTol": 1e-12, "proj.absTol": 1e-14, "proj.maxIters": 20}, **ref_kwargs(c))
    assert int(r.kv["status"]) in (1, 4)
    phi = r["phi"].reshape(nx, order="F")
    phi -= phi.mean()
    assert np.max(np.abs(phi - sol)) <= 1e-9 * np.max(np.abs(sol))


def test_null_space_and_projection_invariants():
    c = dict(nx=(16, 16, 8), L=(1.0, 1.0, 1.0), max_box=(8, 8, 0), bf=4, periodic=(0, 0, 0), relax=6, map="cartesian", ampl=(0, 0, 0))
    r = run_ref("applyop", inp=[np.ones(c["nx"])], **ref_kwargs(c))
    assert np.max(np.abs(r["lhs"])) <= 1e4 * np.finfo(float).eps     # L[1] = 0 to smallReal (PoissonOp.cpp:670-696)
    assert int(r.kv["hasNullSpace"]) == 1
    z = np.load(os.path.join(HERE, "golden", "g_line_cart.npz"))
    assert float(z["proj_finalDivNorm"]) <= 1e-5 * float(z["proj_initDivNorm"])


def _truncation_error(n, ampl):
    """max |L_h[u] - J Lap u| / max |J Lap u| for a smooth u with homogeneous Neumann walls on an n^3
    grid of the stretched map x = xi + A sin(2 pi (xi - xmin) / L) (maps/StretchedMap.cpp:9-38)."""
    nx, L = (n, n, n), (2.0, 1.0, 1.5)
    c = dict(nx=nx, L=L, max_box=(n // 2, n // 2, 0), bf=4, periodic=(0, 0, 0), relax=5, map="stretched", ampl=ampl)
    lo = np.array([0.0, 0.0, -L[2]])           # offset (0, 0, -n) cells: z in [-Lz, 0]
    xs, dxs = [], []
    for d in range(3):
        xi = lo[d] + (np.arange(nx[d]) + 0.5) * L[d] / nx[d]
        k = 2 * np.pi / L[d]
        xs.append(xi + ampl[d] * np.sin(k * (xi - lo[d])))
        dxs.append(1.0 + ampl[d] * k * np.cos(k * (xi - lo[d])))
    m = (2, 1, 3)                                # wall-compatible cosine modes per direction
    w = [m[d] * np.pi / L[d] for d in range(3)]
    cosx = [np.cos(w[d] * (xs[d] - lo[d])) for d in range(3)]
    u = cosx[0][:, None, None] * cosx[1][None, :, None] * cosx[2][None, None, :]
    lap = -(w[0] ** 2 + w[1] ** 2 + w[2] ** 2) * u
    r = run_ref("applyop", inp=[np.asfortranarray(u)], **ref_kwargs(c))
    J = r["J"].reshape(nx, order="F")
    Jexact = dxs[0][:, None, None] * dxs[1][None, :, None] * dxs[2][None, None, :]
    lhs = r["lhs"].reshape(nx, order="F")
    return (np.max(np.abs(lhs - J * lap)) / np.max(np.abs(J * lap)), np.max(np.abs(J - Jexact)) / np.max(Jexact))


@pytest.mark.parametrize("ampl", [(0.0, 0.0, 0.0), (0.08, 0.04, -0.1)])
def test_operator_is_second_order_accurate_on_the_mapped_grid(ampl):
    """An independent pin of the restated coefficient / stencil / boundary leaves (COMPUTEMATRIXELEMENTS,
    COMPUTEDINV, APPLYOP, FILLGHOSTCELLS, the metric fills): against the CONTINUOUS operator J Lap u the
    discrete one must converge at second order, walls included.  A wrong coefficient, index shift or
    ghost formula in any of them shows as O(1) or first-order error."""
    e16, j16 = _truncation_error(16, ampl)
    e32, j32 = _truncation_error(32, ampl)
    e64, j64 = _truncation_error(64, ampl)
    assert e16 < 0.2 and e64 < 0.02
    assert 3.3 <= e16 / e32 <= 4.8 and 3.5 <= e32 / e64 <= 4.5
    if any(ampl):
        assert 3.3 <= j16 / j32 <= 4.8 and 3.5 <= j32 / j64 <= 4.5     # the metric is a centred difference of the map
    else:
        assert j16 == 0.0 and j64 == 0.0


def _divgrad_errors(n, ampl):
    """Relative max errors of levelDivergence and levelGradient against the continuous J div(u) and
    J (dxi/dx) dphi/dx on an n^3 grid of the stretched map."""
    nx, L = (n, n, n), (2.0, 1.0, 1.5)
    c = dict(nx=nx, L=L, max_box=(n // 2, n // 2, 0), bf=4, periodic=(0, 0, 0), relax=5, map="stretched", ampl=ampl)
    lo = np.array([0.0, 0.0, -L[2]])
    k = [2 * np.pi / L[d] for d in range(3)]
    xi_c = [lo[d] + (np.arange(nx[d]) + 0.5) * L[d] / nx[d] for d in range(3)]
    xi_f = [lo[d] + np.arange(nx[d] + 1) * L[d] / nx[d] for d in range(3)]
    mp = lambda d, xi: xi + ampl[d] * np.sin(k[d] * (xi - lo[d]))                 # x(xi)
    dm = lambda d, xi: 1.0 + ampl[d] * k[d] * np.cos(k[d] * (xi - lo[d]))         # dx/dxi
    w = [2 * np.pi / L[0], np.pi / L[1], 3 * np.pi / L[2]]

    def grid(cent):   # coordinates, dx/dxi on the centring cent (tuple of 0 cell / 1 face per direction)
        xi = [xi_f[d] if cent[d] else xi_c[d] for d in range(3)]
        X = np.meshgrid(*[mp(d, xi[d]) for d in range(3)], indexing="ij")
        D = np.meshgrid(*[dm(d, xi[d]) for d in range(3)], indexing="ij")
        return X, D

    # phi = prod cos(w_d (x_d - lo_d)) (Neumann walls); u_d = sin(w_d (x_d - lo_d)) * cos * cos (no flow through walls)
    def phi_of(X):
        return np.cos(w[0] * (X[0] - lo[0])) * np.cos(w[1] * (X[1] - lo[1])) * np.cos(w[2] * (X[2] - lo[2]))

    Xc, Dc = grid((0, 0, 0))
    Jc = Dc[0] * Dc[1] * Dc[2]
    vel, div_exact = [], np.zeros(nx)
    for d in range(3):
        cent = tuple(1 if e == d else 0 for e in range(3))
        Xf, Df = grid(cent)
        f = [np.cos(w[e] * (Xf[e] - lo[e])) for e in range(3)]
        f[d] = np.sin(w[d] * (Xf[d] - lo[d]))
        Jf = Df[0] * Df[1] * Df[2]
        vel.append(np.asfortranarray(Jf / Df[d] * f[0] * f[1] * f[2]))            # advecting velocity J (dxi/dx) u
        g = [np.cos(w[e] * (Xc[e] - lo[e])) for e in range(3)]
        g[d] = w[d] * np.cos(w[d] * (Xc[d] - lo[d]))
        div_exact += g[0] * g[1] * g[2]
    r = run_ref("divgrad", inp=[np.asfortranarray(phi_of(Xc))] + vel, **ref_kwargs(c))
    e_div = np.max(np.abs(r["div"].reshape(nx, order="F") - Jc * div_exact)) / np.max(np.abs(Jc * div_exact))
    e_grad = 0.0
    for d in range(3):
        cent = tuple(1 if e == d else 0 for e in range(3))
        Xf, Df = grid(cent)
        f = [np.cos(w[e] * (Xf[e] - lo[e])) for e in range(3)]
        f[d] = -w[d] * np.sin(w[d] * (Xf[d] - lo[d]))
        want = Df[0] * Df[1] * Df[2] / Df[d] * f[0] * f[1] * f[2]                 # Jg^{dd} dphi/dxi = J (dxi/dx) dphi/dx
        got = r[f"grad{d}"].reshape(want.shape, order="F")
        e_grad = max(e_grad, np.max(np.abs(got - want)) / np.max(np.abs(want)))
    return e_div, e_grad


@pytest.mark.parametrize("ampl", [(0.0, 0.0, 0.0), (0.08, 0.04, -0.1)])
def test_div_and_grad_are_second_order_accurate(ampl):
    """Same kind of pin for the face-centred leaves (FINITEDIFF_DIV3D, FINITEDIFF_PARTIALD_CC2NC, the Jgup
    fill, the wall ghost fill used by levelGradient)."""
    e = [_divgrad_errors(n, ampl) for n in (16, 32, 64)]
    for q in (0, 1):
        assert e[2][q] < 0.01
        assert 3.3 <= e[0][q] / e[1][q] <= 4.8 and 3.5 <= e[1][q] / e[2][q] <= 4.5


@pytest.mark.parametrize("name", ["cube", "flat"])
def test_oracle_operator_matches_the_references_own_python_kit(name):
    """The reference ships an independent implementation of the same operator: PythonScripts/ElliKit.py, the sparse
    Div . Grad Laplacian with mirror (Neumann) ghosts its Python-side solvers use.  tests/golden/independent/ellikit_laplacian.npz holds
    L[phi] computed by importing that file (tests/golden/make_golden_ellikit.py, run where /root/reference exists).  On a
    Cartesian map J = 1 and HomogNeumBC gives the same ghosts, so the oracle's applyOp -- i.e. the restated
    COMPUTEMATRIXELEMENTS / APPLYOP / FILLGHOSTCELLS leaves behind the reference's C++ -- must reproduce it to rounding,
    with several boxes per direction (the exchange path) as well as with one."""
    z = np.load(os.path.join(HERE, "golden", "independent", "ellikit_laplacian.npz"))
    nx, L, phi, want = tuple(int(v) for v in z[f"{name}_nx"]), tuple(float(v) for v in z[f"{name}_L"]), z[f"{name}_phi"], z[f"{name}_lap"]
    scale = np.max(np.abs(want))
    for max_box in ((0, 0, 0), (nx[0] // 2, nx[1] // 2, 0)):
        r = run_ref("applyop", nx=nx, L=L, inp=[phi], max_box=max_box, block_factor=2, relax=5)
        got = r["lhs"].reshape(nx, order="F")
        assert np.max(np.abs(got - want)) <= 1e-12 * scale
