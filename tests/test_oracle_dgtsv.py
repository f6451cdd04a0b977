"""Pins oracle/dgtsv.c (the restated LAPACK routine the reference's line relaxation calls,
PoissonOpF.ChF:794,957) against scipy's LAPACK dgtsv: diagonally dominant systems of the kind
the solver produces, systems that force row interchanges, and the singular INFO = N case."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest
from scipy.linalg import lapack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dgtsv():
    td = tempfile.mkdtemp()
    so = os.path.join(td, "libdgtsv.so")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", os.path.join(ROOT, "oracle", "dgtsv.c"), "-o", so],
                   check=True)
    lib = C.CDLL(so)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.dgtsv_.argtypes = [ip, ip, dp, dp, dp, dp, ip, ip]
    lib.dgtsv_.restype = None

    def call(dl, d, du, b):
        dl, d, du, b = (np.array(a, dtype=np.float64) for a in (dl, d, du, b))
        n, nrhs, info = C.c_int(len(d)), C.c_int(1), C.c_int(0)
        ldb = C.c_int(max(1, len(d)))
        lib.dgtsv_(C.byref(n), C.byref(nrhs), dl.ctypes.data_as(dp), d.ctypes.data_as(dp), du.ctypes.data_as(dp),
                   b.ctypes.data_as(dp), C.byref(ldb), C.byref(info))
        return b, info.value

    return call


@pytest.mark.parametrize("n", [1, 2, 3, 8, 64, 257])
@pytest.mark.parametrize("kind", ["poisson_column", "random_dominant", "pivoting"])
def test_against_lapack(dgtsv, n, kind):
    rng = np.random.default_rng(1000 * n + len(kind))
    if kind == "poisson_column":
        # the matrix of one grid column: -J(MzL + MzR + h) on the diagonal, J*Mz off it, Neumann ends folded in
        J = 1.0 + 0.3 * rng.random(n)
        ml, mr = 60.0 + rng.random(n), 60.0 + rng.random(n)
        h = 2.0
        d = -J * (ml + mr + h)
        d[0] += J[0] * ml[0]
        d[-1] += J[-1] * mr[-1]
        dl, du = (J * ml)[1:], (J * mr)[:-1]
    elif kind == "random_dominant":
        dl, du = rng.standard_normal(max(n - 1, 0)), rng.standard_normal(max(n - 1, 0))
        d = 4.0 + rng.random(n)
    else:
        dl, du = 5.0 * rng.standard_normal(max(n - 1, 0)), rng.standard_normal(max(n - 1, 0))
        d = 0.1 * rng.standard_normal(n)
    b = rng.standard_normal(n)
    x, info = dgtsv(dl, d, du, b)
    if n == 1:  # scipy's wrapper rejects empty off-diagonals; LAPACK's N = 1 path is b / d
        x_ref, info_ref = b / d, 0
    else:
        _, _, _, x_ref, info_ref = lapack.dgtsv(dl, d, du, b)
    assert info == info_ref == 0
    # same algorithm, same operation order: agreement to the last bits (FMA use inside OpenBLAS may differ)
    np.testing.assert_allclose(x, x_ref, rtol=1e-12 if kind != "pivoting" else 1e-9, atol=1e-14)


def test_singular_returns_info_n(dgtsv):
    # all-Neumann single row (N = 1, zero diagonal): INFO = N, which the reference tolerates (PoissonOpF.ChF:960-963)
    _, info = dgtsv([], [0.0], [], [1.0])
    assert info == 1
    # and the last pivot vanishing after elimination (N = 2)
    _, info = dgtsv([1.0], [1.0, 1.0], [1.0], [1.0, 2.0])
    _, _, _, _, info_ref = lapack.dgtsv(np.array([1.0]), np.array([1.0, 1.0]), np.array([1.0]), np.array([1.0, 2.0]))
    assert info == info_ref == 2
