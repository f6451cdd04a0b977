#!/usr/bin/env python
"""Golden vectors from the reference's OWN Python operator kit, PythonScripts/ElliKit.py (the sparse Laplacian / restriction
builder SOMAR's Python-side solvers use) -- code written by the reference's authors, independent of the Fortran leaves and of
this repository's restatement of them (oracle/fort_leaves.cpp).

ElliKit builds L = sum_d Div_d Grad_d on a uniform grid with mirror ghosts for 'Neumann' sides (ghost = first interior cell).
On a Cartesian map SOMAR's PoissonOp computes J * Lap(phi) with J = 1 and, for the projector's HomogNeumBC, the same ghosts, so
the two must agree to rounding.  The fixture pins the oracle's ApplyOp + matrix-element + Neumann ghost-fill leaves
(PoissonOpF.ChF:78-199, BCToolsF.ChF:222-337) against genuinely independent reference code.  (ElliKit's restriction operator is
work in progress upstream -- "eventually, this must be filled with 1/Jinv" -- and is not used.)  Run here only (needs /root/reference); the .npz travels.

    python tests/golden/make_golden_ellikit.py
"""
import importlib.util
import os
import sys

import numpy as np

REF = os.environ.get("SOMAR_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_ellikit():
    spec = importlib.util.spec_from_file_location("ElliKit", os.path.join(REF, "PythonScripts", "ElliKit.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.Ellikit


def main():
    Ellikit = load_ellikit()
    out = {}
    rng = np.random.default_rng(20250829)
    # (name, nx, L): anisotropic spacings, all sides Neumann
    for name, nx, L in [("cube", (16, 12, 8), (1.0, 1.5, 0.5)), ("flat", (24, 8, 4), (8.0, 2.0, 0.25))]:
        dx = tuple(L[d] / nx[d] for d in range(3))
        phi = np.asfortranarray(rng.standard_normal(nx))
        kit = Ellikit(((0, 0, 0), nx, (0, 0, 0)), dx, BC=("Neumann",) * 6)
        lap = np.asarray(kit.Dot(phi, "Laplacian")).reshape(nx, order="F")
        out[f"{name}_nx"] = np.array(nx)
        out[f"{name}_L"] = np.array(L)
        out[f"{name}_phi"] = phi
        out[f"{name}_lap"] = np.asfortranarray(lap)
    np.savez_compressed(os.path.join(HERE, "independent", "ellikit_laplacian.npz"), **out)
    print("wrote", os.path.join(HERE, "independent", "ellikit_laplacian.npz"), sorted(out))


if __name__ == "__main__":
    main()
