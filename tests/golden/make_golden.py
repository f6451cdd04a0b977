#!/usr/bin/env python
"""Generates tests/golden/*.npz from the reference itself: oracle/_ref/d3/somar_ref is SOMAR's own
unmodified C++ solver stack (PoissonOp, MGSolver, BiCGStab, LevelHybridSolver) built from
/root/reference by oracle/build_ref.sh.  Run in the build container (needs the oracle binary):

    python tests/golden/make_golden.py

Each fixture holds the seeded input and everything the reference produced for it, so the GPU tests
can check parity on a box where neither /root/reference nor the oracle binary exists."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _oracle import run_ref  # noqa: E402
from cases import rand_field, rand_velocity, ref_kwargs  # noqa: E402

GOLDEN = {
    "g_line_cart": dict(nx=(16, 16, 8), L=(1.0, 1.0, 1.0), max_box=(8, 8, 0), bf=4, periodic=(0, 0, 0), relax=6, map="cartesian", ampl=(0, 0, 0)),
    "g_line_stretch": dict(nx=(16, 24, 8), L=(1.0, 1.5, 1.0), max_box=(8, 8, 0), bf=4, periodic=(0, 0, 0), relax=6, map="stretched", ampl=(0.05, 0.03, -0.1)),
    "g_line_zstretch_perx": dict(nx=(32, 16, 16), L=(4.0, 2.0, 1.0), max_box=(16, 16, 0), bf=8, periodic=(1, 0, 0), relax=6, map="stretched", ampl=(0.0, 0.0, -0.12)),
    "g_gsrb_cart": dict(nx=(16, 16, 16), L=(1.0, 1.0, 6.0), max_box=(8, 8, 0), bf=4, periodic=(0, 0, 0), relax=5, map="cartesian", ampl=(0, 0, 0)),
    "g_gsrb_stretch_pery": dict(nx=(16, 16, 16), L=(1.0, 1.0, 6.0), max_box=(8, 16, 0), bf=4, periodic=(0, 1, 0), relax=5, map="stretched", ampl=(0.04, 0.0, 0.3)),
}


def main():
    for name, c in GOLDEN.items():
        out = {"case": np.array(repr(c))}
        phi0, rhs0 = rand_field(c, 11), rand_field(c, 12, zero_mean=True)
        r = run_ref("applyop", inp=[phi0], **ref_kwargs(c))
        out.update(apply_in=phi0, apply_out=r["lhs"], apply_norm2=r.kv["norm2"], apply_norm0=r.kv["norm0"],
                   hasNullSpace=int(r.kv["hasNullSpace"]), J=r["J"], Dinv=r["Dinv"], M0=r["M0"], M1=r["M1"], M2=r["M2"])
        r = run_ref("relax", inp=[phi0, rhs0], extra={"drv.relaxIters": 2}, **ref_kwargs(c))
        out.update(relax_phi_in=phi0, relax_rhs=rhs0, relax_out=r["phi"])
        r = run_ref("solve", inp=[rhs0], **ref_kwargs(c))
        out.update(solve_rhs=rhs0, solve_phi=r["phi"], solve_norms=r["norms"], solve_status=int(r.kv["status"]),
                   solve_maxDepth=int(r.kv["maxDepth"]), solve_mode=int(r.kv["solveMode"]))
        vel = rand_velocity(c, 13)
        r = run_ref("project", inp=vel, **ref_kwargs(c))
        out.update(proj_u0=vel[0], proj_u1=vel[1], proj_u2=vel[2], proj_phi=r["phi"], proj_norms=r["norms"],
                   proj_v0=r["vel0"], proj_v1=r["vel1"], proj_v2=r["vel2"], proj_div=r["div"], proj_initDivNorm=r.kv["initDivNorm"],
                   proj_finalDivNorm=r.kv["finalDivNorm"], proj_status=int(r.kv["status"]))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "status", out["solve_status"], "norms", out["solve_norms"])


if __name__ == "__main__":
    main()
