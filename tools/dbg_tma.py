"""Development aid: one line relaxation on a small grid with the TMA kernel forced (SB_LINE_TMA=force)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SB_LINE_TMA"] = "force"
import somar_b200 as sb

nx = tuple(int(v) for v in (sys.argv[1:4] or (128, 8, 32)))
ctx = sb.Context(0, 0, 1)
nxa = np.array(nx)
L = np.array([4.0, 2.0, 1.0])
dXi = L / nxa
lo = np.array([0, 0, -nx[2]])
hi = lo + nxa - 1
op = sb.PoissonOp(ctx, lo, hi, dXi, lo[None, :], hi[None, :], relax_method=sb.RELAX_VERTLINE)
rng = np.random.default_rng(1)
phi, rhs = op.field(data=rng.standard_normal(nx)), op.field(data=rng.standard_normal(nx))
try:
    op.relax(phi, rhs, 2)
    ctx.sync()
    print("ok", float(np.abs(phi.download()).max()))
except Exception as e:
    print("FAILED:", e)
