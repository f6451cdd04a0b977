import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import somar_b200 as sb
from _oracle import run_ref
from cases import CASES, make_op, rand_field, ref_kwargs
from test_parity_gpu import V_OPTS, _proj_overrides
np.set_printoptions(linewidth=200, precision=12)
ctx = sb.Context(0, 0, 1)
c = CASES["line_stretch"]
rhs0 = rand_field(c, 4, zero_mean=True)
over = dict(V_OPTS, prolongOrder=0)
ref = run_ref("solve", inp=[rhs0], extra=_proj_overrides(over), **ref_kwargs(c))
op = make_op(ctx, c)
solver = sb.LevelHybridSolver(op, sb.default_options(**over))
phi, rhs = op.field(), op.field(data=rhs0)
st = solver.solve(phi, rhs)
print("ours", st.status, np.array(st.norms)); print("ref ", int(ref.kv["status"]), ref["norms"][1:])
