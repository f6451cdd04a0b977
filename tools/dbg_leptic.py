import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import somar_b200 as sb
import test_leptic_gpu as tl, test_parity2d_gpu as t2
from _oracle import run_ref
from test_parity_gpu import _proj_overrides
np.set_printoptions(linewidth=200, precision=10)
ctx = sb.Context(0, 0, 1)
for n in sys.argv[1:]:
    c, m = tl.LEPTIC2D[n]
    over = tl.DJL_OPTS if n.startswith("c2_") else {}
    rhs0 = t2.rand_field(c, 4, zero_mean=True)
    r = run_ref("solve", inp=[rhs0], extra=_proj_overrides(over), **t2.ref_kwargs(c))
    op = t2.make_op(ctx, c)
    solver = sb.LevelHybridSolver(op, sb.default_options(**over))
    phi, rhs = op.field(), op.field(data=t2.up(rhs0))
    st = solver.solve(phi, rhs)
    print(n, "status", st.status, int(r.kv["status"]), "nnorms", st.num_norms, len(r["hybridNorms"]))
    a, b = np.array(st.norms), r["hybridNorms"]
    k = min(len(a), len(b))
    print(" ours", a)
    print(" ref ", b)
    print(" rel diff", np.abs(a[:k] - b[:k]) / b[:k])
    print(" phi rel err", float(np.max(np.abs(phi.download().ravel(order="F") - r["phi"])) / np.max(np.abs(r["phi"]))))
