"""Multi-GPU parity (needs >= 2 GPUs on the box; run with `gpurun --gpus 2`): the same solves as
test_parity_gpu.py with the boxes split over 2 ranks (NCCL halo exchange + scalar all-reduces),
checked against the oracle with the same tolerances."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from _oracle import have_ref, run_ref
from cases import CASES, rand_field, ref_kwargs, rel_err
from test_parity_gpu import V_OPTS, _proj_overrides, assert_norms


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(_ngpu() < int(os.environ.get("SB_TEST_WORLD", "2")), reason="needs SB_TEST_WORLD (default 2) GPUs"),
              pytest.mark.skipif(not have_ref(3), reason="oracle/_ref/d3/somar_ref not built")]

HERE = os.path.dirname(os.path.abspath(__file__))
WORLD = int(os.environ.get("SB_TEST_WORLD", "2"))  # ranks = GPUs used by the worker


# agg: SB_AGG_CELLS, the size below which MG depths are agglomerated onto rank 0 (None = library default, which puts every
# depth >= 1 of these small cases on rank 0; "0" = every depth distributed; a number in between = a deeper switch-over).
@pytest.mark.parametrize("name,optset,agg", [("line_cart", "defaults", None), ("line_stretch", "defaults", None),
                                             ("line_perx", "defaults", None), ("gsrb_cart", "defaults", None),
                                             ("gsrb_perxy", "vcycle", None), ("line_aniso", "vcycle", None),
                                             ("line_cart", "defaults", "0"), ("line_stretch", "defaults", "0"),
                                             ("line_perx", "defaults", "0"), ("gsrb_cart", "defaults", "0"),
                                             ("gsrb_perxy", "vcycle", "0"), ("line_aniso", "vcycle", "0"),
                                             ("line_stretch", "defaults", "300"), ("line_aniso", "vcycle", "1100"),
                                             ("gsrb_cart", "vcycle", "600"),
                                             ("lep3d_cart", "defaults", None), ("lep3d_perx", "defaults", "0"),
                                             ("lepmg3d_cart", "defaults", None), ("lepmg3d_zstretch", "defaults", "0"),
                                             # 2-D build (ranks split x): MG, periodic x, and the DJL deck's base level (leptic)
                                             ("gsrb2d_stretch", "defaults", None), ("line2d_zstretch_perx", "vcycle", "0"),
                                             ("c2_djl_base", "defaults", None)])
def test_two_rank_solve(name, optset, agg):
    import test_parity2d_gpu as t2
    from test_leptic_gpu import DJL_OPTS, LEPTIC2D
    if name in t2.CASES2D or name in LEPTIC2D:
        return _two_rank_solve_2d(name, optset, agg)
    leptic = name not in CASES
    if leptic:
        from test_leptic_gpu import LEPTIC3D
    c = LEPTIC3D[name][0] if leptic else CASES[name]
    import somar_b200 as sb
    from cases import geometry
    _, _, _, lo, hi = geometry(c)
    if len(sb.make_base_grids(lo, hi, c["max_box"], (1, 1, 0), c["bf"])[0]) < WORLD:
        pytest.skip(f"{name} has fewer boxes than ranks ({WORLD})")
    ref = run_ref("solve", inp=[rand_field(c, 4, zero_mean=True)], extra=_proj_overrides({} if optset == "defaults" else V_OPTS),
                  **ref_kwargs(c))
    res, phi = _run_worker(name, optset, agg)
    assert res["status"] == int(ref.kv["status"])
    assert res["max_depth"] == int(ref.kv["maxDepth"])
    if leptic:
        assert res["solve_mode"] == int(ref.kv["solveMode"]) == LEPTIC3D[name][1]
        assert_norms(res["norms"], ref["hybridNorms"])
    else:
        assert_norms(res["norms"], ref["norms"][1:])
    assert rel_err(phi, ref["phi"]) <= 1e-9


def _two_rank_solve_2d(name, optset, agg):
    import test_parity2d_gpu as t2
    from test_leptic_gpu import DJL_OPTS, LEPTIC2D
    if not have_ref(2):
        pytest.skip("oracle/_ref/d2/somar_ref not built")
    leptic = name in LEPTIC2D
    c = LEPTIC2D[name][0] if leptic else t2.CASES2D[name]
    if c["nx"][0] // c["max_box"][0] % WORLD:
        pytest.skip(f"{name}: box count not divisible by {WORLD} ranks")
    over = DJL_OPTS if name.startswith("c2_") else ({} if optset == "defaults" else V_OPTS)
    ref = run_ref("solve", inp=[t2.rand_field(c, 4, zero_mean=True)], extra=_proj_overrides(over), **t2.ref_kwargs(c))
    res, phi = _run_worker(name, optset, agg)
    assert res["status"] == int(ref.kv["status"])
    assert res["max_depth"] == int(ref.kv["maxDepth"])
    if leptic:
        assert res["solve_mode"] == int(ref.kv["solveMode"]) == LEPTIC2D[name][1]
        assert_norms(res["norms"], ref["hybridNorms"])
    else:
        assert_norms(res["norms"], ref["norms"][1:])
    assert rel_err(phi, ref["phi"]) <= 1e-9


def _run_worker(name, optset, agg):
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "res.json")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={WORLD}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.join(HERE, "mgpu_worker.py"), name, optset, out]
        env = dict(os.environ)
        env.pop("SB_AGG_CELLS", None)
        if agg is not None:
            env["SB_AGG_CELLS"] = agg
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        res = json.load(open(out))
        phi = np.load(out + ".phi.npy")
    return res, phi
