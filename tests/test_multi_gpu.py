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


def _run_worker(name, optset, agg, extra_env=None):
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "res.json")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={WORLD}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.join(HERE, "mgpu_worker.py"), name, optset, out]
        env = dict(os.environ)
        env.pop("SB_AGG_CELLS", None)
        if agg is not None:
            env["SB_AGG_CELLS"] = agg
        env.update(extra_env or {})
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        res = json.load(open(out))
        phi = np.load(out + ".phi.npy")
    return res, phi


@pytest.mark.parametrize("name,optset", [("line_cart", "defaults"), ("gsrb_perxy", "vcycle")])
def test_peer_halo_is_bitwise_the_nccl_exchange(name, optset):
    """The relaxations' halo exchange by stores into the neighbour's memory (sb_halo.cu; fused into vertline_tma_k where that
    kernel runs) moves the same values as the NCCL send / recv it replaces: the solve is bit-identical either way."""
    import somar_b200 as sb
    from cases import geometry
    c = CASES[name]
    _, _, _, lo, hi = geometry(c)
    if len(sb.make_base_grids(lo, hi, c["max_box"], (1, 1, 0), c["bf"])[0]) < WORLD:
        pytest.skip(f"{name} has fewer boxes than ranks ({WORLD})")
    res_p, phi_p = _run_worker(name, optset, "0")
    res_n, phi_n = _run_worker(name, optset, "0", {"SB_PEER_HALO": "0"})
    assert res_n["halo"] == "nccl" and res_p["halo"] in ("peer", "nccl")
    assert res_p["status"] == res_n["status"] and res_p["norms"] == res_n["norms"]
    assert np.array_equal(phi_p, phi_n)


# ---- AMR hierarchies over several ranks: every level split into tiles, inter-level copies (coarse-fine buffer fill,
# restriction / prolongation between levels, fine flux register) cross ranks through NCCL send/recv (LevelCopier)
@pytest.mark.parametrize("name", ["amr_r2_centre", "amr_r4_periodic", "amr2d_djl_r41", "amr3_c3_mini", "c3_deck"])
def test_multi_rank_amr_solve(name):
    from amr_cases import AMR_CASES, C3_DECK, composite_rhs_levels, level_specs, ndim, num_levels, ref_kwargs_amr3
    import somar_b200 as sb
    c = C3_DECK if name == "c3_deck" else AMR_CASES[name]
    if ndim(c) == 2 and not have_ref(2):
        pytest.skip("oracle/_ref/d2/somar_ref not built")
    for s in level_specs(c):
        try:
            sb.assign_boxes_to_ranks(s["box_lo"], s["box_hi"], WORLD)
        except ValueError:
            pytest.skip(f"{name}: a level's box grid cannot be split over {WORLD} ranks")
    rhs, _ = composite_rhs_levels(c, 3)
    ref = run_ref("amr", inp=rhs, timeout=3000, **ref_kwargs_amr3(c))
    nl = num_levels(c)
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "res.json")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={WORLD}", "--master-addr", "127.0.0.1",
               "--master-port", "29519", os.path.join(HERE, "mgpu_amr_worker.py"), name, out]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ))
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        res = json.load(open(out))
        phis = [np.load(f"{out}.phi{l}.npy") for l in range(nl)]
    assert res["status"] == int(ref.kv["status"])
    lev = ref["amrLevelNorms"].reshape(-1, nl)
    comp = np.sqrt((lev ** 2).sum(axis=1))
    assert len(res["norms"]) == len(comp)
    assert_norms(res["norms"], comp)
    for l in range(nl):
        assert rel_err(phis[l], ref[f"phi{l}"]) <= 1e-9, l
