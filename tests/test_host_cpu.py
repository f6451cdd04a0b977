"""CPU tests (no GPU): the C-ABI library loads and exports every declared symbol, and the host
logic -- box decomposition, MG coarsening schedule -- matches the reference's (values below were
produced by oracle/_ref, i.e. the reference's own MGCoarseningStrategy / makeBaseLevelMesh)."""
import os
import re

import numpy as np
import pytest

import somar_b200 as sb
from somar_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    header = open(os.path.join(ROOT, "include", "somar_b200.h")).read()
    declared = set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", header))
    declared -= {"sb_map_fn"}
    assert len(declared) >= 45
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/somar_b200.h but not exported"
    assert declared == set(capi.PROTOTYPES), declared ^ set(capi.PROTOTYPES)


def test_no_gpu_fails_loudly():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(sb.SomarB200Error, match="no CPU fallback"):
        sb.Context(0, 0, 1)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "somar_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f in (), f"{f} mentions the oracle"


def test_make_base_grids_matches_reference_layout():
    # S5: 1024 x 1024 x 256, maxBaseGridSize 128 128 0, blockFactor 16 -> 64 boxes of 128 x 128 x 256
    lo, hi = sb.make_base_grids((0, 0, -256), (1023, 1023, -1), (128, 128, 0), (1, 1, 0), 16)
    assert lo.shape == (64, 3)
    assert np.all(hi - lo + 1 == np.array([128, 128, 256]))
    assert set(map(tuple, lo)) == {(i, j, -256) for i in range(0, 1024, 128) for j in range(0, 1024, 128)}
    # uneven split: 96 cells, max 64, bf 16 -> 2 boxes of 48 (AnisotropicAMR.cpp:1520-1540 rounds up evenly)
    lo, hi = sb.make_base_grids((0, 0, 0), (95, 31, 7), (64, 0, 0), (1, 1, 0), 16)
    assert sorted((hi - lo + 1)[:, 0].tolist()) == [48, 48]


# (nx, L, max_box, bf, relax) -> schedule; produced by oracle/_ref/d3/somar_ref (maxDepth and ref ratios
# printed by the reference's own coarsening strategies at verbosity 3)
SCHEDULES = [
    ((32, 32, 16), (1.0, 1.0, 1.0), (16, 16, 0), 4, 6, [(2, 2, 2)] * 4 + [(1, 1, 1)]),
    ((1024, 1024, 256), (16.0, 16.0, 1.0), (128, 128, 0), 16, 6, [(2, 2, 2)] * 7 + [(1, 1, 1)]),
    ((16, 16, 32), (1.0, 1.0, 6.0), (8, 8, 0), 4, 5, None),
]


@pytest.mark.parametrize("nx,L,mb,bf,relax,expect", SCHEDULES)
def test_schedule(nx, L, mb, bf, relax, expect):
    lo = np.array([0, 0, -nx[2]])
    hi = lo + np.array(nx) - 1
    blo, bhi = sb.make_base_grids(lo, hi, mb, (1, 1, 0), bf)
    sched = sb.plan_schedule(lo, hi, np.array(L) / np.array(nx), blo, bhi, relax_method=relax)
    assert sched[-1] == (1, 1, 1)
    if expect is not None:
        assert sched == expect
    # every box stays coarsenable along the schedule
    cur_lo, cur_hi = blo.copy(), bhi.copy()
    for r in sched[:-1]:
        r = np.array(r)
        assert np.all(cur_lo % r == 0) and np.all((cur_hi + 1) % r == 0)
        cur_lo, cur_hi = cur_lo // r, (cur_hi + 1) // r - 1


@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
def test_plan_tiles_cover_and_neighbours_are_symmetric(nranks):
    lo, hi = (0, 0, -256), (1023, 1023, -1)
    blo, bhi = sb.make_base_grids(lo, hi, (128, 128, 0), (1, 1, 0), 16)
    ranks = sb.assign_boxes_to_ranks(blo, bhi, nranks)
    plans = [sb.plan_tile(lo, hi, blo, bhi, ranks, r, nranks, periodic=(1, 0, 0)) for r in range(nranks)]
    cells = sum(int(np.prod(np.array(p[1]) - np.array(p[0]) + 1)) for p in plans)
    assert cells == 1024 * 1024 * 256
    assert sum(p[4] for p in plans) == 64
    for r, (tlo, thi, kind, nb, _) in enumerate(plans):
        assert tlo[2] == -256 and thi[2] == -1            # the vertical is never split
        assert kind[4] == 0 and kind[5] == 0
        for s in range(4):
            if kind[s] == 2:
                other = plans[nb[s]]
                assert other[2][s ^ 1] == 2 and other[3][s ^ 1] == r   # my lo neighbour sees me on its hi side
            elif kind[s] == 1:
                assert s < 2 and tlo[0] == 0 and thi[0] == 1023       # periodic in x onto itself


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = (0, 0, -16), (63, 31, -1)
    blo, bhi = sb.make_base_grids(lo, hi, (16, 16, 0), (1, 1, 0), 4)
    ranks = sb.assign_boxes_to_ranks(blo, bhi, world)
    plan = sb.plan_tile(lo, hi, blo, bhi, ranks, rank, world, periodic=(0, 1, 0))
    objs = [None] * world
    dist.all_gather_object(objs, plan)
    ok = True
    for s in range(4):
        if plan[2][s] == 2:
            o = objs[plan[3][s]]
            d = s // 2
            # the two tiles share the face: same extents in the other directions
            for e in range(3):
                if e != d:
                    ok &= o[0][e] == plan[0][e] and o[1][e] == plan[1][e]
            ok &= o[2][s ^ 1] == 2 and o[3][s ^ 1] == rank
    q.put((rank, ok, plan))
    dist.destroy_process_group()


def test_two_rank_plan_over_gloo():
    """world_size-2 gloo run of the host-side decomposition: each rank plans its own tile, the plans
    are all-gathered and must describe a consistent exchange (who sends which face to whom)."""
    import torch.multiprocessing as mp
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = 29400 + os.getpid() % 500
    procs = [ctxmp.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    tiles = {r: (tuple(pl[0]), tuple(pl[1])) for r, _, pl in res}
    assert tiles[0] != tiles[1]


@pytest.mark.parametrize("name", ["amr_r2_centre", "amr_aniso_wall", "amr_r4_periodic", "corner", "thin", "near_wall"])
def test_cf_stencil_plan(name):
    """sb_plan_cf_stencils (host-only C++, the table builder of next round's coarse-fine ghost kernel) against the
    numpy specification tests/amr_cfinterp_spec.py, which itself reproduces the reference bit for bit
    (tests/test_oracle_amr_cpu.py): same coarse cells per patch side and, on random coarse data, the same first,
    second and mixed tangential derivatives -- centred, one-sided, order-dropped and out-of-buffer cases included."""
    import somar_b200 as sb
    from amr_cases import AMR_CASES, fine_shape
    from amr_cfinterp_spec import CFInterpSpec
    extra = {
        "corner": dict(nx=(16, 16, 16), L=(2.0, 1.0, 1.0), offset=(0, 0, -16), periodic=(0, 0, 0), ref=(2, 2, 2),
                       region=(0, 0, -16, 7, 5, -9), fine_max_box=16),
        "thin": dict(nx=(16, 16, 16), L=(1.0, 1.0, 1.0), offset=(0, 0, 0), periodic=(0, 0, 0), ref=(2, 2, 2),
                     region=(4, 6, 4, 11, 7, 11), fine_max_box=8),
        "near_wall": dict(nx=(16, 16, 16), L=(1.0, 1.0, 1.0), offset=(0, 0, 0), periodic=(0, 0, 0), ref=(4, 2, 2),
                          region=(1, 1, 1, 8, 14, 10), fine_max_box=16),
    }
    c = AMR_CASES[name] if name in AMR_CASES else extra[name]
    nx, ref, reg, off = np.array(c["nx"]), np.array(c["ref"]), c["region"], np.array(c["offset"])
    nf = fine_shape(c)
    flo = np.array([reg[d] * ref[d] for d in range(3)])
    fmb = c["fine_max_box"]
    nb = [(nf[d] + fmb - 1) // fmb for d in range(3)]
    sz = np.array([nf[d] // nb[d] for d in range(3)])
    boxes = [(flo + np.array(i) * sz, flo + np.array(i) * sz + sz - 1) for i in np.ndindex(*nb)]
    dxf = np.array(c["L"]) / nx / ref
    dxc = dxf * ref
    p0 = np.random.default_rng(8).standard_normal(tuple(nx))
    phic = lambda cc: p0[tuple(np.array(cc) - off)]
    spec = CFInterpSpec(off, off + nx - 1, c["periodic"], ref, boxes, dxf)
    spec.ghosts(phic, lambda ff: 0.0)
    dom_hi = off + nx - 1
    inside = lambda cc: all(c["periodic"][d] or off[d] <= cc[d] <= dom_hi[d] for d in range(3))
    val = lambda cc: phic(cc) if inside(cc) and all(off[d] <= cc[d] <= dom_hi[d] for d in range(3)) else 0.0
    checked = 0
    for b in range(len(boxes)):
        for d in range(3):
            tr = [t for t in range(3) if t != d]
            for side in (0, 1):
                want = spec.derivs.get((b, d, 1 if side else -1), {})
                cells, w1, w2, wm = sb.plan_cf_stencils(off, dom_hi, c["periodic"], ref, [bx[0] for bx in boxes], [bx[1] for bx in boxes],
                                                        b, d, side)
                assert sorted(map(tuple, cells.tolist())) == sorted(want.keys())
                for n, cc in enumerate(map(tuple, cells.tolist())):
                    slope, curv, mixed = want[cc]
                    ca = np.array(cc)
                    for q, t in enumerate(tr):
                        e = np.eye(3, dtype=int)[t]
                        s1 = sum(w1[n, q, o + 2] * val(tuple(ca + o * e)) for o in range(-2, 3) if w1[n, q, o + 2] != 0.0) / dxc[t]
                        s2 = sum(w2[n, q, o + 2] * val(tuple(ca + o * e)) for o in range(-2, 3) if w2[n, q, o + 2] != 0.0) / (dxc[t] * dxc[t])
                        assert abs(s1 - slope[t]) <= 1e-12 * (1.0 + abs(slope[t]))
                        assert abs(s2 - curv[t]) <= 1e-12 * (1.0 + abs(curv[t]))
                    e0, e1 = np.eye(3, dtype=int)[tr[0]], np.eye(3, dtype=int)[tr[1]]
                    sm = sum(wm[n, o1 + 1, o0 + 1] * val(tuple(ca + o0 * e0 + o1 * e1)) for o1 in range(-1, 2) for o0 in range(-1, 2)
                             if wm[n, o1 + 1, o0 + 1] != 0.0) / (dxc[tr[0]] * dxc[tr[1]])
                    assert abs(sm - mixed) <= 1e-12 * (1.0 + abs(mixed))
                    checked += 1
    assert checked > 0


@pytest.mark.parametrize("nx,ny", [(2, 2), (64, 2), (65, 3), (128, 7), (512, 256), (1000, 33), (3, 40)])
def test_fused_halo_tile_schedule_covers_the_tile_once(nx, ny):
    """vertline_tma_k with the fused neighbour exchange walks the tiles of an exchanged side first (sb_line_tma.cu: tile_of).
    Whatever the shape and the set of exchanged sides, every tile is visited exactly once, the tiles touching a side come
    before every interior tile, and each side is touched by as many tiles as the kernel's arrival counters expect
    (ny along an x side, nbx along a y side)."""
    import ctypes as C
    from somar_b200 import capi
    lib = capi.load()
    nbx = ((nx + 1) // 2 + 31) // 32
    for mask in range(16):
        order = (C.c_int * (3 * nbx * ny))()
        n = C.c_int()
        capi.check(lib.sb_plan_line_tile_order(nx, ny, mask, order, nbx * ny, C.byref(n)))
        assert n.value == nbx * ny
        tiles = [(order[3 * u], order[3 * u + 1], order[3 * u + 2]) for u in range(n.value)]
        assert sorted((bx, j) for bx, j, _ in tiles) == [(bx, j) for bx in range(nbx) for j in range(ny)]
        per_side = [0, 0, 0, 0]
        seen_interior = False
        for bx, j, touch in tiles:
            want = ((mask & 1) and bx == 0) * 1 | ((mask & 2) and bx == nbx - 1) * 2 | ((mask & 4) and j == 0) * 4 | ((mask & 8) and j == ny - 1) * 8
            assert touch == want
            if touch:
                assert not seen_interior
            else:
                seen_interior = True
            for s in range(4):
                per_side[s] += (touch >> s) & 1
        assert per_side == [ny if mask & 1 else 0, ny if mask & 2 else 0, nbx if mask & 4 else 0, nbx if mask & 8 else 0]
