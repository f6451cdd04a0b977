"""Two-level (base level + one refined patch) problems for the AMR oracle:
shared by the CPU tests of the oracle, the fixture generator and -- next round -- the GPU parity tests."""
import numpy as np

# name -> dict(nx, L, offset, max_box, bf, periodic, relax, ref, region (coarse-index box lo..hi), fine_max_box)
AMR_CASES = {
    # a centred patch, refinement 2, Neumann walls
    "amr_r2_centre": dict(nx=(32, 32, 16), L=(1.0, 1.0, 1.0), offset=(0, 0, 0), max_box=(16, 16, 0), bf=4, periodic=(0, 0, 0), relax=5,
                          ref=(2, 2, 2), region=(8, 8, 4, 23, 23, 11), fine_max_box=16),
    # the BuoyantVortexRing deck's shape in miniature: refinement 4, triply periodic base, patch off centre, line relaxation off
    "amr_r4_periodic": dict(nx=(16, 16, 16), L=(10.0, 10.0, 10.0), offset=(-8, -8, -8), max_box=(8, 8, 8), bf=8, periodic=(1, 1, 1),
                            relax=5, ref=(4, 4, 4), region=(-4, -2, -4, 3, 5, 3), fine_max_box=16),
    # anisotropic grid and refinement, patch touching a wall
    "amr_aniso_wall": dict(nx=(32, 16, 16), L=(4.0, 2.0, 1.0), offset=(0, 0, -16), max_box=(16, 16, 0), bf=4, periodic=(0, 0, 0),
                           relax=5, ref=(2, 2, 1), region=(0, 4, -12, 15, 11, -5), fine_max_box=16),
}


DJL_L = 724.07734393502466498646462679537
AMR_CASES.update({
    # 2-D build (CH_SPACEDIM = 2; directions x, z).  The DJL deck's base level with a vertically spanning patch at
    # the deck's two refinement ratios: the level solves inside AMRHybridSolver run in the leptic modes.
    "amr2d_djl_r22": dict(nx=(128, 32), L=(DJL_L, 1.0), offset=(0, -32), max_box=(32, 0), bf=16, periodic=(0, 0), relax=6,
                          ref=(2, 2), region=(32, -32, 95, -1), fine_max_box=128, split=(1, 0)),
    "amr2d_djl_r41": dict(nx=(128, 32), L=(DJL_L, 1.0), offset=(0, -32), max_box=(32, 0), bf=16, periodic=(0, 0), relax=6,
                          ref=(4, 1), region=(32, -32, 95, -1), fine_max_box=128, split=(1, 0)),
    "amr2d_r2_centre": dict(nx=(64, 32), L=(4.0, 1.0), offset=(0, -32), max_box=(32, 0), bf=8, periodic=(0, 0), relax=5,
                            ref=(2, 2), region=(16, -24, 47, -9), fine_max_box=32, split=(1, 1)),
})


def ndim(c):
    return len(c["nx"])


def fine_shape(c):
    D = ndim(c)
    return tuple((c["region"][D + d] - c["region"][d] + 1) * c["ref"][d] for d in range(D))


def region_slices(c):
    D = ndim(c)
    return tuple(slice(c["region"][d] - c["offset"][d], c["region"][D + d] - c["offset"][d] + 1) for d in range(D))


def composite_rhs(c, seed):
    """Random right-hand sides on both levels that satisfy the solvability condition of the all-Neumann /
    periodic composite problem: the coarse values under the patch are the block averages of the fine ones
    and the integral over the composite grid (uncovered coarse cells + fine cells) vanishes."""
    rng = np.random.default_rng(seed)
    nx, ref = c["nx"], c["ref"]
    r0 = rng.standard_normal(nx)
    nf = fine_shape(c)
    r1 = rng.standard_normal(nf)
    sl = region_slices(c)
    shape = [v for d in range(ndim(c)) for v in (nf[d] // ref[d], ref[d])]
    r0[sl] = r1.reshape(shape).mean(axis=tuple(range(1, 2 * ndim(c), 2)))
    mask = np.ones(nx, bool)
    mask[sl] = False
    total = r0[mask].sum() + r1.sum() / np.prod(ref)   # in units of the coarse cell volume
    r0[mask] -= total / mask.sum()
    return np.asfortranarray(r0), np.asfortranarray(r1)


def ref_kwargs_amr(c, **extra):
    kw = dict(nx=c["nx"], L=c["L"], max_box=c["max_box"], block_factor=c["bf"], offset=c["offset"], periodic=c["periodic"],
              relax=c["relax"], split_dirs=c.get("split", (1, 1, 1)), dim=ndim(c),
              extra={"drv.refRatio": " ".join(map(str, c["ref"])), "drv.fineRegion": " ".join(map(str, c["region"])),
                     "drv.fineMaxBox": c["fine_max_box"]})
    kw["extra"].update(extra)
    return kw


def composite_integral(c, f0, f1):
    """Integral over the composite grid in units of the coarse cell volume."""
    mask = np.ones(c["nx"], bool)
    mask[region_slices(c)] = False
    return f0.reshape(c["nx"], order="F")[mask].sum() + f1.sum() / np.prod(c["ref"])


# Three levels: the BuoyantVortexRing deck's hierarchy (64^3 base, two refined levels x (4, 4, 4), triply periodic,
# boxes of 16^3, GSRB) in miniature and at the deck's size, with rectangular patches standing in for the deck's
# vorticity-tagged grids.  region2 is in level-1 indices.
AMR_CASES.update({
    "amr3_r2_r2": dict(nx=(16, 16, 16), L=(1.0, 1.0, 1.0), offset=(0, 0, 0), max_box=(8, 8, 0), bf=4, periodic=(0, 0, 0), relax=5,
                       ref=(2, 2, 2), region=(4, 4, 4, 11, 11, 11), fine_max_box=8,
                       ref2=(2, 2, 2), region2=(12, 12, 12, 19, 19, 19), fine_max_box2=8),
    "amr3_c3_mini": dict(nx=(16, 16, 16), L=(10.0, 10.0, 10.0), offset=(-8, -8, -8), max_box=(8, 8, 8), bf=8, periodic=(1, 1, 1), relax=5,
                         ref=(4, 4, 4), region=(-4, -2, -4, 3, 5, 3), fine_max_box=16,
                         ref2=(4, 4, 4), region2=(-8, 0, -8, -1, 7, -1), fine_max_box2=16),
})
C3_DECK = dict(nx=(64, 64, 64), L=(10.0, 10.0, 10.0), offset=(-32, -32, -32), max_box=(16, 16, 16), bf=16, periodic=(1, 1, 1), relax=5,
               ref=(4, 4, 4), region=(-8, -8, -4, 7, 7, 11), fine_max_box=16,
               ref2=(4, 4, 4), region2=(-16, -16, 0, 15, 15, 31), fine_max_box2=16)


def num_levels(c):
    return 3 if "region2" in c else 2


def level_specs(c):
    """Per level: dict(dom_lo, dom_hi, dXi, box_lo[n,3], box_hi[n,3], reg_lo, reg_hi, ref) in the library's 3-slot convention
    (2-D problems: directions x, z in slots 0 and 2).  Boxes are cut the way oracle/ref_driver.cpp cuts them."""
    import somar_b200 as sb
    D = ndim(c)
    to3 = (lambda v, fill: (v[0], fill, v[1])) if D == 2 else (lambda v, fill: tuple(v))
    nx = np.array(to3(c["nx"], 1))
    L = np.array(to3(c["L"], 1.0), dtype=float)
    off = np.array(to3(c["offset"], 0))
    split = to3(c.get("split", (1,) * D), 0)
    dXi = L / nx
    lo, hi = off, off + nx - 1
    blo, bhi = sb.make_base_grids(lo, hi, to3(c["max_box"], 0), split, c["bf"])
    out = [dict(dom_lo=lo, dom_hi=hi, dXi=dXi, box_lo=blo, box_hi=bhi, reg_lo=lo, reg_hi=hi, ref=np.array([1, 1, 1]))]
    for l in range(1, num_levels(c)):
        sfx = "" if l == 1 else "2"
        ref = np.array(to3(c["ref" + sfx], 1))
        reg = c["region" + sfx]
        rlo, rhi = np.array(to3(reg[:D], 0)), np.array(to3(reg[D:], 0))
        prev = out[-1]
        dom_lo, dom_hi = prev["dom_lo"] * ref, (prev["dom_hi"] + 1) * ref - 1
        flo, fhi = rlo * ref, (rhi + 1) * ref - 1
        n = fhi - flo + 1
        fmb = c["fine_max_box" + sfx]
        nb = np.array([(n[d] + fmb - 1) // fmb if fmb > 0 else 1 for d in range(3)])
        if D == 2:
            nb[1] = 1
        sz = n // nb
        los, his = [], []
        for k in range(nb[2]):
            for j in range(nb[1]):
                for i in range(nb[0]):
                    b = flo + np.array([i, j, k]) * sz
                    los.append(b)
                    his.append(b + sz - 1)
        out.append(dict(dom_lo=dom_lo, dom_hi=dom_hi, dXi=prev["dXi"] / ref, box_lo=np.array(los, dtype=np.int32),
                        box_hi=np.array(his, dtype=np.int32), reg_lo=flo, reg_hi=fhi, ref=ref))
    return out


def level_shapes(c):
    D = ndim(c)
    pick = (lambda v: (v[0], v[2])) if D == 2 else (lambda v: tuple(v))
    return [pick(tuple(int(x) for x in (s["reg_hi"] - s["reg_lo"] + 1))) for s in level_specs(c)]


def make_amr_ops(ctx, c, ranks=None):
    """One PoissonOp per AMR level (Cartesian map, HomogNeumBC), each knowing its coarser level's grids."""
    import somar_b200 as sb
    D = ndim(c)
    per = (c["periodic"][0], 0, c["periodic"][1]) if D == 2 else tuple(c["periodic"])
    ops = []
    for l, s in enumerate(level_specs(c)):
        ops.append(sb.PoissonOp(ctx, s["dom_lo"], s["dom_hi"], s["dXi"], s["box_lo"], s["box_hi"], box_rank=None if ranks is None else ranks[l],
                                periodic=per, dim=D, relax_method=c["relax"], crse_op=ops[-1] if l else None))
    return ops


def ref_kwargs_amr3(c, **extra):
    kw = ref_kwargs_amr(c, **extra)
    if "region2" in c:
        kw["extra"].update({"drv.refRatio2": " ".join(map(str, c["ref2"])), "drv.fineRegion2": " ".join(map(str, c["region2"])),
                            "drv.fineMaxBox2": c["fine_max_box2"]})
    return kw


def composite_rhs_levels(c, seed):
    """Random right-hand sides on every level, consistent under block averaging and with zero composite integral."""
    rng = np.random.default_rng(seed)
    specs = level_specs(c)
    D = ndim(c)
    shapes = level_shapes(c)
    r = [rng.standard_normal(sh) for sh in shapes]
    pick = (lambda v: np.array([v[0], v[2]])) if D == 2 else (lambda v: np.array(v))
    # coarse values under a finer level = block averages, finest level first
    masks = [np.ones(sh, bool) for sh in shapes]
    for l in range(len(specs) - 1, 0, -1):
        ref = pick(specs[l]["ref"])
        lo_c = pick(specs[l]["reg_lo"]) // ref - pick(specs[l - 1]["reg_lo"])
        n_c = np.array(shapes[l]) // ref
        sl = tuple(slice(lo_c[d], lo_c[d] + n_c[d]) for d in range(D))
        shp = [v for d in range(D) for v in (n_c[d], ref[d])]
        r[l - 1][sl] = r[l].reshape(shp).mean(axis=tuple(range(1, 2 * D, 2)))
        masks[l - 1][sl] = False
    # composite integral in units of the level-0 cell volume
    vol = [1.0]
    for l in range(1, len(specs)):
        vol.append(vol[-1] / np.prod(pick(specs[l]["ref"])))
    total = sum(r[l][masks[l]].sum() * vol[l] for l in range(len(specs)))
    r[0][masks[0]] -= total / masks[0].sum()
    for l in range(len(specs) - 1, 0, -1):   # keep the covered coarse values the averages they were
        pass
    return [np.asfortranarray(a) for a in r], masks


# Patches that stress the one-sided / order-dropping branches of the coarse derivative stencils
# (used for the interpolation spec only, not solved).
SPEC_CASES = dict({n: c for n, c in AMR_CASES.items() if len(c["nx"]) == 3 and "region2" not in c}, **{
    "corner": dict(nx=(16, 16, 16), L=(2.0, 1.0, 1.0), offset=(0, 0, -16), max_box=(8, 8, 0), bf=4, periodic=(0, 0, 0), relax=5,
                   ref=(2, 2, 2), region=(0, 0, -16, 7, 5, -9), fine_max_box=16),
    "thin": dict(nx=(16, 16, 16), L=(1.0, 1.0, 1.0), offset=(0, 0, 0), max_box=(8, 8, 0), bf=4, periodic=(0, 0, 0), relax=5,
                 ref=(2, 2, 2), region=(4, 6, 4, 11, 7, 11), fine_max_box=8),
    "near_wall": dict(nx=(16, 16, 16), L=(1.0, 1.0, 1.0), offset=(0, 0, 0), max_box=(8, 8, 0), bf=4, periodic=(0, 0, 0), relax=5,
                      ref=(4, 2, 2), region=(1, 1, 1, 8, 14, 10), fine_max_box=16),
})




def rand_velocity_levels(c, seed):
    """Random advecting face velocity on every level: D face arrays over the level's region, zero normal component on
    non-periodic domain walls, equal values on periodic images (so the composite divergence integrates to zero)."""
    rng = np.random.default_rng(seed)
    D = ndim(c)
    pick = (lambda v: np.array([v[0], v[2]])) if D == 2 else (lambda v: np.array(v))
    out = []
    for s, sh in zip(level_specs(c), level_shapes(c)):
        lo, hi = pick(s["reg_lo"]), pick(s["reg_hi"])
        dlo, dhi = pick(s["dom_lo"]), pick(s["dom_hi"])
        lev = []
        for d in range(D):
            shape = list(sh)
            shape[d] += 1
            u = rng.standard_normal(tuple(shape))
            first, last = [slice(None)] * D, [slice(None)] * D
            first[d], last[d] = 0, -1
            at_lo, at_hi = lo[d] == dlo[d], hi[d] == dhi[d]
            if c["periodic"][d]:
                if at_lo and at_hi:
                    u[tuple(last)] = u[tuple(first)]
            else:
                if at_lo:
                    u[tuple(first)] = 0.0
                if at_hi:
                    u[tuple(last)] = 0.0
            lev.append(np.asfortranarray(u))
        out.append(lev)
    return out
