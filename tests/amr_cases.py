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
