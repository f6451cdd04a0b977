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


def fine_shape(c):
    return tuple((c["region"][3 + d] - c["region"][d] + 1) * c["ref"][d] for d in range(3))


def region_slices(c):
    return tuple(slice(c["region"][d] - c["offset"][d], c["region"][3 + d] - c["offset"][d] + 1) for d in range(3))


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
    r0[sl] = r1.reshape(nf[0] // ref[0], ref[0], nf[1] // ref[1], ref[1], nf[2] // ref[2], ref[2]).mean(axis=(1, 3, 5))
    mask = np.ones(nx, bool)
    mask[sl] = False
    total = r0[mask].sum() + r1.sum() / np.prod(ref)   # in units of the coarse cell volume
    r0[mask] -= total / mask.sum()
    return np.asfortranarray(r0), np.asfortranarray(r1)


def ref_kwargs_amr(c, **extra):
    kw = dict(nx=c["nx"], L=c["L"], max_box=c["max_box"], block_factor=c["bf"], offset=c["offset"], periodic=c["periodic"],
              relax=c["relax"], split_dirs=(1, 1, 1),
              extra={"drv.refRatio": " ".join(map(str, c["ref"])), "drv.fineRegion": " ".join(map(str, c["region"])),
                     "drv.fineMaxBox": c["fine_max_box"]})
    kw["extra"].update(extra)
    return kw


def composite_integral(c, f0, f1):
    """Integral over the composite grid in units of the coarse cell volume."""
    mask = np.ones(c["nx"], bool)
    mask[region_slices(c)] = False
    return f0.reshape(c["nx"], order="F")[mask].sum() + f1.sum() / np.prod(c["ref"])
