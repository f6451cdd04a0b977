"""Shared problem definitions for the parity tests (synthetic family S* of SURVEY.md 8d, small)."""
import numpy as np

import somar_b200 as sb

# name -> dict(nx, L, max_box, bf, periodic, relax, map, ampl)
CASES = {
    "line_cart": dict(nx=(32, 32, 16), L=(1.0, 1.0, 1.0), max_box=(16, 16, 0), bf=4, periodic=(0, 0, 0), relax=6, map="cartesian", ampl=(0, 0, 0)),
    "line_stretch": dict(nx=(32, 32, 16), L=(1.0, 1.0, 1.0), max_box=(16, 16, 0), bf=4, periodic=(0, 0, 0), relax=6, map="stretched", ampl=(0.05, 0.03, -0.1)),
    "line_aniso": dict(nx=(64, 32, 32), L=(8.0, 4.0, 1.0), max_box=(16, 16, 0), bf=8, periodic=(0, 0, 0), relax=6, map="stretched", ampl=(0.0, 0.0, -0.1)),
    "line_perx": dict(nx=(32, 16, 16), L=(2.0, 1.0, 0.5), max_box=(16, 16, 0), bf=4, periodic=(1, 0, 0), relax=6, map="cartesian", ampl=(0, 0, 0)),
    "gsrb_cart": dict(nx=(32, 32, 32), L=(1.0, 1.0, 8.0), max_box=(16, 16, 0), bf=4, periodic=(0, 0, 0), relax=5, map="cartesian", ampl=(0, 0, 0)),
    "gsrb_stretch": dict(nx=(32, 16, 32), L=(1.0, 1.0, 6.0), max_box=(16, 16, 0), bf=4, periodic=(0, 0, 0), relax=5, map="stretched", ampl=(0.04, 0.0, 0.3)),
    "gsrb_perxy": dict(nx=(16, 16, 32), L=(1.0, 1.0, 6.0), max_box=(8, 8, 0), bf=4, periodic=(1, 1, 0), relax=5, map="cartesian", ampl=(0, 0, 0)),
    # the S-family of SURVEY.md 8d at depths that are multiples of 64, so that the production instance of the line kernel
    # (vertline_fused_k<8,4,2,ALIGNED=true>, the one bench.py times) meets the oracle directly
    "s_line64": dict(nx=(64, 64, 64), L=(1.0, 1.0, 1.0), max_box=(32, 32, 0), bf=16, periodic=(0, 0, 0), relax=6, map="cartesian", ampl=(0, 0, 0)),
    "s_line64_zstretch": dict(nx=(64, 32, 64), L=(1.0, 0.5, 1.0), max_box=(32, 32, 0), bf=16, periodic=(0, 0, 0), relax=6, map="stretched", ampl=(0, 0, -0.1)),
    "s_line128": dict(nx=(128, 128, 128), L=(2.0, 2.0, 1.0), max_box=(64, 64, 0), bf=16, periodic=(0, 0, 0), relax=6, map="cartesian", ampl=(0, 0, 0)),
    "s_line256": dict(nx=(128, 128, 256), L=(2.0, 2.0, 1.0), max_box=(0, 0, 0), bf=16, periodic=(0, 0, 0), relax=6, map="cartesian", ampl=(0, 0, 0)),
    # horizontally stretched maps at depths the mapped-grid line kernel (vertline_tma_k<GENERAL>) takes: nz a multiple of 32
    "s_line64_xystretch": dict(nx=(64, 32, 64), L=(1.0, 0.5, 1.0), max_box=(32, 32, 0), bf=16, periodic=(0, 0, 0), relax=6, map="stretched", ampl=(0.05, 0.03, -0.1)),
    "s_line256_xystretch": dict(nx=(64, 64, 256), L=(1.0, 1.0, 1.0), max_box=(32, 32, 0), bf=16, periodic=(0, 0, 0), relax=6, map="stretched", ampl=(0.05, 0.03, -0.1)),
    "onebox": dict(nx=(16, 16, 8), L=(1.0, 1.0, 1.0), max_box=(0, 0, 0), bf=4, periodic=(0, 0, 0), relax=6, map="cartesian", ampl=(0, 0, 0)),
}


def geometry(c):
    nx = np.array(c["nx"])
    L = np.array(c["L"], dtype=float)
    dXi = L / nx
    lo = np.array([0, 0, -nx[2]])
    hi = lo + nx - 1
    return nx, L, dXi, lo, hi


def make_op(ctx, c, box_rank=None):
    nx, L, dXi, lo, hi = geometry(c)
    blo, bhi = sb.make_base_grids(lo, hi, c["max_box"], (1, 1, 0), c["bf"])
    xmin = lo * dXi
    kind = sb.MAP_CARTESIAN if c["map"] == "cartesian" else sb.MAP_STRETCHED
    return sb.PoissonOp(ctx, lo, hi, dXi, blo, bhi, box_rank=box_rank, periodic=c["periodic"], map_kind=kind, map_xmin=xmin,
                        map_xmax=xmin + L, map_ampl=c["ampl"], relax_method=c["relax"])


def ref_kwargs(c):
    nx, L, dXi, lo, hi = geometry(c)
    return dict(nx=c["nx"], L=c["L"], max_box=c["max_box"], block_factor=c["bf"], offset=lo, periodic=c["periodic"],
                relax=c["relax"], mapname=c["map"], ampl=c["ampl"])


def rand_field(c, seed, zero_mean=False):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(c["nx"])
    if zero_mean:
        a -= a.mean()
    return np.asfortranarray(a)


def rand_velocity(c, seed):
    """Random face velocity with zero normal component on non-periodic walls and equal values on
    periodic images (so that Sum div = 0, SURVEY.md 8d RHS-B in miniature)."""
    rng = np.random.default_rng(seed)
    nx = np.array(c["nx"])
    out = []
    for d in range(3):
        shape = nx.copy()
        shape[d] += 1
        u = rng.standard_normal(tuple(shape))
        sl_lo = [slice(None)] * 3
        sl_hi = [slice(None)] * 3
        sl_lo[d], sl_hi[d] = 0, -1
        if c["periodic"][d]:
            u[tuple(sl_hi)] = u[tuple(sl_lo)]
        else:
            u[tuple(sl_lo)] = 0.0
            u[tuple(sl_hi)] = 0.0
        out.append(np.asfortranarray(u))
    return out


def rel_err(a, b):
    a, b = np.asarray(a).ravel(order="F"), np.asarray(b).ravel(order="F")
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))
