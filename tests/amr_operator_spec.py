"""TEST INFRASTRUCTURE (numpy): an executable specification of the reference's two-level composite operator
on a Cartesian grid (alpha = 0, beta = 1) -- the level operator with quadratically interpolated
coarse-fine ghosts on the fine level and the REFLUXED operator on the coarse level
(PoissonOp::AMROperatorNF / AMROperatorNC / reflux, PoissonOp.cpp:1156-1200, 1296-1440, with
AnisotropicFluxRegister::incrementCoarse / incrementFine / reflux, AnisotropicFluxRegister.cpp:329-384,
411-470, 600-620).  Checked against the oracle in tests/test_oracle_amr_cpu.py; blueprint for next round's
CUDA path."""
import numpy as np

from amr_cfinterp_spec import CFInterpSpec


def _level_op(phi_g, dx):
    """J (sum_d M_d (phi(i-1) + phi(i+1))) + phi / Dinv on the interior of a ghosted array, Cartesian: M = 1/dx^2."""
    m = [1.0 / (dx[d] * dx[d]) for d in range(3)]
    c = phi_g[1:-1, 1:-1, 1:-1]
    s = m[0] * phi_g[:-2, 1:-1, 1:-1] + m[0] * phi_g[2:, 1:-1, 1:-1] + m[1] * phi_g[1:-1, :-2, 1:-1] + m[1] * phi_g[1:-1, 2:, 1:-1]
    s = s + m[2] * phi_g[1:-1, 1:-1, :-2] + m[2] * phi_g[1:-1, 1:-1, 2:]
    dinv = 1.0 / (1.0 * (0.0 - 1.0 * (m[0] + m[0] + (m[1] + m[1] + m[2] + m[2]))))
    return 1.0 * 1.0 * s + c / dinv


def composite_minus_L(c, p0, p1, fine_boxes):
    """Returns (-L[phi] on level 0 over the domain, -L[phi] on level 1 over the patch).  Values of the level-0
    result under the patch are the plain level operator (the reference leaves them; norms mask them out)."""
    nx, ref, reg, off = np.array(c["nx"]), np.array(c["ref"]), c["region"], np.array(c["offset"])
    per = c["periodic"]
    dxc = np.array(c["L"]) / nx
    dxf = dxc / ref
    nf = np.array(p1.shape)
    flo = np.array([reg[d] * ref[d] for d in range(3)])
    rlo, rhi = np.array(reg[:3]), np.array(reg[3:])

    # ---- level 1: ghosted patch array; walls Neumann (ghost = first interior), coarse-fine faces interpolated
    g1 = np.zeros(nf + 2)
    g1[1:-1, 1:-1, 1:-1] = p1
    spec = CFInterpSpec(off, off + nx - 1, per, ref, fine_boxes, dxf)
    cf = spec.ghosts(lambda cc: p0[tuple(np.array(cc) - off)], lambda ff: p1[tuple(np.array(ff) - flo)])
    for f, v in cf.items():
        g1[tuple(np.array(f) - flo + 1)] = v
    for d in range(3):
        if not per[d]:
            if rlo[d] == off[d]:
                sl_g, sl_i = [slice(1, -1)] * 3, [slice(1, -1)] * 3
                sl_g[d], sl_i[d] = 0, 1
                g1[tuple(sl_g)] = g1[tuple(sl_i)]
            if rhi[d] == off[d] + nx[d] - 1:
                sl_g, sl_i = [slice(1, -1)] * 3, [slice(1, -1)] * 3
                sl_g[d], sl_i[d] = -1, -2
                g1[tuple(sl_g)] = g1[tuple(sl_i)]
    L1 = _level_op(g1, dxf)

    # ---- level 0: periodic wrap or Neumann walls, level operator everywhere, then reflux next to the patch
    g0 = np.pad(p0, 1, mode="wrap")
    for d in range(3):
        if not per[d]:
            sl_g, sl_i = [slice(None)] * 3, [slice(None)] * 3
            sl_g[d], sl_i[d] = 0, 1
            g0[tuple(sl_g)] = g0[tuple(sl_i)]
            sl_g[d], sl_i[d] = -1, -2
            g0[tuple(sl_g)] = g0[tuple(sl_i)]
    L0 = _level_op(g0, dxc)
    coarse_reg, fine_reg = np.zeros(nx), np.zeros(nx)
    for d in range(3):
        denom = float(np.prod(ref) // ref[d])
        s = 1.0 / dxc[d]
        tr = [t for t in range(3) if t != d]
        for sgn in (-1, 1):
            cd = rlo[d] - 1 if sgn < 0 else rhi[d] + 1         # coarse cells just outside the patch on this side
            if not per[d] and not (off[d] <= cd <= off[d] + nx[d] - 1):
                continue                                        # physical wall: no coarse-fine interface
            for ct0 in range(rlo[tr[0]], rhi[tr[0]] + 1):
                for ct1 in range(rlo[tr[1]], rhi[tr[1]] + 1):
                    cc = np.zeros(3, dtype=int)
                    cc[d], cc[tr[0]], cc[tr[1]] = cd, ct0, ct1
                    inner = cc.copy()
                    inner[d] -= sgn                             # the coarse cell under the patch across the face
                    a, b = (cc, inner) if sgn < 0 else (inner, cc)          # face between a (low) and b (high)
                    Fc = (p0[tuple(b - off)] - p0[tuple(a - off)]) / dxc[d] * 1.0
                    coarse_reg[tuple(cc - off)] += Fc * (-sgn * s)
                    # the fine faces of this coarse face, x fastest (the Fortran loop of ANISOTROPICINCREMENTFINE)
                    fl = inner * ref
                    acc = 0.0
                    rng_ = [range(fl[e], fl[e] + ref[e]) for e in range(3)]
                    rng_[d] = [fl[d] + (ref[d] - 1 if sgn > 0 else 0)]      # the fine layer that touches the face
                    for k in rng_[2]:
                        for j in rng_[1]:
                            for i in rng_[0]:
                                fi = np.array([i, j, k])
                                fo = fi.copy()
                                fo[d] += sgn                     # the fine ghost cell across the face
                                lo_, hi_ = (fo, fi) if sgn < 0 else (fi, fo)
                                Ff = (g1[tuple(hi_ - flo + 1)] - g1[tuple(lo_ - flo + 1)]) / dxf[d] * 1.0
                                acc = acc + (sgn * s / denom) * Ff
                    fine_reg[tuple(cc - off)] += acc
    L0 = (L0 + coarse_reg * (-1.0)) + fine_reg * (-1.0)
    return -L0, -L1
