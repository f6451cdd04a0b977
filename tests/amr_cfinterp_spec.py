"""TEST INFRASTRUCTURE (numpy): an executable specification of the reference's quadratic coarse-fine
ghost interpolation in 3-D, checked against the oracle (tests/test_oracle_amr_cpu.py).  It is the
blueprint for the CUDA kernel of SURVEY.md 8 rows a15 / f2 (next round): the host side precomputes, per
coarse cell next to a coarse-fine face, which derivative stencils apply; the device side evaluates
phistar and the normal quadratic.

Follows, for one refined patch made of fine boxes (reference paths relative to src/):
  Grade2_AnisotropicChombo/QuadCFInterp/MappedQuadCFInterp.cpp:68-217   (define: per fine box and side)
  .../MappedCFStencil.cpp:833-995   (which coarse cells are usable / 'standard')
  .../MappedCFStencil.cpp:997-1237  (one-sided stencils, order dropping; the quadrant boxes of the mixed
                                      stencil are taken as the code builds them, sign pattern included)
  .../MappedCFStencil.cpp:379-598   (derivative evaluation)
  .../MappedQuadCFInterp.cpp:271-505 + MappedQuadCFInterpF.ChF:50-127 (phistar)
  .../MappedQuadCFInterpF.ChF:9-46  (the normal quadratic)
Restrictions (asserted): 3-D, patch not touching a periodic boundary, base level covers the domain."""
import itertools

import numpy as np


def _cells(lo, hi):
    return itertools.product(*[range(lo[d], hi[d] + 1) for d in range(3)])


class CFInterpSpec:
    def __init__(self, dom_lo, dom_hi, periodic, ref, fine_boxes, dxf):
        """dom_*: coarse domain box; fine_boxes: list of (lo, hi) in fine indices; dxf: fine dXi (3)."""
        self.dlo, self.dhi = np.array(dom_lo), np.array(dom_hi)
        self.periodic, self.ref = tuple(periodic), np.array(ref)
        self.fb = [(np.array(l), np.array(h)) for l, h in fine_boxes]
        self.dxf = np.array(dxf, dtype=float)
        self.dxc = self.dxf * self.ref
        self.cb = [(l // self.ref, h // self.ref) for l, h in self.fb]     # coarsened fine boxes
        for l, h in self.cb:
            for d in range(3):
                if periodic[d]:
                    assert l[d] - 2 >= self.dlo[d] and h[d] + 2 <= self.dhi[d], "patch too close to a periodic boundary"

    def _in_domain(self, c):
        return all(self.periodic[d] or self.dlo[d] <= c[d] <= self.dhi[d] for d in range(3))

    def _covered(self, c):
        return any(all(l[d] <= c[d] <= h[d] for d in range(3)) for l, h in self.cb)

    def _fine_covered(self, f):
        return any(all(l[d] <= f[d] <= h[d] for d in range(3)) for l, h in self.fb)

    def ghosts(self, phic, phif):
        """phic(c) -> coarse value, phif(f) -> fine value (callables on index tuples).  Returns
        {fine ghost index: value} for every coarse-fine ghost cell of the patch."""
        out = {}
        self.derivs = {}     # (box, dir, side) -> {coarse cell: (slope{t}, curv{t}, mixed)}
        for b in range(len(self.fb)):
            for d in range(3):
                for side in (-1, 1):
                    self._one_side(b, d, side, phic, phif, out)
        return out

    def _one_side(self, b, d, side, phic, phif, out):
        flo, fhi = self.fb[b]
        clo, chi = self.cb[b]
        tr = [t for t in range(3) if t != d]                  # vinttran: tangential directions, ascending
        # fine ghost cells of this side that are coarse-fine ghosts
        glo, ghi = flo.copy(), fhi.copy()
        glo[d] = ghi[d] = (fhi[d] + 1) if side > 0 else (flo[d] - 1)
        fine_ivs = [f for f in _cells(glo, ghi) if self._in_domain(np.array(f) // self.ref) and not self._fine_covered(f)]
        if not fine_ivs:
            return
        base = sorted({tuple(np.array(f) // self.ref) for f in fine_ivs})
        # the coarse slab next to the face, grown by 2 (allGood) / 1 (standard) in the tangential directions
        slo, shi = clo.copy(), chi.copy()
        slo[d] = shi[d] = (chi[d] + 1) if side > 0 else (clo[d] - 1)
        g2lo, g2hi, g1lo, g1hi = slo.copy(), shi.copy(), slo.copy(), shi.copy()
        for t in tr:
            g2lo[t] -= 2; g2hi[t] += 2; g1lo[t] -= 1; g1hi[t] += 1
        good = {c for c in _cells(g2lo, g2hi) if self._in_domain(c) and not self._covered(c)}
        in_g1 = lambda c: all(g1lo[e] <= c[e] <= g1hi[e] for e in range(3))
        std = {c for c in good if in_g1(c)}
        for t in tr:                                           # IntVectSet::grow(t, -1): erosion
            e = np.eye(3, dtype=int)[t]
            std = {c for c in std if tuple(np.array(c) - e) in std and tuple(np.array(c) + e) in std}
        buf_lo, buf_hi = clo - 2, chi + 2                      # a_phic.box(): coarsened fine box grown by 2
        in_buf = lambda c: all(buf_lo[e] <= c[e] <= buf_hi[e] for e in range(3))

        def sten_sum(st):                                      # DerivStencil evaluation with the keepzer rule
            acc, zero = 0.0, False
            for idx, w in st:
                if in_buf(idx):
                    acc += w * phic(idx)
                else:
                    zero = True
            return 0.0 if zero else acc

        box_good = lambda lo, hi: all(c in good for c in _cells(lo, hi))
        deriv = {}
        for c in base:
            ca = np.array(c)
            slope, curv, mixed = {}, {}, 0.0
            if c in std:
                for t in tr:
                    e = np.eye(3, dtype=int)[t]
                    hi_, lo_, cc = phic(tuple(ca + e)), phic(tuple(ca - e)), phic(c)
                    slope[t] = (hi_ - lo_) / (2.0 * self.dxc[t])
                    curv[t] = (hi_ + lo_ - 2.0 * cc) / (self.dxc[t] * self.dxc[t])
                # computeMixedDerivative: basex = the lower tangential direction, basey the higher
                ex, ey = np.eye(3, dtype=int)[tr[0]], np.eye(3, dtype=int)[tr[1]]
                mixed = (phic(tuple(ca + ex + ey)) + phic(tuple(ca - ex - ey)) - phic(tuple(ca + ex - ey)) - phic(tuple(ca - ex + ey))) / (
                    4.0 * self.dxc[tr[1]] * self.dxc[tr[0]])
            else:
                e1, e2 = np.eye(3, dtype=int)[tr[0]], np.eye(3, dtype=int)[tr[1]]      # itran1, itran2 of buildStencils
                # quadrant boxes in the order the code tests them (ur, ul, lr, ll) and as it builds them
                quads = [(ca - e1, ca + e2), (ca, ca + e1 + e2), (ca - e2, ca + e1), (ca - e1 - e2, ca)]
                sten, nq = [], 0
                for qlo, qhi in quads:
                    if box_good(qlo, qhi):
                        nq += 1
                        for idx in _cells(qlo, qhi):           # BoxIterator order: x fastest
                            ia = np.array(idx)
                            w = -1.0 if (np.array_equal(ia, qlo) or np.array_equal(ia, qhi)) else 1.0
                            for n_, (j, wj) in enumerate(sten):
                                if j == idx:
                                    sten[n_] = (j, wj + w)
                                    break
                            else:
                                sten.append((idx, w))
                drop = nq == 0
                if nq:
                    sten = [(j, wj / float(nq)) for j, wj in sten]
                firsts, seconds = {}, {}
                for t in tr:
                    e = np.eye(3, dtype=int)[t]
                    if drop:
                        continue
                    p, m_ = tuple(ca + e), tuple(ca - e)
                    p2, m2 = tuple(ca + 2 * e), tuple(ca - 2 * e)
                    if box_good(ca - e, ca + e):
                        seconds[t] = [(m_, 1.0), (c, -2.0), (p, 1.0)]
                        firsts[t] = [(m_, -0.5), (c, 0.0), (p, 0.5)]
                    elif box_good(ca, ca + 2 * e):
                        seconds[t] = [(c, 1.0), (p, -2.0), (p2, 1.0)]
                        firsts[t] = [(c, -1.5), (p, 2.0), (p2, -0.5)]
                    elif box_good(ca - 2 * e, ca):
                        seconds[t] = [(m2, 1.0), (m_, -2.0), (c, 1.0)]
                        firsts[t] = [(m2, 0.5), (m_, -2.0), (c, 1.5)]
                    else:
                        drop = True                            # m_dropOrd(iv) = true: affects this and later directions
                        if p in good:
                            firsts[t] = [(c, -1.0), (p, 1.0)]
                        elif m_ in good:
                            firsts[t] = [(m_, -1.0), (c, 1.0)]
                        else:
                            firsts[t] = [(c, 0.0)]
                for t in tr:
                    slope[t] = sten_sum(firsts[t]) / self.dxc[t] if t in firsts else 0.0   # no stencil was built: empty sum
                    curv[t] = 0.0 if (drop or t not in seconds) else sten_sum(seconds[t]) / (self.dxc[t] * self.dxc[t])
                mixed = 0.0 if drop else sten_sum(sten) / (self.dxc[tr[1]] * self.dxc[tr[0]])
            deriv[c] = (slope, curv, mixed)
        self.derivs[(b, d, side)] = deriv

        n_hat = np.eye(3, dtype=int)[d] * side
        h = self.dxf[d]
        nref = int(self.ref[d])
        for f in fine_ivs:
            fa = np.array(f)
            c = tuple(fa // self.ref)
            slope, curv, mixed = deriv[c]
            x = [(f[t] + 0.5) * self.dxf[t] - (c[t] + 0.5) * self.dxc[t] for t in tr]
            ps = phic(c) + (slope[tr[0]] * x[0] + curv[tr[0]] * x[0] * x[0] * 0.5) + (slope[tr[1]] * x[1] + curv[tr[1]] * x[1] * x[1] * 0.5) \
                + mixed * x[0] * x[1]
            # MAPPEDQUADINTERP: pa, pb = second / first fine cell inside, phistar sits at the ghost position
            pa, pb = phif(tuple(fa - 2 * n_hat)), phif(tuple(fa - n_hat))
            mult = (2.0 / (h * h)) / float(nref * nref + 4 * nref + 3)
            a_ = mult * (2.0 * ps + float(nref + 1) * pa - float(nref + 3) * pb)
            b_ = (pb - pa) * (1.0 / h) - a_ * h
            out[f] = (4.0 * h * h) * a_ + b_ * (2.0 * h) + pa
