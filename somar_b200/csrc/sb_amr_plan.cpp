// sb_amr_plan.cpp -- host-only planning for the quadratic coarse-fine ghost interpolation of a refined patch
// (the kernel that consumes these records is next round's work).
//
// Reference: MappedQuadCFStencil::define / buildStencils
// (Grade2_AnisotropicChombo/QuadCFInterp/MappedCFStencil.cpp:833-995, 997-1237) decide, for every coarse
// cell under the fine ghost layer of one side of one fine box, how the tangential first, second and mixed
// derivatives of the coarse field are formed: centred where the cell and all its tangential neighbours are
// usable ("standard"), one-sided three-point stencils otherwise, and reduced order ("dropOrd") when neither
// side has room; MappedQuadCFStencil::compute*Derivative (:379-598) evaluate them and zero a derivative
// whose stencil leaves the coarse buffer (the coarsened fine box grown by 2).  Here the outcome is one
// record of plain weights per coarse cell, so that a kernel only multiplies and adds.
//
// The same logic, in numpy, is tests/amr_cfinterp_spec.py, which reproduces the reference bit for bit; the
// CPU test tests/test_host_cpu.py::test_cf_stencil_plan compares the two.
#include <algorithm>
#include <array>
#include <set>

#include "sb_core.h"

namespace sb {

namespace {
using C3 = std::array<int, 3>;

struct Planner {
    Box3              dom;
    int               periodic[3];
    int               ref[3];
    std::vector<Box3> fine, crse;  // fine boxes and their coarsened images
    const std::vector<Box3>* crseGrids = nullptr;  // the coarser level's own boxes (null: it covers its domain)
    // "points in stencils which fell off coarse grid" (ivsEdgeOfEarth, MappedCFStencil.cpp:948-969)
    bool onCoarseGrid(const std::array<int, 3>& c) const
    {
        if (!crseGrids) return true;
        for (const Box3& b : *crseGrids)
            if (inBox(b, c)) return true;
        return false;
    }

    bool inDomain(const C3& c) const
    {
        for (int d = 0; d < 3; ++d)
            if (!periodic[d] && (c[d] < dom.lo[d] || c[d] > dom.hi[d])) return false;
        return true;
    }
    static bool inBox(const Box3& b, const C3& c)
    {
        for (int d = 0; d < 3; ++d)
            if (c[d] < b.lo[d] || c[d] > b.hi[d]) return false;
        return true;
    }
    bool covered(const C3& c) const
    {
        for (const Box3& b : crse)
            if (inBox(b, c)) return true;
        return false;
    }
    bool fineCovered(const C3& f) const
    {
        for (const Box3& b : fine)
            if (inBox(b, f)) return true;
        return false;
    }
};

C3 plus(const C3& a, int d, int n) { C3 r = a; r[d] += n; return r; }
}  // namespace

// One record per coarse cell: first[2][5] and second[2][5] are weights of phi(c + o e_t), o = -2..2, for the
// two tangential directions t in ascending order (to be divided by dx_t and dx_t^2); mixed[3][3] are weights
// of phi(c + o0 e_t0 + o1 e_t1), index [o1 + 1][o0 + 1] (to be divided by dx_t0 dx_t1).
void planCFStencils(const Box3& dom, const int periodic[3], const int ref[3], const std::vector<Box3>& fineBoxes, int box, int dir,
                    int side, std::vector<int>& cells, std::vector<double>& wFirst, std::vector<double>& wSecond,
                    std::vector<double>& wMixed, int dim, const std::vector<Box3>* crseBoxes)
{
    if (box < 0 || box >= (int)fineBoxes.size() || dir < 0 || dir > 2 || (side != 0 && side != 1)) SB_FAIL("bad box / dir / side");
    Planner P;
    P.dom = dom;
    for (int d = 0; d < 3; ++d) { P.periodic[d] = periodic[d]; P.ref[d] = ref[d]; if (ref[d] < 1) SB_FAIL("bad refinement ratio"); }
    P.fine = fineBoxes;
    P.crseGrids = crseBoxes;
    for (const Box3& b : fineBoxes) {
        if (!coarsenable(b, ref)) SB_FAIL("fine boxes must be coarsenable by the refinement ratio");
        P.crse.push_back(coarsen(b, ref));
    }
    for (const Box3& c : P.crse)
        for (int d = 0; d < 3; ++d)
            if (periodic[d] && (c.lo[d] - 2 < dom.lo[d] || c.hi[d] + 2 > dom.hi[d]))
                SB_FAIL("refined patch within 2 coarse cells of a periodic boundary: not supported");
    const Box3& fb = fineBoxes[box];
    const Box3& cb = P.crse[box];
    // tangential directions, ascending (vinttran); a 2-D build (slots 0 and 2) has one and no mixed derivative
    // (MappedCFStencil.cpp:1003-1015, 1022: the mixed stencils exist "only in the case of 3 dimensions")
    int         tr[2] = {0, 0}, nt = 0;
    for (int t = 0; t < 3; ++t)
        if (t != dir && !(dim == 2 && t == 1)) tr[nt++] = t;
    if (dim == 2 && dir == 1) SB_FAIL("direction 1 is not a direction of a 2-D build");

    // coarse cells under the fine ghost layer of this side that are coarse-fine ghosts
    std::set<C3> base;
    {
        Box3 g = fb;
        g.lo[dir] = g.hi[dir] = side ? fb.hi[dir] + 1 : fb.lo[dir] - 1;
        for (int k = g.lo[2]; k <= g.hi[2]; ++k)
            for (int j = g.lo[1]; j <= g.hi[1]; ++j)
                for (int i = g.lo[0]; i <= g.hi[0]; ++i) {
                    const C3 f{i, j, k};
                    const C3 c{fdiv(i, ref[0]), fdiv(j, ref[1]), fdiv(k, ref[2])};
                    if (P.inDomain(c) && !P.fineCovered(f)) base.insert(c);
                }
    }
    cells.clear(); wFirst.clear(); wSecond.clear(); wMixed.clear();
    if (base.empty()) return;

    // usable cells of the coarse slab next to the face, and the standard ones
    Box3 g2 = cb, g1 = cb;
    g2.lo[dir] = g2.hi[dir] = g1.lo[dir] = g1.hi[dir] = side ? cb.hi[dir] + 1 : cb.lo[dir] - 1;
    for (int q = 0; q < nt; ++q) { g2.lo[tr[q]] -= 2; g2.hi[tr[q]] += 2; g1.lo[tr[q]] -= 1; g1.hi[tr[q]] += 1; }
    std::set<C3> good, stdc;
    for (int k = g2.lo[2]; k <= g2.hi[2]; ++k)
        for (int j = g2.lo[1]; j <= g2.hi[1]; ++j)
            for (int i = g2.lo[0]; i <= g2.hi[0]; ++i) {
                const C3 c{i, j, k};
                if (P.inDomain(c) && !P.covered(c) && P.onCoarseGrid(c)) good.insert(c);
            }
    for (const C3& c : good)
        if (Planner::inBox(g1, c)) stdc.insert(c);
    for (int q = 0; q < nt; ++q) {  // IntVectSet::grow(t, -1)
        std::set<C3> e;
        for (const C3& c : stdc)
            if (stdc.count(plus(c, tr[q], -1)) && stdc.count(plus(c, tr[q], 1))) e.insert(c);
        stdc.swap(e);
    }
    Box3 buf = cb;
    for (int d = 0; d < 3; ++d) { if (dim == 2 && d == 1) continue; buf.lo[d] -= 2; buf.hi[d] += 2; }
    auto boxGood = [&](C3 lo, C3 hi) {
        for (int k = lo[2]; k <= hi[2]; ++k)
            for (int j = lo[1]; j <= hi[1]; ++j)
                for (int i = lo[0]; i <= hi[0]; ++i)
                    if (!good.count(C3{i, j, k})) return false;
        return true;
    };

    for (const C3& c : base) {
        double f1[2][5] = {}, f2[2][5] = {}, mx[3][3] = {};
        bool   touched[3][3] = {};  // cells of the 3 x 3 neighbourhood that some quadrant put into the mixed stencil
        if (stdc.count(c)) {
            for (int q = 0; q < nt; ++q) {
                f1[q][1] = -0.5; f1[q][3] = 0.5;
                f2[q][1] = 1.0; f2[q][2] = -2.0; f2[q][3] = 1.0;
            }
            if (nt == 2) { mx[2][2] = 0.25; mx[0][0] = 0.25; mx[0][2] = -0.25; mx[2][0] = -0.25; }  // (ur + ll - lr - ul) / 4
        } else {
            // mixed derivative: the quadrant boxes as the reference builds them, weight -1 on the box's low and
            // high corner and +1 on the other two, averaged over the usable quadrants
            const int e0 = tr[0], e1 = tr[1];
            C3        qlo[4] = {plus(c, e0, -1), c, plus(c, e1, -1), plus(plus(c, e0, -1), e1, -1)};
            C3        qhi[4] = {plus(c, e1, 1), plus(plus(c, e0, 1), e1, 1), plus(c, e0, 1), c};
            int       nq = 0;
            for (int q = 0; q < 4 && nt == 2; ++q)
                if (boxGood(qlo[q], qhi[q])) {
                    ++nq;
                    for (int b1 = 0; b1 < 2; ++b1)
                        for (int b0 = 0; b0 < 2; ++b0) {
                            const int o0 = qlo[q][e0] + b0 - c[e0], o1 = qlo[q][e1] + b1 - c[e1];
                            mx[o1 + 1][o0 + 1] += (b0 == b1) ? -1.0 : 1.0;
                            touched[o1 + 1][o0 + 1] = true;
                        }
                }
            bool drop = nt == 2 && nq == 0;  // m_dropOrd starts false in a 2-D build (MappedCFStencil.cpp:1018-1019)
            if (nq)
                for (auto& row : mx)
                    for (double& w : row) w /= (double)nq;
            bool haveFirst[2] = {false, false}, haveSecond[2] = {false, false};
            for (int q = 0; q < nt && !drop; ++q) {
                const int t = tr[q];
                if (boxGood(plus(c, t, -1), plus(c, t, 1))) {
                    f2[q][1] = 1.0; f2[q][2] = -2.0; f2[q][3] = 1.0;
                    f1[q][1] = -0.5; f1[q][3] = 0.5;
                    haveFirst[q] = haveSecond[q] = true;
                } else if (boxGood(c, plus(c, t, 2))) {
                    f2[q][2] = 1.0; f2[q][3] = -2.0; f2[q][4] = 1.0;
                    f1[q][2] = -1.5; f1[q][3] = 2.0; f1[q][4] = -0.5;
                    haveFirst[q] = haveSecond[q] = true;
                } else if (boxGood(plus(c, t, -2), c)) {
                    f2[q][0] = 1.0; f2[q][1] = -2.0; f2[q][2] = 1.0;
                    f1[q][0] = 0.5; f1[q][1] = -2.0; f1[q][2] = 1.5;
                    haveFirst[q] = haveSecond[q] = true;
                } else {
                    drop = true;  // m_dropOrd(iv) = true: later directions get no stencil at all
                    if (good.count(plus(c, t, 1))) { f1[q][2] = -1.0; f1[q][3] = 1.0; }
                    else if (good.count(plus(c, t, -1))) { f1[q][1] = -1.0; f1[q][2] = 1.0; }
                    haveFirst[q] = true;
                }
            }
            if (drop) {  // second and mixed derivatives are dropped for the cell
                for (auto& row : f2) std::fill(row, row + 5, 0.0);
                for (auto& row : mx) std::fill(row, row + 3, 0.0);
            }
            (void)haveFirst; (void)haveSecond;
            // a derivative whose stencil reaches outside the coarse buffer is zero (keepzer)
            auto outside = [&](const C3& p) { return !Planner::inBox(buf, p); };
            for (int q = 0; q < nt; ++q) {
                bool z1 = false, z2 = false;
                for (int o = -2; o <= 2; ++o) {
                    if (f1[q][o + 2] != 0.0 && outside(plus(c, tr[q], o))) z1 = true;
                    if (f2[q][o + 2] != 0.0 && outside(plus(c, tr[q], o))) z2 = true;
                }
                if (z1) std::fill(f1[q], f1[q] + 5, 0.0);
                if (z2) std::fill(f2[q], f2[q] + 5, 0.0);
            }
            bool zm = false;
            for (int o1 = -1; o1 <= 1; ++o1)
                for (int o0 = -1; o0 <= 1; ++o0)
                    if (touched[o1 + 1][o0 + 1] && outside(plus(plus(c, tr[0], o0), tr[1], o1))) zm = true;
            if (zm)
                for (auto& row : mx) std::fill(row, row + 3, 0.0);
        }
        cells.insert(cells.end(), c.begin(), c.end());
        for (int q = 0; q < 2; ++q) wFirst.insert(wFirst.end(), f1[q], f1[q] + 5);
        for (int q = 0; q < 2; ++q) wSecond.insert(wSecond.end(), f2[q], f2[q] + 5);
        for (auto& row : mx) wMixed.insert(wMixed.end(), row, row + 3);
    }
}

}  // namespace sb
