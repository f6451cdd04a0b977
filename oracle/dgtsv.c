/* TEST INFRASTRUCTURE ONLY.  Nothing in somar_b200/ may include, link or call this file.
 *
 * dgtsv.c -- LAPACK's DGTSV restated in plain C: the one third-party routine on SOMAR's hot path
 * that is not under /root/reference.  The reference links a bare, unpinned system -llapack
 * (SConstruct_ex.py:94) and calls dgtsv at Grade3_Calculus/Elliptic/PoissonOpF.ChF:794,831,957,1002
 * (one tridiagonal solve per grid column, NRHS = 1).  This follows the published LAPACK 3.x
 * reference algorithm (Gaussian elimination with partial pivoting, row interchange when
 * |d(i)| < |dl(i)|) statement by statement, with no FMA contraction, and is pinned against
 * scipy.linalg.lapack.dgtsv (OpenBLAS's LAPACK) in tests/test_oracle_dgtsv.py.
 */
#include <math.h>

void
dgtsv_(const int* N_, const int* NRHS_, double* DL, double* D, double* DU, double* B, const int* LDB_, int* INFO)
{
    const int N = *N_, NRHS = *NRHS_, LDB = *LDB_;
    *INFO = 0;
    if (N < 0) { *INFO = -1; return; }
    if (NRHS < 0) { *INFO = -2; return; }
    if (LDB < (N > 1 ? N : 1)) { *INFO = -7; return; }
    if (N == 0) return;
    // 1-based helpers
#define dl(i) DL[(i)-1]
#define d(i) D[(i)-1]
#define du(i) DU[(i)-1]
#define b(i, j) B[((i)-1) + (long)((j)-1) * LDB]
    double fact, temp;
    if (NRHS == 1) {
        for (int i = 1; i <= N - 2; ++i) {
            if (fabs(d(i)) >= fabs(dl(i))) {
                // No row interchange required
                if (d(i) != 0.0) {
                    fact     = dl(i) / d(i);
                    d(i + 1) = d(i + 1) - fact * du(i);
                    b(i + 1, 1) = b(i + 1, 1) - fact * b(i, 1);
                } else {
                    *INFO = i;
                    return;
                }
                dl(i) = 0.0;
            } else {
                // Interchange rows I and I+1
                fact      = d(i) / dl(i);
                d(i)      = dl(i);
                temp      = d(i + 1);
                d(i + 1)  = du(i) - fact * temp;
                dl(i)     = du(i + 1);
                du(i + 1) = -fact * dl(i);
                du(i)     = temp;
                temp      = b(i, 1);
                b(i, 1)   = b(i + 1, 1);
                b(i + 1, 1) = temp - fact * b(i + 1, 1);
            }
        }
        if (N > 1) {
            const int i = N - 1;
            if (fabs(d(i)) >= fabs(dl(i))) {
                if (d(i) != 0.0) {
                    fact     = dl(i) / d(i);
                    d(i + 1) = d(i + 1) - fact * du(i);
                    b(i + 1, 1) = b(i + 1, 1) - fact * b(i, 1);
                } else {
                    *INFO = i;
                    return;
                }
            } else {
                fact     = d(i) / dl(i);
                d(i)     = dl(i);
                temp     = d(i + 1);
                d(i + 1) = du(i) - fact * temp;
                du(i)    = temp;
                temp     = b(i, 1);
                b(i, 1)  = b(i + 1, 1);
                b(i + 1, 1) = temp - fact * b(i + 1, 1);
            }
        }
        if (d(N) == 0.0) {
            *INFO = N;
            return;
        }
    } else {
        for (int i = 1; i <= N - 2; ++i) {
            if (fabs(d(i)) >= fabs(dl(i))) {
                if (d(i) != 0.0) {
                    fact     = dl(i) / d(i);
                    d(i + 1) = d(i + 1) - fact * du(i);
                    for (int j = 1; j <= NRHS; ++j) b(i + 1, j) = b(i + 1, j) - fact * b(i, j);
                } else {
                    *INFO = i;
                    return;
                }
                dl(i) = 0.0;
            } else {
                fact      = d(i) / dl(i);
                d(i)      = dl(i);
                temp      = d(i + 1);
                d(i + 1)  = du(i) - fact * temp;
                dl(i)     = du(i + 1);
                du(i + 1) = -fact * dl(i);
                du(i)     = temp;
                for (int j = 1; j <= NRHS; ++j) {
                    temp        = b(i, j);
                    b(i, j)     = b(i + 1, j);
                    b(i + 1, j) = temp - fact * b(i + 1, j);
                }
            }
        }
        if (N > 1) {
            const int i = N - 1;
            if (fabs(d(i)) >= fabs(dl(i))) {
                if (d(i) != 0.0) {
                    fact     = dl(i) / d(i);
                    d(i + 1) = d(i + 1) - fact * du(i);
                    for (int j = 1; j <= NRHS; ++j) b(i + 1, j) = b(i + 1, j) - fact * b(i, j);
                } else {
                    *INFO = i;
                    return;
                }
            } else {
                fact     = d(i) / dl(i);
                d(i)     = dl(i);
                temp     = d(i + 1);
                d(i + 1) = du(i) - fact * temp;
                du(i)    = temp;
                for (int j = 1; j <= NRHS; ++j) {
                    temp        = b(i, j);
                    b(i, j)     = b(i + 1, j);
                    b(i + 1, j) = temp - fact * b(i + 1, j);
                }
            }
        }
        if (d(N) == 0.0) {
            *INFO = N;
            return;
        }
    }
    // Back solve with the matrix U from the factorization.
    for (int j = 1; j <= NRHS; ++j) {
        b(N, j) = b(N, j) / d(N);
        if (N > 1) b(N - 1, j) = (b(N - 1, j) - du(N - 1) * b(N, j)) / d(N - 1);
        for (int i = N - 2; i >= 1; --i) b(i, j) = (b(i, j) - du(i) * b(i + 1, j) - dl(i) * b(i + 2, j)) / d(i);
    }
#undef dl
#undef d
#undef du
#undef b
}
