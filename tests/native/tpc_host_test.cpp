// Host check of the thread-per-chain linear algebra core (csrc/tpc_core.h) against plain dense arithmetic.
// Build + run: g++ -O2 -std=c++17 -o tpc_host_test tpc_host_test.cpp && ./tpc_host_test   (tests/test_host_cpu.py does it)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../riemannhamiltonianmontecarlo_b200/csrc/tpc_core.h"

using namespace rmhmc;

static double urand() { return rand() / (double)RAND_MAX - 0.5; }

template <int TILE>
static int check(int D, int lane) {
    const int ST = kTpcStride;
    std::vector<double> G(D * D), b(D);
    // SPD: B B^T + D I
    std::vector<double> B(D * D);
    for (auto& v : B) v = urand();
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
            double s = i == j ? 0.1 : 0.0;
            for (int k = 0; k < D; ++k) s += B[i * D + k] * B[j * D + k];
            G[i * D + j] = s;
        }
    for (auto& v : b) v = urand();
    // dense reference: Cholesky, solve, inverse
    std::vector<double> L(D * D, 0.0);
    for (int j = 0; j < D; ++j) {
        double s = G[j * D + j];
        for (int k = 0; k < j; ++k) s -= L[j * D + k] * L[j * D + k];
        L[j * D + j] = sqrt(s);
        for (int i = j + 1; i < D; ++i) {
            double t = G[i * D + j];
            for (int k = 0; k < j; ++k) t -= L[i * D + k] * L[j * D + k];
            L[i * D + j] = t / L[j * D + j];
        }
    }
    std::vector<double> y(D), x(D);
    for (int i = 0; i < D; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i * D + k] * y[k]; y[i] = s / L[i * D + i]; }
    for (int i = D - 1; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < D; ++k) s -= L[k * D + i] * x[k]; x[i] = s / L[i * D + i]; }
    double logdet = 0.0;
    for (int i = 0; i < D; ++i) logdet += log(L[i * D + i]);
    int bad = 0;
    double worst = 0.0;
    auto cmp = [&](double got, double want, const char* what, int i, int j) {
        double e = fabs(got - want) / (1e-300 + fabs(want) + 1e-3);
        if (e > worst) worst = e;
        if (!(e < 1e-11)) { if (bad < 5) printf("  D=%d TILE=%d %s(%d,%d): got %.17g want %.17g\n", D, TILE, what, i, j, got, want); ++bad; }
    };
    // dense M = L^-1 and G^-1 = M^T M
    std::vector<double> M(D * D, 0.0), IG(D * D, 0.0);
    for (int j = 0; j < D; ++j) {
        M[j * D + j] = 1.0 / L[j * D + j];
        for (int i = j + 1; i < D; ++i) {
            double s = 0.0;
            for (int k = j; k < i; ++k) s += L[i * D + k] * M[k * D + j];
            M[i * D + j] = -s / L[i * D + i];
        }
    }
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += M[k * D + i] * M[k * D + j];
            IG[i * D + j] = s;
        }
    for (int DL = D; DL <= D + 1; ++DL) {   // plain triangle, and augmented with the right-hand side as row D
        const size_t n_el = (size_t)tpc_elems(D, DL) + TILE;
        std::vector<double> A((n_el + D) * ST, 777.0);
        double* a = A.data() + lane;
        for (int j = 0; j < D; ++j) {
            for (int i = j; i < D; ++i) a[(tpc_col(j, DL) + i - j) * ST] = G[i * D + j];
            if (DL > D) a[(tpc_col(j, DL) + D - j) * ST] = b[j];
        }
        std::vector<double> A2(A), dg1((size_t)D * ST, 0.0), dg2((size_t)D * ST, 0.0);
        double ld = tpc_cholesky<TILE>(a, D, DL, dg1.data());
        cmp(ld, logdet, "logdet", 0, DL);
        {   // the two-column variant must be bit-identical (same accumulation order per entry)
            double ld2 = tpc_cholesky2<TILE>(A2.data() + lane, D, DL, dg2.data());
            cmp(ld2, logdet, "logdet2", 0, DL);
            for (int j = 0; j < D; ++j) {
                if (dg1[j * ST] != dg2[j * ST]) { printf("  diag_out differs at %d\n", j); ++bad; }
                cmp(dg1[j * ST], L[j * D + j], "diag_out", j, j);
                for (int i = j; i < DL; ++i) {
                    const size_t e = (size_t)(tpc_col(j, DL) + i - j) * ST + lane;
                    if (A[e] != A2[e]) { if (bad < 5) printf("  D=%d DL=%d cholesky2 differs at (%d,%d): %.17g vs %.17g\n", D, DL, i, j, A2[e], A[e]); ++bad; }
                }
            }
            for (size_t e = 0; e < A2.size(); ++e)
                if ((int)(e % ST) != lane && A2[e] != 777.0) { printf("  cholesky2 wrote another chain's element %zu\n", e); ++bad; break; }
        }
        for (int j = 0; j < D; ++j) {
            cmp(a[tpc_col(j, DL) * ST], 1.0 / L[j * D + j], "dinv", j, j);
            for (int i = j + 1; i < D; ++i) cmp(a[(tpc_col(j, DL) + i - j) * ST], L[i * D + j], "L", i, j);
            if (DL > D) cmp(a[(tpc_col(j, DL) + D - j) * ST], y[j], "y", j, 0);
        }
        if (DL > D) {
            double* X = a + n_el * ST;
            tpc_backsolve(a, X, D);
            for (int i = 0; i < D; ++i) cmp(X[i * ST], x[i], "x", i, 0);
        }
        tpc_invert_lower<TILE>(a, D, DL);
        for (int j = 0; j < D; ++j)
            for (int i = j; i < D; ++i) cmp(a[(tpc_col(j, DL) + i - j) * ST], M[i * D + j], "M", i, j);
        tpc_mtm_lower<TILE>(a, D, DL);
        for (int j = 0; j < D; ++j)
            for (int i = j; i < D; ++i) cmp(a[(tpc_col(j, DL) + i - j) * ST], IG[i * D + j], "Ginv", i, j);
        for (size_t e = 0; e < A.size(); ++e)
            if ((int)(e % ST) != lane && A[e] != 777.0) { printf("  wrote another chain's element %zu\n", e); ++bad; break; }
    }
    printf("D=%2d TILE=%d lane=%2d worst rel err %.2e %s\n", D, TILE, lane, worst, bad ? "FAIL" : "ok");
    return bad;
}

int main() {
    srand(12345);
    int bad = 0;
    for (int D : {1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 25, 31, 32}) {
        bad += check<4>(D, D % 32);
        bad += check<6>(D, (3 * D) % 32);
        bad += check<8>(D, (7 * D) % 32);
    }
    printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad ? 1 : 0;
}
