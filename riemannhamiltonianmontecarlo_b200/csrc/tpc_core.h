// Thread-per-chain dense linear algebra on a packed triangle (D <= 32): the arithmetic core of chain_tpc.cuh.
//
// The warp-per-chain kernels (chain_kernels.cuh) spend ~2400 warp instructions per 25 x 25 Cholesky solve, most of them
// on lanes that hold the unused triangle, column broadcasts and per-column reciprocal square roots executed by all 32
// lanes.  Here ONE THREAD owns one chain: every instruction of a warp does useful work for 32 chains, there is no
// cross-lane traffic at all, and the only per-column overhead is one rsqrt per chain.  The price is residency (the
// matrices live in shared memory: 32 chains x 351 packed doubles = 93 KB per warp at D = 25, two warps per SM), so the
// routines below are written for instruction-level parallelism inside one thread: TILE independent accumulators per
// inner loop, loads that do not depend on the accumulators.
//
// Storage: element e of the thread's array sits at A[e * kTpcStride] (kTpcStride = 33 doubles: the 32 threads of a
// warp hit 32 different banks when they touch the same element, and so do the 32 lanes of a transposing fill that
// writes 32 consecutive elements of ONE chain).
//
// Augmented packed layout ("aug"): the lower triangle of the (D+1) x (D+1) matrix [[G, .], [b^T, .]] by columns,
// column j holding rows j..D: element (i, j) at tpc_col(j, D + 1) + i - j.  Row D carries a right-hand side, so the forward
// substitution of a solve is simply the last row of the factorisation (rmhmc.py:121 solves G u = p).
//
// These functions compile as host code too (tests/native/tpc_host_test.cpp checks them against a plain dense
// reference on the CPU); on the device they are inlined into the kernels of chain_tpc.cuh.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define RMHMC_HD __host__ __device__ __forceinline__
#else
#define RMHMC_HD inline
#endif

namespace rmhmc {

constexpr int kTpcStride = 33;

// first element of column j when every column j holds DL - j rows (DL = D: plain triangle, rows j..D-1, the layout
// the metric builds write, pair_index(j, i, D); DL = D + 1: augmented, rows j..D)
RMHMC_HD int tpc_col(int j, int dl) { return j * dl - j * (j - 1) / 2; }
RMHMC_HD int tpc_elems(int d, int dl) { return tpc_col(d, dl); }      // elements of columns 0..D-1

RMHMC_HD double tpc_rsqrt(double a) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));      // 20-bit seed, two Newton steps: ~1 ulp; NaN for a < 0
    const double h = 0.5 * a;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
#else
    return 1.0 / sqrt(a);
#endif
}

// In-place Cholesky factorisation, column by column (Crout), of the packed matrix with DL - j rows in column j (rows
// j..DL-1; DL = D + 1 carries the right-hand side as row D): on exit element (i, j), j < i < D, holds L[i][j]; the
// DIAGONAL holds 1 / L[j][j]; row D (if any) holds y = L^-1 b.  Rows are processed TILE at a time with independent
// accumulators; a tile may run past the last row (it then reads the head of the next column -- the caller pads the array
// by TILE elements -- and the surplus accumulators are dropped).
// Returns 0.5 log|G| = log prod_j L[j][j] (rmhmc.py:171), the product folded into the sum of logs every 8 columns.
// diag_out (may be null): receives the true diagonal, diag_out[j * kTpcStride] = L[j][j].
template <int TILE>
RMHMC_HD double tpc_cholesky(double* A, int D, int DL, double* diag_out = nullptr) {
    constexpr int ST = kTpcStride;
    double logdet = 0.0, prod = 1.0;
    int cj = 0;                                     // tpc_col(j) - j: element (i, j) at cj + i
    for (int j = 0; j < D; ++j) {
        double dinv = 0.0;
        for (int i0 = j; i0 < DL; i0 += TILE) {
            double acc[TILE];
#pragma unroll
            for (int m = 0; m < TILE; ++m) acc[m] = A[(cj + i0 + m) * ST];
            const double* pk = A;                   // column k: element (i, k) at pk[i * ST]
            for (int k = 0; k < j; ++k) {
                const double ljk = pk[j * ST];
#pragma unroll
                for (int m = 0; m < TILE; ++m) acc[m] = fma(-ljk, pk[(i0 + m) * ST], acc[m]);
                pk += (DL - k - 1) * ST;
            }
            int m0 = 0;
            if (i0 == j) {
                dinv = tpc_rsqrt(acc[0]);
                A[(cj + j) * ST] = dinv;
                const double ljj = acc[0] * dinv;
                prod *= ljj;
                if (diag_out) diag_out[j * ST] = ljj;
                m0 = 1;
            }
#pragma unroll
            for (int m = 0; m < TILE; ++m)
                if (m >= m0 && i0 + m < DL) A[(cj + i0 + m) * ST] = acc[m] * dinv;
        }
        if ((j & 7) == 7 || j == D - 1) { logdet += log(prod); prod = 1.0; }
        cj += DL - j - 1;
    }
    return logdet;
}

// The same factorisation two columns at a time: the pair (j, j + 1) shares the loads of L[i][k], k < j, so a k-step of
// a TR-row tile is TR + 2 shared-memory loads for 2 TR FMAs (tpc_cholesky: TR + 1 for TR), and 2 TR independent
// accumulators cover the FP64 latency of a single resident warp per scheduler.  Every entry is accumulated in exactly
// the order of tpc_cholesky (k ascending, the k = j term of column j + 1 last), so the two are bit-identical.
template <int TR>
RMHMC_HD double tpc_cholesky2(double* A, int D, int DL, double* diag_out = nullptr) {
    constexpr int ST = kTpcStride;
    double logdet = 0.0, prod = 1.0;
    int cj = 0;                                     // tpc_col(j) - j: element (i, j) at cj + i
    int j = 0;
    for (; j + 1 < D; j += 2) {
        const int j1 = j + 1, cj1 = cj + DL - j - 1;
        double dinv0 = 0.0, dinv1 = 0.0, lj1j = 0.0;
        for (int i0 = j; i0 < DL; i0 += TR) {
            double a0[TR], a1[TR];
#pragma unroll
            for (int m = 0; m < TR; ++m) {
                a0[m] = A[(cj + i0 + m) * ST];
                a1[m] = A[(cj1 + i0 + m) * ST];     // row j of column j + 1 does not exist: reads a neighbour, never stored
            }
            // software pipeline: the operands of step k + 1 are loaded before the FMAs of step k (a single resident
            // warp per scheduler has nobody else to hide the shared-memory latency behind); the load after the last
            // step reads column j itself, which exists, and is dropped
            const double* pk = A;                   // column k: element (i, k) at pk[i * ST]
            double l0 = pk[j * ST], l1 = pk[j1 * ST], li[TR];
#pragma unroll
            for (int m = 0; m < TR; ++m) li[m] = pk[(i0 + m) * ST];
            for (int k = 0; k < j; ++k) {
                pk += (DL - k - 1) * ST;
                const double n0 = pk[j * ST], n1 = pk[j1 * ST];
                double ni[TR];
#pragma unroll
                for (int m = 0; m < TR; ++m) ni[m] = pk[(i0 + m) * ST];
#pragma unroll
                for (int m = 0; m < TR; ++m) {
                    a0[m] = fma(-l0, li[m], a0[m]);
                    a1[m] = fma(-l1, li[m], a1[m]);
                }
                l0 = n0; l1 = n1;
#pragma unroll
                for (int m = 0; m < TR; ++m) li[m] = ni[m];
            }
            if (i0 == j) {                          // rows j, j + 1 head the first tile: both pivots
                dinv0 = tpc_rsqrt(a0[0]);
                const double ljj = a0[0] * dinv0;
                lj1j = a0[1] * dinv0;
                const double p1 = fma(-lj1j, lj1j, a1[1]);
                dinv1 = tpc_rsqrt(p1);
                const double lj1j1 = p1 * dinv1;
                prod *= ljj;
                prod *= lj1j1;
                if (diag_out) { diag_out[j * ST] = ljj; diag_out[j1 * ST] = lj1j1; }
                A[(cj + j) * ST] = dinv0;
                A[(cj + j1) * ST] = lj1j;
                A[(cj1 + j1) * ST] = dinv1;
#pragma unroll
                for (int m = 2; m < TR; ++m) {
                    if (i0 + m < DL) {
                        const double l0 = a0[m] * dinv0;
                        A[(cj + i0 + m) * ST] = l0;
                        A[(cj1 + i0 + m) * ST] = fma(-l0, lj1j, a1[m]) * dinv1;
                    }
                }
            } else {
#pragma unroll
                for (int m = 0; m < TR; ++m) {
                    if (i0 + m < DL) {
                        const double l0 = a0[m] * dinv0;
                        A[(cj + i0 + m) * ST] = l0;
                        A[(cj1 + i0 + m) * ST] = fma(-l0, lj1j, a1[m]) * dinv1;
                    }
                }
            }
        }
        if ((j & 7) == 6 || j1 == D - 1) { logdet += log(prod); prod = 1.0; }
        cj = cj1 + DL - j1 - 1;
    }
    if (j < D) {                                    // odd D: the last column on its own
        double dinv = 0.0;
        for (int i0 = j; i0 < DL; i0 += TR) {
            double acc[TR];
#pragma unroll
            for (int m = 0; m < TR; ++m) acc[m] = A[(cj + i0 + m) * ST];
            const double* pk = A;
            for (int k = 0; k < j; ++k) {
                const double ljk = pk[j * ST];
#pragma unroll
                for (int m = 0; m < TR; ++m) acc[m] = fma(-ljk, pk[(i0 + m) * ST], acc[m]);
                pk += (DL - k - 1) * ST;
            }
            int m0 = 0;
            if (i0 == j) {
                dinv = tpc_rsqrt(acc[0]);
                A[(cj + j) * ST] = dinv;
                const double ljj = acc[0] * dinv;
                prod *= ljj;
                if (diag_out) diag_out[j * ST] = ljj;
                m0 = 1;
            }
#pragma unroll
            for (int m = 0; m < TR; ++m)
                if (m >= m0 && i0 + m < DL) A[(cj + i0 + m) * ST] = acc[m] * dinv;
        }
        logdet += log(prod);
    }
    return logdet;
}

// Back substitution L^T x = y with the factor left by tpc_cholesky (DL = D + 1, y in row D).  X[i * kTpcStride] <- x_i.
RMHMC_HD void tpc_backsolve(const double* A, double* X, int D) {
    constexpr int ST = kTpcStride;
    const int DL = D + 1;
    for (int i = D - 1; i >= 0; --i) {
        const double* col = A + (tpc_col(i, DL) - i) * ST;      // element (k, i) at col[k * ST]
        double s0 = col[D * ST], s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int k = i + 1;
        for (; k + 3 < D; k += 4) {
            s0 = fma(-col[k * ST], X[k * ST], s0);
            s1 = fma(-col[(k + 1) * ST], X[(k + 1) * ST], s1);
            s2 = fma(-col[(k + 2) * ST], X[(k + 2) * ST], s2);
            s3 = fma(-col[(k + 3) * ST], X[(k + 3) * ST], s3);
        }
        for (; k < D; ++k) s0 = fma(-col[k * ST], X[k * ST], s0);
        X[i * ST] = ((s0 + s1) + (s2 + s3)) * col[i * ST];
    }
}

// In place: L (diagonal holding 1 / L[j][j]) -> M = L^-1 (lower triangular, TRUE diagonal 1 / L[j][j]).  Row by row:
// M[i][j] = -(sum_{k=j}^{i-1} L[i][k] M[k][j]) / L[i][i]; rows above i are already M, row i is still L.  Within row i
// the columns go in ascending tiles of TILE: tile j0 needs L[i][k] only for k >= j0, so overwriting L[i][j0..] after the
// tile is safe.  The TILE columns of a tile are independent sums sharing the L[i][k] loads (M[k][j] = 0 above the
// diagonal: the first TILE-1 values of k take part in fewer columns).  Columns >= i of the last tile are clamped to
// i - 1 (computed twice, stored once).
template <int TILE>
RMHMC_HD void tpc_invert_lower(double* A, int D, int DL) {
    constexpr int ST = kTpcStride;
    for (int i = 1; i < D; ++i) {
        const double dinv_i = A[tpc_col(i, DL) * ST];
        for (int j0 = 0; j0 < i; j0 += TILE) {
            const double* mc[TILE];                 // column j0 + m: element (k, j) at mc[m][k * ST]
            double acc[TILE];
#pragma unroll
            for (int m = 0; m < TILE; ++m) {
                const int j = j0 + m < i ? j0 + m : i - 1;
                mc[m] = A + (tpc_col(j, DL) - j) * ST;
                acc[m] = 0.0;
            }
            const double* pk = A + ((tpc_col(j0, DL) - j0) + i) * ST;        // L[i][k], k = j0
            int k = j0;
#pragma unroll
            for (int kk = 0; kk < TILE - 1; ++kk) {                          // head: column j0 + m starts at k = j0 + m
                if (k < i) {
                    const double lik = *pk;
#pragma unroll
                    for (int m = 0; m < TILE; ++m)
                        if (m <= kk && j0 + m < i) acc[m] = fma(lik, mc[m][k * ST], acc[m]);
                    pk += (DL - k - 1) * ST;
                    ++k;
                }
            }
            for (; k < i; ++k) {
                const double lik = *pk;
#pragma unroll
                for (int m = 0; m < TILE; ++m) acc[m] = fma(lik, mc[m][k * ST], acc[m]);
                pk += (DL - k - 1) * ST;
            }
#pragma unroll
            for (int m = 0; m < TILE; ++m)
                if (j0 + m < i) const_cast<double*>(mc[m])[i * ST] = -acc[m] * dinv_i;
        }
    }
}

// In place: M (lower triangular) -> the lower triangle of M^T M = G^-1: entry (i, j), j <= i, is
// sum_{k >= i} M[k][i] M[k][j].  Rows ascending: row i reads rows k >= i only, and of row i itself each tile reads its
// own entries (k = i term) before it overwrites them; the tile holding the diagonal comes last.
template <int TILE>
RMHMC_HD void tpc_mtm_lower(double* A, int D, int DL) {
    constexpr int ST = kTpcStride;
    for (int i = 0; i < D; ++i) {
        const double* coli = A + (tpc_col(i, DL) - i) * ST;    // element (k, i) at coli[k * ST]
        for (int j0 = 0; j0 <= i; j0 += TILE) {
            double* mc[TILE];
            double acc[TILE];
#pragma unroll
            for (int m = 0; m < TILE; ++m) {
                const int j = j0 + m <= i ? j0 + m : i;
                mc[m] = A + (tpc_col(j, DL) - j) * ST;
                acc[m] = 0.0;
            }
            for (int k = i; k < D; ++k) {
                const double mki = coli[k * ST];
#pragma unroll
                for (int m = 0; m < TILE; ++m) acc[m] = fma(mki, mc[m][k * ST], acc[m]);
            }
#pragma unroll
            for (int m = 0; m < TILE; ++m)
                if (j0 + m <= i) mc[m][i * ST] = acc[m];
        }
    }
}

}  // namespace rmhmc
