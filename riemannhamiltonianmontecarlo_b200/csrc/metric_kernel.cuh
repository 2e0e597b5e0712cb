// Fused metric build: for a tile of chains, one pass over the design matrix computes
//   f = X theta, p = sigma(f), v = p(1-p)                        (rmhmc.py:51-53, :116-118, :134-136)
//   G = X^T diag(v) X + I/alpha   (packed upper triangle)        (rmhmc.py:57, :119, :137)
// and, in CLOSING mode (the build that ends a leapfrog step),
//   X^T (t - p)                                                  (rmhmc.py:140, first term)
//   loglik = f^T t - sum log(1+e^f)                              (rmhmc.py:167-168)
//   c_n = v_n (1 - 2 p_n)  -> cbuf, the A operand of the partials build (rmhmc.py:148-149)
//
// Formulated as a chain-batched contraction on the FP64 tensor cores (DMMA.8x8x4):
//   F^T[chains x rows]   = Theta[chains x D] . X^T                (K = D)
//   G  [chains x pairs]  = V[chains x rows]  . KR2(X)[rows x pairs]   (K = N rows)
// where KR2(X)[n, (a,b)] = x_na x_nb is never materialised: each lane forms its B-fragment
// element from two staged X values.  X row blocks arrive by 1-D bulk TMA (UBLKCP) into a
// 3-stage mbarrier ring; V never leaves the SM.
#pragma once
#include "common.cuh"

namespace rmhmc {

constexpr int kMetricChains = 32;    // chains per CTA
constexpr int kMetricRows = 32;      // design-matrix rows per staged block
constexpr int kMetricWarps = 8;
constexpr int kMetricStages = 3;
constexpr int kMetricVS = kMetricRows + 4;   // smem stride of the V/R tiles (4*odd)

struct MetricArgs {
    const double* x;          // [Np][XS]  zero-padded rows/cols, label t in column XS-1
    const uchar2* pair_tab;   // [P2p]     packed column -> (a, b), a <= b
    const double* theta;      // [C][D]    positions to evaluate at
    double* g_out;            // [C][P2p]  packed metric
    double* grad_out;         // [C][D]    (closing) X^T (t - p)
    double* loglik_out;       // [C]       (closing)
    double* cbuf;             // [C][Np]   (closing)
    const unsigned char* skip;// [C] or null: chains whose outputs nobody will read
    int n_chains, n_rows, n_rows_pad, dim, xs, p2, p2p;
    double alpha_inv;
};

__host__ inline size_t metric_smem_bytes(int xs) {
    size_t b = 0;
    b += (size_t)kMetricStages * kMetricRows * xs * 8;  // X ring
    b += (size_t)kMetricChains * xs * 8;                // Theta tile
    b += 2 * (size_t)kMetricChains * kMetricVS * 8;     // V and R tiles
    b += (size_t)kMetricWarps * 8 * 8;                  // loglik partials
    b += 64;                                            // mbarriers
    return b;
}

#ifdef __CUDACC__
// Logistic terms of one (chain, row) pair.  One exp; p and 1-p are both formed without
// cancellation.  Overflow quirk of the reference kept: exp(f) overflows for f > ~709.78, which
// makes its gradient NaN (inf/inf, rmhmc.py:100) and its log-likelihood -inf (rmhmc.py:168).
__device__ __forceinline__ void logistic_terms(double f, double& v, double& om_minus_p, double& p_out,
                                               double& e_out) {
    double e = exp(-fabs(f));
    double q = 1.0 / (1.0 + e);
    double eq = e * q;
    bool pos = f >= 0.0;
    double p = pos ? q : eq;
    double om = pos ? eq : q;     // 1 - p
    v = eq * q;                   // p (1-p)
    om_minus_p = om - p;          // 1 - 2p
    p_out = p;
    e_out = e;
}

// MODE 0: G only (position fixed-point iterates); 1: closing build (G, gradient, log-likelihood,
// cbuf); 2: gradient and log-likelihood only (Euclidean HMC, hmc.py:52-53,60-61,65-66).
template <int NT, int MODE>
__global__ void __launch_bounds__(kMetricWarps * 32, 1) k_metric(MetricArgs a) {
    constexpr bool CLOSING = MODE >= 1, WITH_G = MODE <= 1, WITH_C = MODE == 1;
    constexpr int MC = kMetricChains, NB = kMetricRows, VS = kMetricVS, ST = kMetricStages;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int xs = a.xs;
    double* xs_ring = reinterpret_cast<double*>(smem_raw);
    double* th = xs_ring + (size_t)ST * NB * xs;
    double* vs = th + (size_t)MC * xs;
    double* rs = vs + (size_t)MC * VS;
    double* ll_s = rs + (size_t)MC * VS;
    uint64_t* full = reinterpret_cast<uint64_t*>(ll_s + kMetricWarps * 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int chain0 = blockIdx.x * MC;
    if (chain0 >= a.n_chains) return;
    const int n_blocks = a.n_rows_pad / NB;
    const uint32_t stage_bytes = (uint32_t)(NB * xs * 8);

    if (tid == 0) {
        for (int s = 0; s < ST; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    // Theta tile, zero padded (pad columns multiply the staged label column by zero)
    for (int i = tid; i < MC * xs; i += blockDim.x) {
        int m = i / xs, d = i - m * xs, c = chain0 + m;
        th[i] = (c < a.n_chains && d < a.dim) ? a.theta[(size_t)c * a.dim + d] : 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < ST && s < n_blocks; ++s) {
            mbar_expect_tx(&full[s], stage_bytes);
            tma_bulk_g2s(xs_ring + (size_t)s * NB * xs, a.x + (size_t)s * NB * xs, stage_bytes, &full[s]);
        }
    }

    // columns owned by this warp: n-tiles warp, warp+8, ...
    const int n_tiles = a.p2p / 8;
    int col_a[NT], col_b[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        int nt = warp + j * kMetricWarps;
        uchar2 ab = make_uchar2(0, 0);
        if (nt < n_tiles) ab = a.pair_tab[nt * 8 + g];
        col_a[j] = ab.x; col_b[j] = ab.y;
    }
    double acc[NT][4][2];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int m = 0; m < 4; ++m) acc[j][m][0] = acc[j][m][1] = 0.0;
    // closing: gradient tiles (mt, dt) = flattened index warp, warp+8 over 4 x ceil(D/8)
    const int d_tiles = (a.dim + 7) / 8;
    double gacc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    double ll_acc = 0.0;

    const int mt1 = warp & 3;                // chain tile of this warp's f tiles
    const int k_steps_f = (a.dim + 3) / 4;
    const int tcol = xs - 1;                 // label column

    for (int rb = 0; rb < n_blocks; ++rb) {
        const int stage = rb % ST;
        const uint32_t parity = (uint32_t)((rb / ST) & 1);
        mbar_wait(&full[stage], parity);
        const double* xb = xs_ring + (size_t)stage * NB * xs;

        // ---- phase 1: f^T tiles and the logistic terms
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int rt = (warp >> 2) + 2 * h;          // row tile 0..3
            double f0 = 0.0, f1 = 0.0;
            const double* ta = th + (size_t)(mt1 * 8 + g) * xs + q;
            const double* xb_ = xb + (size_t)(rt * 8 + g) * xs + q;
            for (int ks = 0; ks < k_steps_f; ++ks) dmma884(f0, f1, ta[ks * 4], xb_[ks * 4]);
            const int r_local = rt * 8 + 2 * q;
            double vv[2], rr[2], cc[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                double f = j ? f1 : f0;
                double v, omp, p, e;
                logistic_terms(f, v, omp, p, e);
                vv[j] = v;
                if (CLOSING) {
                    double t = xb[(size_t)(r_local + j) * xs + tcol];
                    bool ovf = f > 709.782712893384;
                    rr[j] = ovf ? __longlong_as_double(0x7ff8000000000000LL) : t - p;
                    cc[j] = WITH_C ? v * omp : 0.0;
                    int row = rb * NB + r_local + j;
                    if (row < a.n_rows) {
                        double l1pe = ovf ? __longlong_as_double(0x7ff0000000000000LL) : fmax(f, 0.0) + log1p(e);
                        ll_acc += t * f - l1pe;
                    }
                }
            }
            const int m_local = mt1 * 8 + g;
            if (WITH_G)
                *reinterpret_cast<double2*>(vs + (size_t)m_local * VS + r_local) = make_double2(vv[0], vv[1]);
            if (CLOSING)
                *reinterpret_cast<double2*>(rs + (size_t)m_local * VS + r_local) = make_double2(rr[0], rr[1]);
            if (WITH_C) {
                int c = chain0 + m_local;
                if (c < a.n_chains)
                    *reinterpret_cast<double2*>(a.cbuf + (size_t)c * a.n_rows_pad + rb * NB + r_local) =
                        make_double2(cc[0], cc[1]);
            }
        }
        __syncthreads();

        // ---- phase 2: G += V . KR2(X) (and X^T r) over the 32 staged rows
#pragma unroll 2
        for (int ks = 0; ks < NB / 4; ++ks) {
            const double* xr = xb + (size_t)(ks * 4 + q) * xs;
            if (WITH_G) {
                double af[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) af[m] = vs[(size_t)(m * 8 + g) * VS + ks * 4 + q];
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    double b = xr[col_a[j]] * xr[col_b[j]];
#pragma unroll
                    for (int m = 0; m < 4; ++m) dmma884(acc[j][m][0], acc[j][m][1], af[m], b);
                }
            }
            if (CLOSING) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int tix = warp + h * kMetricWarps;
                    int mt = tix & 3, dt = tix >> 2;
                    if (dt < d_tiles) {
                        double ar = rs[(size_t)(mt * 8 + g) * VS + ks * 4 + q];
                        int dcol = dt * 8 + g;
                        double b = dcol < a.dim ? xr[dcol] : 0.0;
                        dmma884(gacc[h][0], gacc[h][1], ar, b);
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0 && rb + ST < n_blocks) {
            mbar_expect_tx(&full[stage], stage_bytes);
            tma_bulk_g2s(xs_ring + (size_t)stage * NB * xs, a.x + (size_t)(rb + ST) * NB * xs, stage_bytes,
                         &full[stage]);
        }
    }

    // ---- epilogue: packed G (+ I/alpha on the diagonal pairs)
#pragma unroll
    for (int j = 0; j < (WITH_G ? NT : 0); ++j) {
        int nt = warp + j * kMetricWarps;
        if (nt >= n_tiles) continue;
        int col = nt * 8 + 2 * q;
        uchar2 ab0 = a.pair_tab[col], ab1 = a.pair_tab[col + 1];
        double d0 = (col < a.p2 && ab0.x == ab0.y) ? a.alpha_inv : 0.0;
        double d1 = (col + 1 < a.p2 && ab1.x == ab1.y) ? a.alpha_inv : 0.0;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            int c = chain0 + m * 8 + g;
            if (c < a.n_chains) {
                double o0 = col < a.p2 ? acc[j][m][0] + d0 : 0.0;
                double o1 = col + 1 < a.p2 ? acc[j][m][1] + d1 : 0.0;
                *reinterpret_cast<double2*>(a.g_out + (size_t)c * a.p2p + col) = make_double2(o0, o1);
            }
        }
    }
    if (CLOSING) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int tix = warp + h * kMetricWarps;
            int mt = tix & 3, dt = tix >> 2;
            int c = chain0 + mt * 8 + g;
            if (dt < d_tiles && c < a.n_chains) {
                int dcol = dt * 8 + 2 * q;
                if (dcol < a.dim) a.grad_out[(size_t)c * a.dim + dcol] = gacc[h][0];
                if (dcol + 1 < a.dim) a.grad_out[(size_t)c * a.dim + dcol + 1] = gacc[h][1];
            }
        }
        // log-likelihood: fixed-order reduction (q lanes, then the two warps sharing a chain tile)
        ll_acc += __shfl_xor_sync(0xffffffffu, ll_acc, 1);
        ll_acc += __shfl_xor_sync(0xffffffffu, ll_acc, 2);
        if (q == 0) ll_s[warp * 8 + g] = ll_acc;
        __syncthreads();
        if (tid < MC) {
            int mt = tid >> 3, gg = tid & 7, c = chain0 + tid;
            if (c < a.n_chains) a.loglik_out[c] = ll_s[mt * 8 + gg] + ll_s[(mt + 4) * 8 + gg];
        }
    }
}
#endif  // __CUDACC__

}  // namespace rmhmc
