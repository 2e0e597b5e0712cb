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
#include "chain_kernels.cuh"
#include "common.cuh"

namespace rmhmc {

constexpr int kMetricChains = 32;    // chains per CTA
constexpr int kMetricRows = 32;      // design-matrix rows per staged block
constexpr int kMetricGWarps = 8;     // warps accumulating G (and X^T r).  16 (four per SM sub-partition, 3 tiles each, 96 registers) was
                                     // measured slower twice (1.99 vs 1.78 ms per German-shaped build: spills + more A-fragment traffic); 12 (4 tiles
                                     // each, 128 registers): 1.98 ms.  Fewer, larger per-warp tiles win: A fragments are reused 5x
constexpr int kMetricFWarps = 4;     // warps producing f = X theta and the logistic terms one block ahead
constexpr int kMetricThreads = (kMetricGWarps + kMetricFWarps) * 32;
constexpr int kMetricStages = 4;     // X ring depth
constexpr int kMetricVS = kMetricRows + 4;   // smem stride of the V/R tiles (4*odd)

struct MetricArgs {
    const double* x;          // [Np][XS]  zero-padded rows/cols, label t in column XS-1
    const uchar2* pair_tab;   // [P2p]     packed column -> (a, b), a <= b
    const double* theta;      // [C][D]    positions to evaluate at
    double* g_out;            // [C][P2p]  packed metric
    double* grad_out;         // [C][D]    (closing) X^T (t - p)
    double* loglik_out;       // [C]       (closing)
    double* cbuf;             // [C][Np]   (closing: written; MODE 3/4: read) c_n = v_n (1 - 2 p_n)
    // c_n slot selection: closing writes slot cw_cur[c] ^ cw_flip, the data passes read slot aslot[c]
    const int* cw_cur;        // [C] or null (single buffer)
    int cw_flip;
    size_t cw_slot;           // doubles between the two c_n slots
    const int* aslot;         // [C] or null
    const double* hbuf;       // [Cpad][Np] leverages (MODE 4)
    double* vout;             // [Cpad][Np] v_n = p_n (1 - p_n) (MODE 5)
    int n_chains, n_rows, n_rows_pad, dim, xs, p2, p2p;
    int extra_tile;           // n-tile split over the chain tiles of G-warps 0..3, or -1
    // row splitting (gridDim.z > 1: few chains, very many rows): split z handles a contiguous range of row blocks and
    // writes PARTIAL g / grad / loglik at offset z * split_* ; k_reduce_splits adds them in split order
    size_t split_g, split_grad, split_ll;
    int tiles_per_cta;        // packed-column tiles owned by one CTA (blockIdx.y selects the range)
    int n_main_tiles;         // tiles distributed over the G-warps (all tiles except extra_tile)
    double alpha_inv;
};

__host__ inline size_t metric_smem_bytes(int xs) {
    size_t b = 0;
    b += (size_t)kMetricStages * kMetricRows * xs * 8;  // X ring
    b += 4 * (size_t)kMetricChains * kMetricVS * 8;     // V and R tiles, double buffered
    b += (size_t)kMetricChains * xs * 8;                // Theta tile
    b += (size_t)kMetricFWarps * 32 * 8;                // loglik partials
    b += 256 * 8;                                       // exp table
    b += 256 * 8;                                       // log table (1/c_j, log c_j)
    b += 16 * 8;                                        // mbarriers
    return b;
}

// Per-chain work fused into the kernel's epilogue (the chain tile's packed metric stays in shared
// memory and is factored there by the CTA's 12 warps, one chain per warp at a time):
//   kFuseSolve  (MODE 0): Cholesky solve G(theta_w) u = p and the position fixed-point update
//                         theta_w <- theta + s eps/2 (u0 + u) (+ the ||theta|| > 10 hack on the last
//                         iterate)                                            rmhmc.py:121-130
//   kFuseFactor (MODE 1): L = chol(G), G^-1, 0.5 log|G| of the new position into the proposal slot
//                         (slot `cur` when init)                              rmhmc.py:138,171
enum { kFuseNone = 0, kFuseSolve = 1, kFuseFactor = 2 };
struct FuseArgs {
    int mode, is_last, init;
    double step_size;
    long long it_stop;
    const double* mom; const double* theta; const double* u0; double* theta_w;
    const int* dir; const int* step; const int* cur; const long long* iter; const int* nsteps;
    int* renorm_pos;
    double* lfac; double* invg; double* logdet;
    size_t slot_theta, slot_invg, slot_scalar;
};

__host__ inline size_t metric_smem_bytes(int xs, int p2p, bool fused) {
    size_t b = metric_smem_bytes(xs);
    if (fused) b += (size_t)kMetricChains * p2p * 8;     // the chain tile's packed metric / factor
    return b;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------- fast FP64 exp / reciprocal
// The F-warps share the FP64 pipe with the DMMAs, so the logistic terms are kept short:
//   exp(x), x <= 0:  x = (256 k + j) ln2/256 + r, |r| <= ln2/512;  e^x = 2^k * 2^(j/256) * P5(r)
//                    (256-entry table in shared memory, degree-5 Taylor: truncation < 1e-20, ~1 ulp)
//   1/y, y in [1,2]: rcp.approx.ftz.f64 (20-bit seed) + two Newton steps (~1 ulp)
// instead of the libdevice exp() and IEEE division (~4x fewer FP64 instructions).
__device__ __forceinline__ double exp_table_entry(int j) { return exp2((double)j * (1.0 / 256.0)); }

__device__ __forceinline__ double fast_exp_nonpos(double x, const double* __restrict__ tab) {
    const double kInv = 369.3299304675746271;           // 256 / ln 2
    const double kHi = 0.00270760617331689;              // ln2/256, upper 32 bits (n * kHi is exact)
    const double kLo = 7.453964567463233e-13;            // ln2/256 - kHi
    const double kMagic = 6755399441055744.0;            // 1.5 * 2^52
    x = x < -746.0 ? -746.0 : x;                          // e^-746 = 0 in FP64; keeps n inside int32 for any finite x (NaN passes)
    double t = fma(x, kInv, kMagic);
    int n = __double2loint(t);
    double nf = t - kMagic;
    double r = fma(nf, -kHi, x);
    r = fma(nf, -kLo, r);
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    int k = n >> 8;
    double y = p * tab[n & 255];
    // y in [1, 4); scale by 2^k through the exponent field; below the normal range flush to zero
    double out = __hiloint2double(__double2hiint(y) + (k << 20), __double2loint(y));
    return (k < -1020 || !(x == x)) ? (x == x ? 0.0 : x) : out;
}

__device__ __forceinline__ double fast_rcp_1to2(double y) {
    double q;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(y));
    double e = fma(-y, q, 1.0);
    q = fma(q, e, q);
    e = fma(-y, q, 1.0);
    q = fma(q, e, q);
    return q;
}

// log(1 + e), e in [0, 1]:  y = 1 + e = c_j (1 + r), c_j = 1 + (j + 0.5)/128 from the top 7 mantissa
// bits, |r| < 2^-8;  log y = log c_j + (r - r^2/2 + ... + r^7/7)   (truncation < 1e-20; absolute
// error ~1e-16, the same as the reference's log(1 + exp(f)) which also rounds 1 + e first).
__device__ __forceinline__ void log_table_entry(int j, double& inv_c, double& log_c) {
    double c = 1.0 + ((double)j + 0.5) * (1.0 / 128.0);
    inv_c = 1.0 / c;
    log_c = log(c);
}
__device__ __forceinline__ double fast_log1p_01(double e, const double* __restrict__ tab) {
    double y = 1.0 + e;
    int j = (__double2hiint(y) >> 13) & 127;
    if (y >= 2.0) j = 127;
    double r = fma(y, tab[2 * j], -1.0);
    double p = fma(r, 1.0 / 7.0, -1.0 / 6.0);
    p = fma(p, r, 0.2);
    p = fma(p, r, -0.25);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    p = fma(p, r, 1.0);
    return fma(p, r, tab[2 * j + 1]);
}


// MODE 0: G only (position fixed-point iterates); 1: closing build (G, gradient, log-likelihood,
// cbuf); 2: gradient and log-likelihood only (Euclidean HMC, hmc.py:52-53,60-61,65-66).
// Matrix-free partials (the D matrices dG/dw_d = X^T diag(c x_d) X are never formed):
// MODE 3: quadratic forms  out[c][d] = sum_n c_n (x_n . u_c)^2 x_nd = u^T dG_d u with u = a.theta
//         (LastTerm of rmhmc.py:105-107 / :159-161 without the 0.5)
// MODE 4: traces           out[c][d] = sum_n c_n h_n x_nd = tr(G^-1 dG_d), h = a.hbuf  (rmhmc.py:76-77,155-156)
// Both reuse the gradient contraction R . X of the closing build with a different R.
// MODE 5: v only, written to a.vout: the A operand of the plain-GEMM metric build G = V . KR2(X) used for 32 < D (one
//         CTA cannot hold all packed columns there, and the column CTAs of MODE 0 would each recompute f and v).
// MODE 6: the closing build without G: MODE 2 (gradient, log-likelihood) + c_n + v -> a.vout, again followed by the GEMM.
//
// Warp-specialised: 4 F-warps (one 8-row tile each, all 32 chains) compute f^T = Theta X^T on the
// tensor cores, the logistic terms, and publish V (and R) for row block rb+1 while the 8 G-warps
// accumulate G += V . KR2(X) for row block rb.  Everything is handed over through mbarriers
// (X ring full/empty, V double buffer full/empty); there is no CTA-wide barrier in the main loop.
// MODE >= 2 has no G accumulators and little work per row block: two CTAs per SM hide each other's latencies.
// NT = packed-column tiles (8 columns) per G-warp (tile t belongs to warp t mod GW); when the tile
// count is 1 mod 4 the last tile is split over the four chain tiles of G-warps 0..3 (a.extra_tile)
// so that all four SM sub-partitions issue the same number of DMMAs.
template <int NT, int MODE>
__global__ void __launch_bounds__(kMetricThreads, MODE >= 2 ? 2 : 1) k_metric(MetricArgs a, FuseArgs fz) {
    constexpr bool CLOSING = MODE == 1 || MODE == 2 || MODE == 6, WITH_G = MODE <= 1, WITH_C = MODE == 1 || MODE == 6;
    constexpr bool APPLY = MODE == 3 || MODE == 4, WITH_R = (MODE >= 1 && MODE <= 4) || MODE == 6, WITH_F = MODE != 4;
    constexpr bool VOUT = MODE == 5 || MODE == 6;
    constexpr int MC = kMetricChains, NB = kMetricRows, VS = kMetricVS, ST = kMetricStages;
    constexpr int GW = kMetricGWarps, FW = kMetricFWarps;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int xs = a.xs;
    double* xs_ring = reinterpret_cast<double*>(smem_raw);
    double* v_buf = xs_ring + (size_t)ST * NB * xs;           // [2][MC][VS]
    double* r_buf = v_buf + 2 * (size_t)MC * VS;              // [2][MC][VS]
    double* ll_s = r_buf + 2 * (size_t)MC * VS;               // [FW][32]
    double* exp_tab = ll_s + FW * 32;                         // [256] 2^(j/256)
    double* log_tab = exp_tab + 256;                          // [128][2] 1/c_j, log c_j
    double* th = log_tab + 256;                               // [MC][xs] Theta tile, zero padded
    uint64_t* bars = reinterpret_cast<uint64_t*>(th + (size_t)MC * xs);
    uint64_t* x_full = bars;            // [ST]
    uint64_t* x_empty = bars + ST;      // [ST]
    uint64_t* v_full = bars + 2 * ST;   // [2]
    uint64_t* v_empty = v_full + 2;     // [2]
    double* g_s = reinterpret_cast<double*>(bars + 16);   // [MC][P2p] packed metric tile (fused epilogues only)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int chain0 = blockIdx.x * MC;
    if (chain0 >= a.n_chains) return;
    const int n_blocks_all = a.n_rows_pad / NB;
    const int rb_begin = (int)((long long)n_blocks_all * blockIdx.z / gridDim.z);
    const int n_blocks = (int)((long long)n_blocks_all * (blockIdx.z + 1) / gridDim.z) - rb_begin;      // this split's row blocks
    if (gridDim.z > 1) {
        a.g_out += blockIdx.z * a.split_g; a.grad_out += blockIdx.z * a.split_grad; a.loglik_out += blockIdx.z * a.split_ll;
        if (blockIdx.z > 0) a.alpha_inv = 0.0;
    }
    const uint32_t stage_bytes = (uint32_t)(NB * xs * 8);

    if (tid == 0) {
        for (int s = 0; s < ST; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], GW + FW); }
        for (int s = 0; s < 2; ++s) { mbar_init(&v_full[s], FW); mbar_init(&v_empty[s], GW); }
        mbar_fence_init();
    }
    if (tid < 256) exp_tab[tid] = exp_table_entry(tid);
    if (CLOSING && tid < 128) log_table_entry(tid, log_tab[2 * tid], log_tab[2 * tid + 1]);
    // Theta tile, zero padded (pad columns meet the staged label column and must contribute 0)
    for (int i = tid; i < MC * xs; i += kMetricThreads) {
        int m = i / xs, d = i - m * xs, c = chain0 + m;
        th[i] = (c < a.n_chains && d < a.dim) ? a.theta[(size_t)c * a.dim + d] : 0.0;
    }
    __syncthreads();

    if (warp >= GW) {
        // =================================================================== F-warps
        const int fw = warp - GW;               // row tile of this warp inside every block
        if (fw == 0 && lane == 0) {
            for (int s = 0; s < ST && s < n_blocks; ++s) {
                mbar_expect_tx(&x_full[s], stage_bytes);
                tma_bulk_g2s(xs_ring + (size_t)s * NB * xs, a.x + (size_t)(rb_begin + s) * NB * xs, stage_bytes, &x_full[s]);
            }
        }
        const int k_steps_f = WITH_F ? (a.dim + 3) / 4 : 0;
        double ll_acc[4] = {0.0, 0.0, 0.0, 0.0};
        const int tcol = xs - 1;
        // MODE 3/4: this lane's c_n (and h_n) rows: chain m*8+g, columns rb*NB + fw*8 + 2q, +1
        const double* cw_row[4];
        const double* h_row[4];
        if (APPLY) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int c = chain0 + m * 8 + g;
                const size_t slot = (a.aslot && c < a.n_chains) ? (size_t)a.aslot[c] * a.cw_slot : 0;
                cw_row[m] = a.cbuf + slot + (size_t)c * a.n_rows_pad + fw * 8 + 2 * q;
                h_row[m] = MODE == 4 ? a.hbuf + (size_t)c * a.n_rows_pad + fw * 8 + 2 * q : nullptr;
            }
        }
        for (int rb = 0; rb < n_blocks; ++rb) {
            const int stage = rb % ST, buf = rb & 1;
            double2 cwv[4], hv[4];
            if (APPLY) {              // issued before the barrier waits: independent of the staged X block
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    cwv[m] = *reinterpret_cast<const double2*>(cw_row[m] + (size_t)(rb_begin + rb) * NB);
                    if (MODE == 4) hv[m] = *reinterpret_cast<const double2*>(h_row[m] + (size_t)(rb_begin + rb) * NB);
                }
            }
            // refill the stage freed two blocks ago (the G-warps have released it: see v_empty below)
            if (rb >= 2) mbar_wait(&v_empty[buf], (uint32_t)(((rb - 2) >> 1) & 1));
            if (fw == 0 && lane == 0 && rb >= 2 && rb + ST - 2 < n_blocks) {
                const int nb = rb + ST - 2, ns = nb % ST;
                mbar_wait(&x_empty[ns], (uint32_t)(((nb / ST) - 1) & 1));
                mbar_expect_tx(&x_full[ns], stage_bytes);
                tma_bulk_g2s(xs_ring + (size_t)ns * NB * xs, a.x + (size_t)(rb_begin + nb) * NB * xs, stage_bytes, &x_full[ns]);
            }
            mbar_wait(&x_full[stage], (uint32_t)((rb / ST) & 1));
            const double* xb = xs_ring + (size_t)stage * NB * xs;
            double f[4][2];
#pragma unroll
            for (int m = 0; m < 4; ++m) f[m][0] = f[m][1] = 0.0;
            const double* xrow = xb + (size_t)(fw * 8 + g) * xs + q;
            const double* trow = th + (size_t)g * xs + q;
            for (int ks = 0; ks < k_steps_f; ++ks) {
                double bx = xrow[ks * 4];
#pragma unroll
                for (int m = 0; m < 4; ++m) dmma884(f[m][0], f[m][1], trow[(size_t)m * 8 * xs + ks * 4], bx);
            }
            const int r_local = fw * 8 + 2 * q;
            double t0 = 0.0, t1 = 0.0;
            if (CLOSING) {
                t0 = xb[(size_t)r_local * xs + tcol];
                t1 = xb[(size_t)(r_local + 1) * xs + tcol];
            }
            double* vdst = v_buf + (size_t)buf * MC * VS;
            double* rdst = r_buf + (size_t)buf * MC * VS;
            if (APPLY) {
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    double2 rr;
                    if (MODE == 3) rr = make_double2(cwv[m].x * f[m][0] * f[m][0], cwv[m].y * f[m][1] * f[m][1]);
                    else rr = make_double2(cwv[m].x * hv[m].x, cwv[m].y * hv[m].y);
                    *reinterpret_cast<double2*>(rdst + (size_t)(m * 8 + g) * VS + r_local) = rr;
                }
            } else {
            // the 8 (chain, row) pairs of this lane, evaluated in lock-step for instruction-level parallelism
#pragma unroll
            for (int half = 0; half < 2; ++half) {
            double ev[4], eq[4], qq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) ev[i] = fast_exp_nonpos(-fabs(f[2 * half + (i >> 1)][i & 1]), exp_tab);
#pragma unroll
            for (int i = 0; i < 4; ++i) qq[i] = fast_rcp_1to2(1.0 + ev[i]);
#pragma unroll
            for (int i = 0; i < 4; ++i) eq[i] = ev[i] * qq[i];
#pragma unroll
            for (int mm = 0; mm < 2; ++mm) {
                const int m = 2 * half + mm;
                double vv[2], rr[2], cc[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int i = 2 * mm + j;
                    double fv = f[m][j];
                    bool pos = fv >= 0.0;
                    double p = pos ? qq[i] : eq[i];
                    double om = pos ? eq[i] : qq[i];          // 1 - p
                    vv[j] = eq[i] * qq[i];                    // p (1 - p)
                    if (CLOSING) {
                        double t = j ? t1 : t0;
                        bool ovf = fv > 709.782712893384;
                        rr[j] = ovf ? __longlong_as_double(0x7ff8000000000000LL) : t - p;
                        cc[j] = WITH_C ? vv[j] * (om - p) : 0.0;
                        int row = (rb_begin + rb) * NB + r_local + j;
                        if (row < a.n_rows) {
                            double l1pe = ovf ? __longlong_as_double(0x7ff0000000000000LL) : fmax(fv, 0.0) + fast_log1p_01(ev[i], log_tab);
                            ll_acc[m] += t * fv - l1pe;
                        }
                    }
                }
                const int m_local = m * 8 + g;
                if (WITH_G) *reinterpret_cast<double2*>(vdst + (size_t)m_local * VS + r_local) = make_double2(vv[0], vv[1]);
                if (VOUT && blockIdx.y == 0 && chain0 + m_local < a.n_chains)
                    *reinterpret_cast<double2*>(a.vout + (size_t)(chain0 + m_local) * a.n_rows_pad + (size_t)(rb_begin + rb) * NB + r_local) =
                        make_double2(vv[0], vv[1]);
                if (WITH_R) *reinterpret_cast<double2*>(rdst + (size_t)m_local * VS + r_local) = make_double2(rr[0], rr[1]);
                if (WITH_C && blockIdx.y == 0) {
                    int c = chain0 + m_local;
                    if (c < a.n_chains) {
                        const size_t slot = a.cw_cur ? (size_t)(a.cw_cur[c] ^ a.cw_flip) * a.cw_slot : 0;
                        *reinterpret_cast<double2*>(a.cbuf + slot + (size_t)c * a.n_rows_pad + (size_t)(rb_begin + rb) * NB + r_local) =
                            make_double2(cc[0], cc[1]);
                    }
                }
            }
            }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&v_full[buf]);
                mbar_arrive(&x_empty[stage]);
            }
        }
        if (CLOSING) {
            // log-likelihood: fixed-order reduction (q lanes, then the four F-warps)
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                double v = ll_acc[m];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (q == 0) ll_s[fw * 32 + m * 8 + g] = v;
            }
        }
    } else {
        // =================================================================== G-warps
        const int gw = warp;
        // this CTA's packed-column tiles: [tile0, tile_end), tile t of the range belongs to warp t mod GW
        const int tile0 = blockIdx.y * a.tiles_per_cta;
        const int tile_end = min(tile0 + a.tiles_per_cta, a.n_main_tiles);
        int col_a[NT], col_b[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            int nt = tile0 + gw + j * GW;
            uchar2 ab = make_uchar2(0, 0);
            if (nt < tile_end) ab = a.pair_tab[nt * 8 + g];
            col_a[j] = ab.x; col_b[j] = ab.y;
        }
        const bool has_extra = WITH_G && a.extra_tile >= 0 && gw < 4 && blockIdx.y == 0;
        int ex_a = 0, ex_b = 0;
        if (has_extra) { uchar2 ab = a.pair_tab[a.extra_tile * 8 + g]; ex_a = ab.x; ex_b = ab.y; }
        double acc[NT][4][2];
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int m = 0; m < 4; ++m) acc[j][m][0] = acc[j][m][1] = 0.0;
        double eacc[2] = {0.0, 0.0};
        // closing: gradient tiles (chain tile mt, parameter tile dt) = flattened index gw, gw + GW, ...
        // MODE 0/1: blockIdx.y (the column CTA) owns 4 parameter tiles = 16 (chain, parameter) tiles, 2 per G-warp.
        // The modes without G run ONE CTA per chain tile (f / v are formed once, not once per parameter-tile group):
        // up to 16 parameter tiles (D <= 128) = 64 tiles, 8 per G-warp.
        constexpr int GT = MODE <= 1 ? 16 / GW : 64 / GW;
        const int dt_base = MODE <= 1 ? 4 * blockIdx.y : 0;
        const int d_tiles = (a.dim + 7) / 8;
        double gacc[GT][2];
#pragma unroll
        for (int h = 0; h < GT; ++h) gacc[h][0] = gacc[h][1] = 0.0;

        for (int rb = 0; rb < n_blocks; ++rb) {
            const int stage = rb % ST, buf = rb & 1;
            mbar_wait(&x_full[stage], (uint32_t)((rb / ST) & 1));
            mbar_wait(&v_full[buf], (uint32_t)((rb >> 1) & 1));
            const double* xb = xs_ring + (size_t)stage * NB * xs;
            const double* vs = v_buf + (size_t)buf * MC * VS;
            const double* rs = r_buf + (size_t)buf * MC * VS;
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
                    if (has_extra) {
                        double b = xr[ex_a] * xr[ex_b];
                        double asel = gw == 0 ? af[0] : (gw == 1 ? af[1] : (gw == 2 ? af[2] : af[3]));
                        dmma884(eacc[0], eacc[1], asel, b);
                    }
                }
                if (WITH_R) {
#pragma unroll
                    for (int h = 0; h < GT; ++h) {
                        int tix = gw + h * GW;
                        int mt = tix & 3, dt = (tix >> 2) + dt_base;
                        if (dt < d_tiles) {
                            double ar = rs[(size_t)(mt * 8 + g) * VS + ks * 4 + q];
                            int dcol = dt * 8 + g;
                            double b = dcol < a.dim ? xr[dcol] : 0.0;
                            dmma884(gacc[h][0], gacc[h][1], ar, b);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&v_empty[buf]);
                mbar_arrive(&x_empty[stage]);
            }
        }

        // ---- epilogue: packed G (+ I/alpha on the diagonal pairs)
        auto store_tile = [&](int nt, int m, double o0, double o1) {
            int col = nt * 8 + 2 * q;
            uchar2 ab0 = a.pair_tab[col], ab1 = a.pair_tab[col + 1];
            double d0 = (col < a.p2 && ab0.x == ab0.y) ? a.alpha_inv : 0.0;
            double d1 = (col + 1 < a.p2 && ab1.x == ab1.y) ? a.alpha_inv : 0.0;
            int c = chain0 + m * 8 + g;
            double2 val = make_double2(col < a.p2 ? o0 + d0 : 0.0, col + 1 < a.p2 ? o1 + d1 : 0.0);
            if (fz.mode != kFuseNone)
                *reinterpret_cast<double2*>(g_s + (size_t)(m * 8 + g) * a.p2p + col) = val;
            else if (c < a.n_chains)
                *reinterpret_cast<double2*>(a.g_out + (size_t)c * a.p2p + col) = val;
        };
        if (WITH_G) {
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                int nt = tile0 + gw + j * GW;
                if (nt >= tile_end) continue;
#pragma unroll
                for (int m = 0; m < 4; ++m) store_tile(nt, m, acc[j][m][0], acc[j][m][1]);
            }
            if (has_extra) store_tile(a.extra_tile, gw, eacc[0], eacc[1]);
        }
        if (WITH_R) {
#pragma unroll
            for (int h = 0; h < GT; ++h) {
                int tix = gw + h * GW;
                int mt = tix & 3, dt = (tix >> 2) + dt_base;
                int c = chain0 + mt * 8 + g;
                if (dt < d_tiles && c < a.n_chains) {
                    int dcol = dt * 8 + 2 * q;
                    if (dcol < a.dim) a.grad_out[(size_t)c * a.dim + dcol] = gacc[h][0];
                    if (dcol + 1 < a.dim) a.grad_out[(size_t)c * a.dim + dcol + 1] = gacc[h][1];
                }
            }
        }
    }
    if (CLOSING || (WITH_G && fz.mode != kFuseNone)) __syncthreads();
    if (CLOSING && tid < MC && blockIdx.y == 0) {
        int c = chain0 + tid;
        if (c < a.n_chains) a.loglik_out[c] = (ll_s[tid] + ll_s[32 + tid]) + (ll_s[64 + tid] + ll_s[96 + tid]);
    }
    if (WITH_G && fz.mode != kFuseNone) {
        // ---- fused per-chain epilogue: every warp takes chains warp, warp + 12, ... of the tile
        constexpr int DMAX = NT <= 2 ? 16 : 32;          // NT <= 2  <=>  at most 16 parameters
        const int D = a.dim;
        double* m_scratch = xs_ring + (size_t)warp * a.p2p;      // X ring / V buffers are dead by now
        for (int ci = warp; ci < MC; ci += GW + FW) {
            const int c = chain0 + ci;
            if (c >= a.n_chains) continue;
            if (!fz.init && (fz.iter[c] >= fz.it_stop || fz.nsteps[c] <= 0)) continue;
            double* gp = g_s + (size_t)ci * a.p2p;
            double lrow[DMAX], dinv;
            load_packed_rows<DMAX>(gp, lrow, D, lane);
            __syncwarp();
            double logdet = chol_regs<DMAX>(lrow, D, lane, dinv);
            store_rows_packed<DMAX>(gp, lrow, D, lane);
            const bool live = lane < D;
            if (fz.mode == kFuseSolve) {
                const int cur = fz.cur[c];
                const int in_slot = fz.step[c] == 0 ? cur : 1 - cur;
                double p = live ? fz.mom[(size_t)c * D + lane] : 0.0;
                double w = live ? fz.theta[in_slot * fz.slot_theta + (size_t)c * D + lane] : 0.0;
                double u0 = live ? fz.u0[(size_t)c * D + lane] : 0.0;
                double u = chol_solve_packed<DMAX>(lrow, gp, dinv, D, lane, p);          // rmhmc.py:121
                double pw = w + (fz.dir[c] * fz.step_size / 2) * (u0 + u);                 // rmhmc.py:122
                if (fz.is_last) pw = clamp_position(pw, lane, D, &fz.renorm_pos[c]);
                if (live) fz.theta_w[(size_t)c * D + lane] = pw;
            } else {
                const int out = fz.init ? fz.cur[c] : 1 - fz.cur[c];
                if (lane == 0) fz.logdet[out * fz.slot_scalar + c] = logdet;
                double* ld = fz.lfac + out * fz.slot_invg + (size_t)c * D * D;
#pragma unroll
                for (int j = 0; j < DMAX; ++j)
                    if (j < D && live) ld[lane * D + j] = j <= lane ? lrow[j] : 0.0;
                double ig[DMAX];
                chol_inverse_packed<DMAX>(gp, m_scratch, dinv, ig, D, lane);
                double* igd = fz.invg + out * fz.slot_invg + (size_t)c * D * D;
#pragma unroll
                for (int b = 0; b < DMAX; ++b)
                    if (b < D && live) igd[lane * D + b] = ig[b];
            }
        }
    }
}

// out[i] = sum over splits z (in order) of part[z * stride + i]
__global__ void k_reduce_splits(const double* __restrict__ part, size_t stride, int n_splits, double* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = part[i];
    for (int z = 1; z < n_splits; ++z) s += part[(size_t)z * stride + i];
    out[i] = s;
}
#endif  // __CUDACC__

}  // namespace rmhmc
