// Metric build, second formulation (D <= 25, many chains): one WARP owns 8 chains (one DMMA m-tile) and ALL packed
// columns of their metric, so nothing is handed from warp to warp:
//   f[8 chains x 32 rows] = Theta . X^T              (DMMA, K = D, Theta fragments in registers)      rmhmc.py:51,116,134
//   v = p (1 - p), p = sigma(f)                        (registers; closing: t - p, c_n, log-likelihood)  rmhmc.py:52-53
//   V's C-fragment -> A-fragment by four warp shuffles
//   G[8 chains x P2] += V . KR2(X)                     (DMMA, K = rows; TILES x 2 accumulator registers)  rmhmc.py:57
// The B operand KR2(X)[n, (a,b)] = x_na x_nb of a 32-row block is formed ONCE per CTA in shared memory (every warp
// forms 4 of the 32 rows of the next block's tile before it starts on the current one), so the DMMA stream of a warp
// contains no FP64 multiplies and one shared-memory load per DMMA.  In the first formulation (metric_kernel.cuh:
// 4 F-warps -> shared V tile -> 8 G-warps, each lane forming its own B elements) ncu attributed 26 % of the G-warps'
// time to DMULs stalled behind the other warps' DMMAs (math-pipe throttle) and 13 % to waiting for the V hand-off.
// X row blocks arrive by 1-D bulk TMA into a 3-stage mbarrier ring; there is no CTA-wide barrier in the main loop.
#pragma once
#include "common.cuh"
#include "metric_kernel.cuh"

namespace rmhmc {

constexpr int kM2Warps = 8;            // m-tiles (8 chains each) per CTA
constexpr int kM2Threads = kM2Warps * 32;
constexpr int kM2Rows = 32;            // rows per staged block
constexpr int kM2Stages = 3;           // X ring depth

__host__ __device__ inline int m2_kr_stride(int tiles) { return tiles * 8 + 4; }      // = 4 or 12 mod 16: conflict-free B fragments
__host__ inline size_t metric2_smem_bytes(int xs, int tiles) {
    return ((size_t)kM2Stages * kM2Rows * xs + 2 * (size_t)kM2Rows * m2_kr_stride(tiles) + 256 + 256) * 8 + 16 * 8;
}

#ifdef __CUDACC__
// MODE 0: G only (position fixed-point iterates); 1: closing build (G, X^T (t - p), log-likelihood, c_n).
// TILES = ceil(P2 / 8) rounded up to an instantiated value; surplus tiles see pair (0,0) and are never stored.
template <int TILES, int MODE>
__global__ void __launch_bounds__(kM2Threads, 1) k_metric2(MetricArgs a) {
    constexpr bool CLOSING = MODE == 1;
    constexpr int W = kM2Warps, NB = kM2Rows, ST = kM2Stages, RS = TILES * 8 + 4;
    constexpr int PC = (TILES * 8 + 31) / 32;        // tile columns formed per lane and row
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int xs = a.xs;
    double* xs_ring = reinterpret_cast<double*>(smem_raw);                // [ST][NB][xs]
    double* kr = xs_ring + (size_t)ST * NB * xs;                          // [2][NB][RS]
    double* exp_tab = kr + 2 * (size_t)NB * RS;                           // [256]
    double* log_tab = exp_tab + 256;                                      // [128][2]
    uint64_t* x_full = reinterpret_cast<uint64_t*>(log_tab + 256);        // [ST]
    uint64_t* x_empty = x_full + ST;                                      // [ST]
    uint64_t* kr_full = x_empty + ST;                                     // [2]
    uint64_t* kr_empty = kr_full + 2;                                     // [2]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int D = a.dim;
    const int chain0 = (blockIdx.x * W + warp) * 8;
    const int n_blocks = a.n_rows_pad / NB;
    const uint32_t stage_bytes = (uint32_t)(NB * xs * 8);

    if (tid == 0) {
        for (int s = 0; s < ST; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], W); }
        for (int s = 0; s < 2; ++s) { mbar_init(&kr_full[s], W); mbar_init(&kr_empty[s], W); }
        mbar_fence_init();
    }
    exp_tab[tid] = exp_table_entry(tid);
    if (CLOSING && tid < 128) log_table_entry(tid, log_tab[2 * tid], log_tab[2 * tid + 1]);
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < ST && s < n_blocks; ++s) {
            mbar_expect_tx(&x_full[s], stage_bytes);
            tma_bulk_g2s(xs_ring + (size_t)s * NB * xs, a.x + (size_t)s * NB * xs, stage_bytes, &x_full[s]);
        }
    }

    // ---- per-lane constants
    const int c = chain0 + g;
    const bool valid = c < a.n_chains;
    double ua[8];                 // Theta A-fragments: theta[c][4 ks + q], zero beyond D (the staged label column meets a zero)
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const int d = ks * 4 + q;
        ua[ks] = (valid && d < D) ? a.theta[(size_t)c * D + d] : 0.0;
    }
    int pcol[PC];                 // (a, b) of the tile columns this lane forms: lane, lane + 32, ...
#pragma unroll
    for (int i = 0; i < PC; ++i) {
        const int col = lane + 32 * i;
        uchar2 ab = make_uchar2(0, 0);
        if (col < a.p2p) ab = a.pair_tab[col];
        pcol[i] = ab.x | (ab.y << 8);
    }
    auto produce = [&](int rb) {          // rows 4 warp .. 4 warp + 3 of block rb's KR2 tile (its X rows must have landed)
        const double* xb = xs_ring + (size_t)(rb % ST) * NB * xs;
        double* dst = kr + (size_t)(rb & 1) * NB * RS;
#pragma unroll
        for (int rr = 0; rr < NB / W; ++rr) {
            const int r = warp * (NB / W) + rr;
            const double* xr = xb + (size_t)r * xs;
#pragma unroll
            for (int i = 0; i < PC; ++i) {
                const int col = lane + 32 * i;
                if (col < TILES * 8) dst[(size_t)r * RS + col] = xr[pcol[i] & 255] * xr[pcol[i] >> 8];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&kr_full[rb & 1]);
    };
    const int k_steps = (D + 3) / 4;
    const int d_tiles = (D + 7) / 8;
    const int src0 = g * 4 + (q >> 1), src1 = src0 + 2;      // shuffle sources of the C -> A fragment conversion
    const bool odd = q & 1;
    const int tcol = xs - 1;
    const size_t cw_off = CLOSING ? (a.cw_cur && valid ? (size_t)(a.cw_cur[c] ^ a.cw_flip) * a.cw_slot : 0) : 0;

    double acc[TILES][2];
#pragma unroll
    for (int j = 0; j < TILES; ++j) acc[j][0] = acc[j][1] = 0.0;
    double gacc[4][2];
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) gacc[dt][0] = gacc[dt][1] = 0.0;
    double ll_acc = 0.0;

    mbar_wait(&x_full[0], 0);
    produce(0);

    for (int rb = 0; rb < n_blocks; ++rb) {
        const int stage = rb % ST, buf = rb & 1;
        if (tid == 0 && rb >= 1 && rb - 1 + ST < n_blocks) {
            // refill the stage of block rb-1 once every warp has released it
            const int nb = rb - 1 + ST, ns = nb % ST;
            mbar_wait(&x_empty[ns], (uint32_t)(((rb - 1) / ST) & 1));
            mbar_expect_tx(&x_full[ns], stage_bytes);
            tma_bulk_g2s(xs_ring + (size_t)ns * NB * xs, a.x + (size_t)nb * NB * xs, stage_bytes, &x_full[ns]);
        }
        __syncwarp();
        if (rb + 1 < n_blocks) {
            // this warp's slice of the NEXT block's KR2 tile (its buffer was last read for block rb-1)
            mbar_wait(&x_full[(rb + 1) % ST], (uint32_t)(((rb + 1) / ST) & 1));
            if (rb >= 1) mbar_wait(&kr_empty[(rb + 1) & 1], (uint32_t)(((rb - 1) >> 1) & 1));
            produce(rb + 1);
        }
        const double* xb = xs_ring + (size_t)stage * NB * xs;
        // ---- stage 1: f for the block's four 8-row groups (four independent DMMA chains)
        double sv[4][2];
#pragma unroll
        for (int r8 = 0; r8 < 4; ++r8) sv[r8][0] = sv[r8][1] = 0.0;
        {
            const double* xrow = xb + (size_t)g * xs + q;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                if (ks < k_steps) {
#pragma unroll
                    for (int r8 = 0; r8 < 4; ++r8) dmma884(sv[r8][0], sv[r8][1], ua[ks], xrow[(size_t)(r8 * 8) * xs + ks * 4]);
                }
            }
        }
        // ---- logistic terms of this lane's 8 (chain, row) pairs; C fragment -> A fragments
        double av[4][2], ar[4][2];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            double ev[4], eq[4], qq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) ev[i] = fast_exp_nonpos(-fabs(sv[2 * half + (i >> 1)][i & 1]), exp_tab);
#pragma unroll
            for (int i = 0; i < 4; ++i) qq[i] = fast_rcp_1to2(1.0 + ev[i]);
#pragma unroll
            for (int i = 0; i < 4; ++i) eq[i] = ev[i] * qq[i];
#pragma unroll
            for (int mm = 0; mm < 2; ++mm) {
                const int r8 = 2 * half + mm;
                const int r_local = r8 * 8 + 2 * q;
                double vv[2], rr[2] = {0.0, 0.0}, cc[2] = {0.0, 0.0};
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int i = 2 * mm + j;
                    const double fv = sv[r8][j];
                    const bool pos = fv >= 0.0;
                    const double p = pos ? qq[i] : eq[i];
                    const double om = pos ? eq[i] : qq[i];          // 1 - p
                    vv[j] = eq[i] * qq[i];                          // p (1 - p)
                    if (CLOSING) {
                        const double t = xb[(size_t)(r_local + j) * xs + tcol];
                        const bool ovf = fv > 709.782712893384;
                        rr[j] = ovf ? __longlong_as_double(0x7ff8000000000000LL) : t - p;
                        cc[j] = vv[j] * (om - p);
                        if (rb * NB + r_local + j < a.n_rows) {
                            const double l1pe = ovf ? __longlong_as_double(0x7ff0000000000000LL)
                                                    : fmax(fv, 0.0) + fast_log1p_01(ev[i], log_tab);
                            ll_acc += t * fv - l1pe;
                        }
                    }
                }
                if (CLOSING && valid)
                    *reinterpret_cast<double2*>(a.cbuf + cw_off + (size_t)c * a.n_rows_pad + rb * NB + r_local) =
                        make_double2(cc[0], cc[1]);
                {
                    const double e0 = __shfl_sync(kFull, vv[0], src0), o0 = __shfl_sync(kFull, vv[1], src0);
                    const double e1 = __shfl_sync(kFull, vv[0], src1), o1 = __shfl_sync(kFull, vv[1], src1);
                    av[r8][0] = odd ? o0 : e0;
                    av[r8][1] = odd ? o1 : e1;
                }
                if (CLOSING) {
                    const double e0 = __shfl_sync(kFull, rr[0], src0), o0 = __shfl_sync(kFull, rr[1], src0);
                    const double e1 = __shfl_sync(kFull, rr[0], src1), o1 = __shfl_sync(kFull, rr[1], src1);
                    ar[r8][0] = odd ? o0 : e0;
                    ar[r8][1] = odd ? o1 : e1;
                }
            }
        }
        // ---- stage 2: G += V . KR2 over the block's eight 4-row k-steps
        mbar_wait(&kr_full[buf], (uint32_t)((rb >> 1) & 1));
        const double* kb = kr + (size_t)buf * NB * RS + (size_t)q * RS + g;
#pragma unroll
        for (int r8 = 0; r8 < 4; ++r8) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const double* krow = kb + (size_t)(r8 * 8 + kk * 4) * RS;
                const double afrag = av[r8][kk];
#pragma unroll
                for (int j = 0; j < TILES; ++j) dmma884(acc[j][0], acc[j][1], afrag, krow[j * 8]);
                if (CLOSING) {
                    const double* xr = xb + (size_t)(r8 * 8 + kk * 4 + q) * xs;
#pragma unroll
                    for (int dt = 0; dt < 4; ++dt) {
                        if (dt < d_tiles) {
                            const int dcol = dt * 8 + g;
                            const double b = dcol < D ? xr[dcol] : 0.0;
                            dmma884(gacc[dt][0], gacc[dt][1], ar[r8][kk], b);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&kr_empty[buf]);
            mbar_arrive(&x_empty[stage]);
        }
    }

    // ---- epilogue: packed G (+ I/alpha on the diagonal pairs), gradient, log-likelihood
    if (valid) {
#pragma unroll
        for (int j = 0; j < TILES; ++j) {
            const int col = j * 8 + 2 * q;
            if (col < a.p2p) {
                const uchar2 ab0 = a.pair_tab[col], ab1 = a.pair_tab[col + 1];
                const double d0 = (col < a.p2 && ab0.x == ab0.y) ? a.alpha_inv : 0.0;
                const double d1 = (col + 1 < a.p2 && ab1.x == ab1.y) ? a.alpha_inv : 0.0;
                *reinterpret_cast<double2*>(a.g_out + (size_t)c * a.p2p + col) =
                    make_double2(col < a.p2 ? acc[j][0] + d0 : 0.0, col + 1 < a.p2 ? acc[j][1] + d1 : 0.0);
            }
        }
        if (CLOSING) {
#pragma unroll
            for (int dt = 0; dt < 4; ++dt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int d = dt * 8 + 2 * q + j;
                    if (d < D) a.grad_out[(size_t)c * D + d] = gacc[dt][j];
                }
        }
    }
    if (CLOSING) {
        double v = ll_acc;                                   // fixed-order reduction over the four q lanes of the chain
        v += __shfl_xor_sync(kFull, v, 1);
        v += __shfl_xor_sync(kFull, v, 2);
        if (valid && q == 0) a.loglik_out[c] = v;
    }
}
#endif  // __CUDACC__

}  // namespace rmhmc
