// Implicit momentum half-step of the generalized leapfrog (rmhmc.py:102-110), all F fixed-point iterates in ONE
// launch (matrix-free partials, mf_kernels.cuh):
//     PM_0 = p;   PM_{k+1} = p + s eps/2 (grad - tr/2 + LastTerm(PM_k)),
//     LastTerm_d = 0.5 u^T dG_d u = 0.5 sum_n c_n (x_n . u)^2 x_nd,   u = G^-1 PM_k
// A CTA owns 32 chains for the whole loop.  Every iterate is one pass over the design matrix on the FP64
// tensor cores (DMMA.8x8x4), warp-specialised exactly like the metric build (metric_kernel.cuh):
//     F-warps:  S[chains x rows] = U . X^T  (K = D),  R = c .* S .* S          (one row block ahead)
//     G-warps:  Q[chains x D]   += R . X                                         (K = rows)
// X row blocks stream through the same bulk-TMA mbarrier ring (the ring keeps running across iterates), c_n is
// read from the slot the chain integrates in.  Between two passes the CTA's 12 warps update PM and form the next
// u = G^-1 PM (G^-1 from global memory: 5 KB per chain, L2-resident across the F iterates).  The last iterate
// also does rmhmc.py:110,113 and the first position iterate (whose metric is the one already held).
// Replaces F x { quadratic-form pass, k_mf_mom_iter }: 2F launches and 2F round trips of u / quad through HBM.
#pragma once
#include "chain_kernels.cuh"
#include "common.cuh"
#include "metric_kernel.cuh"

namespace rmhmc {

constexpr int kMomThreads = kMetricThreads;

__host__ inline size_t momfp_smem_bytes(int xs) {
    size_t b = 0;
    b += (size_t)kMetricStages * kMetricRows * xs * 8;  // X ring
    b += 2 * (size_t)kMetricChains * kMetricVS * 8;     // R tiles, double buffered
    b += (size_t)kMetricChains * xs * 8;                // U tile
    b += 4 * (size_t)kMetricChains * 32 * 8;            // p, grad - tr/2, Q, PM tiles
    b += (size_t)kMetricChains * 8;                     // s eps/2 per chain
    b += 16 * 8;                                        // mbarriers
    return b;
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(kMomThreads, 2) k_mom_fp(EngineParams P, ChainArrays S, const double* __restrict__ x, int xs) {
    constexpr int MC = kMetricChains, NB = kMetricRows, VS = kMetricVS, ST = kMetricStages;
    constexpr int GW = kMetricGWarps, FW = kMetricFWarps, NW = GW + FW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* xs_ring = reinterpret_cast<double*>(smem_raw);
    double* r_buf = xs_ring + (size_t)ST * NB * xs;           // [2][MC][VS]
    double* u_t = r_buf + 2 * (size_t)MC * VS;                // [MC][xs]  zero padded
    double* p_t = u_t + (size_t)MC * xs;                      // [MC][32]
    double* base_t = p_t + MC * 32;                           // [MC][32]  grad - tr/2
    double* q_t = base_t + MC * 32;                           // [MC][32]  u^T dG_d u
    double* pm_t = q_t + MC * 32;                             // [MC][32]
    double* h_s = pm_t + MC * 32;                             // [MC]      s eps/2, 0 for idle chains
    uint64_t* bars = reinterpret_cast<uint64_t*>(h_s + MC);
    uint64_t* x_full = bars;            // [ST]
    uint64_t* x_empty = bars + ST;      // [ST]
    uint64_t* v_full = bars + 2 * ST;   // [2]
    uint64_t* v_empty = v_full + 2;     // [2]
    __shared__ int slot_s[MC];          // slot the chain integrates in, -1 = idle

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int chain0 = blockIdx.x * MC, D = P.dim;
    if (chain0 >= P.n_chains) return;
    const int n_blocks = P.n_rows_pad / NB;
    const int n_total = n_blocks * P.n_fixed;
    const uint32_t stage_bytes = (uint32_t)(NB * xs * 8);

    if (tid == 0) {
        for (int s = 0; s < ST; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], NW); }
        for (int s = 0; s < 2; ++s) { mbar_init(&v_full[s], FW); mbar_init(&v_empty[s], GW); }
        mbar_fence_init();
    }
    if (tid < MC) {
        const int c = chain0 + tid;
        int slot = -1;
        double h = 0.0;
        if (c < P.n_chains && S.iter[c] < P.it_stop && S.nsteps[c] > 0) {
            const int cur = S.cur[c];
            slot = S.step[c] == 0 ? cur : 1 - cur;
            h = S.dir[c] * P.step_size / 2;
        }
        slot_s[tid] = slot;
        h_s[tid] = h;
    }
    __syncthreads();
    for (int i = tid; i < MC * xs; i += kMomThreads) {
        const int m = i / xs, d = i - m * xs, c = chain0 + m;
        u_t[i] = (slot_s[m] >= 0 && d < D) ? S.uvec[(size_t)c * D + d] : 0.0;
    }
    for (int i = tid; i < MC * 32; i += kMomThreads) {
        const int m = i >> 5, d = i & 31, c = chain0 + m, slot = slot_s[m];
        double pv = 0.0, bv = 0.0;
        if (slot >= 0 && d < D) {
            const size_t so = slot * P.slot_theta + (size_t)c * D + d;
            pv = S.mom[(size_t)c * D + d];
            bv = S.grad[so] - 0.5 * S.trace[so];
        }
        p_t[i] = pv;
        base_t[i] = bv;
    }
    __syncthreads();

    // ---- per-lane constants of the two roles
    const int fw = warp - GW;
    const double* cw_row[4] = {nullptr, nullptr, nullptr, nullptr};
    if (warp >= GW) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int ml = m * 8 + g, c = chain0 + ml;
            const size_t slot = slot_s[ml] > 0 ? P.slot_cw : 0;
            cw_row[m] = S.cw + slot + (size_t)c * P.n_rows_pad + fw * 8 + 2 * q;      // rows beyond n_chains: zero padding of cw
        }
        if (fw == 0 && lane == 0) {
            for (int s = 0; s < ST && s < n_total; ++s) {
                mbar_expect_tx(&x_full[s], stage_bytes);
                tma_bulk_g2s(xs_ring + (size_t)s * NB * xs, x + (size_t)(s % n_blocks) * NB * xs, stage_bytes, &x_full[s]);
            }
        }
    }
    const int k_steps_f = (D + 3) / 4;
    const int d_tiles = (D + 7) / 8;
    constexpr int GT = 16 / GW;

    for (int fi = 0; fi < P.n_fixed; ++fi) {
        if (warp >= GW) {
            // =============================================================== F-warps
            for (int rb = 0; rb < n_blocks; ++rb) {
                const int gb = fi * n_blocks + rb;
                const int stage = gb % ST, buf = gb & 1;
                double2 cwv[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) cwv[m] = *reinterpret_cast<const double2*>(cw_row[m] + rb * NB);
                if (gb >= 2) mbar_wait(&v_empty[buf], (uint32_t)(((gb - 2) >> 1) & 1));
                if (fw == 0 && lane == 0 && gb >= 2 && gb + ST - 2 < n_total) {
                    const int nb = gb + ST - 2, ns = nb % ST;
                    mbar_wait(&x_empty[ns], (uint32_t)(((nb / ST) - 1) & 1));
                    mbar_expect_tx(&x_full[ns], stage_bytes);
                    tma_bulk_g2s(xs_ring + (size_t)ns * NB * xs, x + (size_t)(nb % n_blocks) * NB * xs, stage_bytes, &x_full[ns]);
                }
                mbar_wait(&x_full[stage], (uint32_t)((gb / ST) & 1));
                const double* xb = xs_ring + (size_t)stage * NB * xs;
                double f[4][2];
#pragma unroll
                for (int m = 0; m < 4; ++m) f[m][0] = f[m][1] = 0.0;
                const double* xrow = xb + (size_t)(fw * 8 + g) * xs + q;
                const double* trow = u_t + (size_t)g * xs + q;
                for (int ks = 0; ks < k_steps_f; ++ks) {
                    const double bx = xrow[ks * 4];
#pragma unroll
                    for (int m = 0; m < 4; ++m) dmma884(f[m][0], f[m][1], trow[(size_t)m * 8 * xs + ks * 4], bx);
                }
                const int r_local = fw * 8 + 2 * q;
                double* rdst = r_buf + (size_t)buf * MC * VS;
#pragma unroll
                for (int m = 0; m < 4; ++m)
                    *reinterpret_cast<double2*>(rdst + (size_t)(m * 8 + g) * VS + r_local) =
                        make_double2(cwv[m].x * f[m][0] * f[m][0], cwv[m].y * f[m][1] * f[m][1]);
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&v_full[buf]);
                    mbar_arrive(&x_empty[stage]);
                }
            }
        } else {
            // =============================================================== G-warps
            const int gw = warp;
            double gacc[GT][2];
#pragma unroll
            for (int h = 0; h < GT; ++h) gacc[h][0] = gacc[h][1] = 0.0;
            for (int rb = 0; rb < n_blocks; ++rb) {
                const int gb = fi * n_blocks + rb;
                const int stage = gb % ST, buf = gb & 1;
                mbar_wait(&x_full[stage], (uint32_t)((gb / ST) & 1));
                mbar_wait(&v_full[buf], (uint32_t)((gb >> 1) & 1));
                const double* xb = xs_ring + (size_t)stage * NB * xs;
                const double* rs = r_buf + (size_t)buf * MC * VS;
#pragma unroll 2
                for (int ks = 0; ks < NB / 4; ++ks) {
                    const double* xr = xb + (size_t)(ks * 4 + q) * xs;
#pragma unroll
                    for (int h = 0; h < GT; ++h) {
                        const int tix = gw + h * GW;
                        const int mt = tix & 3, dt = tix >> 2;
                        if (dt < d_tiles) {
                            const double ar = rs[(size_t)(mt * 8 + g) * VS + ks * 4 + q];
                            const int dcol = dt * 8 + g;
                            const double b = dcol < D ? xr[dcol] : 0.0;
                            dmma884(gacc[h][0], gacc[h][1], ar, b);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&v_empty[buf]);
                    mbar_arrive(&x_empty[stage]);
                }
            }
#pragma unroll
            for (int h = 0; h < GT; ++h) {
                const int tix = gw + h * GW;
                const int mt = tix & 3, dt = tix >> 2;
                *reinterpret_cast<double2*>(q_t + (mt * 8 + g) * 32 + dt * 8 + 2 * q) = make_double2(gacc[h][0], gacc[h][1]);
            }
        }
        __syncthreads();
        // =================================================================== per-chain update, every warp
        const bool last = fi + 1 == P.n_fixed;
        for (int ci = warp; ci < MC; ci += NW) {
            const int slot = slot_s[ci];
            if (slot < 0) continue;
            const int c = chain0 + ci;
            const double h = h_s[ci];
            double pm = 0.0;
            if (lane < D) pm = p_t[ci * 32 + lane] + h * (base_t[ci * 32 + lane] + 0.5 * q_t[ci * 32 + lane]);   // rmhmc.py:108
            pm_t[ci * 32 + lane] = pm;
            __syncwarp();
            double y0 = 0.0, y1 = 0.0;
            if (lane < D) {
                const double* col = S.invg + slot * P.slot_invg + (size_t)c * D * D + lane;
                const double* xv = pm_t + ci * 32;
                int b = 0;
#pragma unroll 4
                for (; b + 1 < D; b += 2) {
                    y0 = fma(col[(size_t)b * D], xv[b], y0);
                    y1 = fma(col[(size_t)(b + 1) * D], xv[b + 1], y1);
                }
                if (b < D) y0 = fma(col[(size_t)b * D], xv[b], y0);
            }
            const double u = y0 + y1;                                                    // G^-1 PM, rmhmc.py:103
            if (lane < D) {
                u_t[(size_t)ci * xs + lane] = u;
                if (last) {
                    const size_t cd = (size_t)c * D + lane;
                    S.mom[cd] = pm;                                                      // rmhmc.py:110
                    S.u0[cd] = u;                                                        // rmhmc.py:113
                    S.theta_w[cd] = S.theta[slot * P.slot_theta + cd] + h * (u + u);     // rmhmc.py:116-122, first iterate
                }
            }
        }
        __syncthreads();
    }
}
#endif  // __CUDACC__

}  // namespace rmhmc
