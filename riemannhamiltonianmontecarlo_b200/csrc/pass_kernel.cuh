// Passes over the design matrix of the matrix-free partials (mf_kernels.cuh), D <= 32:
//   QUAD    quad[c][d]  = sum_n c_n (x_n . u_c)^2 x_nd  = u^T dG_d u         (rmhmc.py:105-107, :159-161)
//   TRACE   trace[c][d] = sum_n c_n h_n x_nd            = tr(G^-1 dG_d)      (rmhmc.py:76-77, :155-156)
//   PAIR    both in one pass (the closing half of a leapfrog step)
//   MOMFP   the whole implicit momentum half-step (rmhmc.py:102-110): F fixed-point iterates, each a QUAD pass
//           followed by PM <- p + s eps/2 (grad - tr/2 + quad/2), u <- G^-1 PM; the last iterate also does
//           rmhmc.py:110,113 and the first position iterate (whose metric is the one already held)
// on the FP64 tensor cores (DMMA.8x8x4).  One WARP owns 8 chains (one DMMA m-tile) for the whole kernel and does
// both contractions of a pass itself:
//   S[8 chains x 8 rows] = U . X^T (K = D; U fragments live in registers), R = c .* S .* S (or c .* h),
//   R's C-fragment used directly as A-fragment (the S stage's columns are fed the rows in the order that makes the two
//   layouts coincide and keeps both stages free of bank conflicts),  Q[8 chains x D] += R . X (K = 8 rows)
// so there is no hand-off between warps: the only shared object is the X row-block ring (bulk TMA + full/empty
// mbarriers, one wait and one arrive per warp per 32 rows).  The warp-specialised formulation these replace
// (k_metric MODE 3/4, k_mom_fp: F-warps -> shared R tile -> G-warps) was issue-bound on its barrier traffic
// (ncu: 2.06 G instructions, 3.0 ms for the six passes of one momentum half-step at 65 536 chains).
#pragma once
#include <type_traits>

#include "chain_kernels.cuh"
#include "common.cuh"

namespace rmhmc {

constexpr int kPassWarps = 8;          // m-tiles (8 chains each) per CTA; kPassWarpsSmall when the grid would not fill the GPU
constexpr int kPassWarpsSmall = 2;
constexpr int kPassRows = 32;          // rows per staged X block
constexpr int kPassStages = 4;
enum { kPassQuad = 0, kPassTrace = 1, kPassPair = 2, kPassMomFp = 3 };

// per-warp scratch tiles of [8 chains][32]: PM and u; the 8-warp momentum fixed point also parks p and grad - tr/2 there
__host__ __device__ constexpr int pass_scratch_tiles(int kind, int warps) { return (kind == kPassMomFp && warps >= 4) ? 4 : 2; }
__host__ inline size_t pass_smem_bytes(int xs, int warps, int kind) {
    return (size_t)kPassStages * kPassRows * xs * 8 + (size_t)warps * pass_scratch_tiles(kind, warps) * 8 * 32 * 8 + 2 * kPassStages * 8;
}

#ifdef __CUDACC__
template <int KIND, int W, bool TAIL>
__global__ void __launch_bounds__(W * 32, 512 / (W * 32)) k_pass(EngineParams P, ChainArrays S, const double* __restrict__ x, int xs) {
    constexpr int NB = kPassRows, ST = kPassStages;
    constexpr bool WITH_U = KIND != kPassTrace, WITH_H = KIND == kPassTrace || KIND == kPassPair;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* xs_ring = reinterpret_cast<double*>(smem_raw);                    // [ST][NB][xs]
    constexpr int SCR = pass_scratch_tiles(KIND, W);
    constexpr bool PARK = SCR == 4;           // p and grad - tr/2 in shared memory (the 2-warp CTAs keep them in registers:
                                              // their residency, 6 CTAs per SM, leaves no shared memory for it)
    double* scratch = xs_ring + (size_t)ST * NB * xs;                         // [W][SCR][8][32]  PM, u (, p, grad - tr/2) of the warp's chains
    uint64_t* x_full = reinterpret_cast<uint64_t*>(scratch + (size_t)W * SCR * 8 * 32);
    uint64_t* x_empty = x_full + ST;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int D = P.dim;
    const int chain0 = (blockIdx.x * W + warp) * 8;          // first chain of this warp
    const int n_blocks = P.n_rows_pad / NB;
    const int n_iter = KIND == kPassMomFp ? P.n_fixed : 1;
    const int n_total = n_blocks * n_iter;
    const uint32_t stage_bytes = (uint32_t)(NB * xs * 8);

    if (tid == 0) {
        for (int s = 0; s < ST; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], W); }
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < ST && s < n_total; ++s) {
            mbar_expect_tx(&x_full[s], stage_bytes);
            tma_bulk_g2s(xs_ring + (size_t)s * NB * xs, x + (size_t)(s % n_blocks) * NB * xs, stage_bytes, &x_full[s]);
        }
    }

    // ---- this lane's chain (m-tile row g): slot, step direction, operand pointers
    const int c = chain0 + g;
    int slot = -1;                 // slot whose c_n this pass reads; -1: idle chain (contributes zeros, stores nothing)
    double hstep = 0.0;
    if (c < P.n_chains) {
        if (KIND == kPassMomFp) {
            if (S.iter[c] < P.it_stop && S.nsteps[c] > 0) {
                const int cur = S.cur[c];
                slot = S.step[c] == 0 ? cur : 1 - cur;
                hstep = S.dir[c] * P.step_size / 2;
            }
        } else {
            slot = S.aslot[c];
        }
    }
    const bool active = slot >= 0;
    // rows beyond n_chains exist in the c_n / leverage buffers (chain padding), so the loads below stay in bounds
    // The S stage's DMMA columns are fed the data rows of an 8-row group in an order that makes this lane's two S values
    // (columns 2q, 2q + 1) belong to exactly the two rows it supplies as A fragments to the two 4-row k-steps of the Q
    // stage -- no C -> A shuffle.
    // Which rows: column 2q carries row q, column 2q + 1 row q_hi = 4 + (q ^ 2).  With the row stride 4 * odd, rows r and
    // r + 4 of a group share their shared-memory banks; this pairing keeps the four rows of every half-warp distinct
    // mod 4 in BOTH stages (S stage: lanes g = 0..3 read rows 0, 6, 1, 7, lanes g = 4..7 rows 2, 4, 3, 5; Q stage:
    // lanes q = 0..3 read rows 0..3, then 6, 7, 4, 5), so neither stage has bank conflicts (the plain q, q + 4 pairing
    // cost the S stage a 2-way conflict on every load: ncu r02 v8, 92 M excess wavefronts per launch)
    const int q_hi = 4 + (q ^ 2);
    const double* cw_row = S.cw + (slot > 0 ? P.slot_cw : 0) + (size_t)c * P.n_rows_pad;
    const double* h_row = WITH_H ? S.hbuf + (size_t)c * P.n_rows_pad : nullptr;
    double ua[8];                  // U A-fragments: u[c][4 ks + q], zero beyond D (the staged label column meets a zero)
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const int d = ks * 4 + q;
        ua[ks] = (WITH_U && active && d < D) ? S.uvec[(size_t)c * D + d] : 0.0;
    }
    // momentum fixed point: p and grad - tr/2 of this lane's (chain, parameter) pairs, in accumulator layout; parked in
    // shared memory (read once per iterate) -- the row-block loop needs the registers
    double* pm_s = scratch + (size_t)warp * SCR * 8 * 32;    // [8][32]
    double* u_s = pm_s + 8 * 32;                             // [8][32]
    double* pl_s = u_s + 8 * 32;                             // [8][32] (PARK)
    double* bl_s = pl_s + 8 * 32;                            // [8][32] (PARK)
    double pl[4][2], bl[4][2];
    if (KIND == kPassMomFp) {
#pragma unroll
        for (int dt = 0; dt < 4; ++dt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int d = dt * 8 + 2 * q + j;
                double pv = 0.0, bv = 0.0;
                if (active && d < D) {
                    const size_t so = slot * P.slot_theta + (size_t)c * D + d;
                    pv = S.mom[(size_t)c * D + d];
                    bv = S.grad[so] - 0.5 * S.trace[so];
                }
                if (PARK) { pl_s[g * 32 + d] = pv; bl_s[g * 32 + d] = bv; }
                else { pl[dt][j] = pv; bl[dt][j] = bv; }
            }
        __syncwarp();
    }
    // D = 8 k + 1 (German credit: 25): the last parameter would cost a whole k-step of the S stage and a whole d-tile of
    // the Q stage (12 of 60 DMMAs per row block for one column of 25).  It is carried by plain FMAs instead: S gets
    // u_{D-1} x_{n,D-1} added per row, and column D-1 of Q is a per-lane dot product over the lane's own rows, reduced
    // over the four lanes of a chain at the end of the pass.
    // (TAIL is a template parameter: the launcher passes (D & 7) == 1 && D > 8, and the other instantiation is exactly
    // the kernel without this path)
    constexpr bool tail = TAIL;
    const int k_steps = tail ? D / 4 : (D + 3) / 4;
    const int d_tiles = tail ? D / 8 : (D + 7) / 8;
    double u_tail = (tail && WITH_U && active) ? S.uvec[(size_t)c * D + D - 1] : 0.0;
    const int s_row = (g & 1) ? 4 + ((g >> 1) ^ 2) : (g >> 1);   // data row (within an 8-row group) behind S-stage column g

    int gb = 0;
    for (int fi = 0; fi < n_iter; ++fi) {
        double acc[4][2], acc2[4][2];                        // acc: QUAD / TRACE; acc2: QUAD of PAIR
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) acc[dt][0] = acc[dt][1] = acc2[dt][0] = acc2[dt][1] = 0.0;
        double acc_t = 0.0, acc2_t = 0.0;                    // column D - 1 of acc / acc2 when `tail`
        for (int rb = 0; rb < n_blocks; ++rb, ++gb) {
            const int stage = gb % ST;
            // c_n (and h_n) of the block's four row groups: issued before the barrier wait
            double2 cwv[4], hv[4];                           // .x: row q, .y: row q_hi of the group
            // (requesting them one block earlier was measured slower: 1.95 -> 2.09 ms at 65 536 chains, 0.39 -> 0.45 ms at 8192)
#pragma unroll
            for (int r8 = 0; r8 < 4; ++r8) {
                cwv[r8] = make_double2(cw_row[rb * NB + r8 * 8 + q], cw_row[rb * NB + r8 * 8 + q_hi]);
                if (WITH_H) hv[r8] = make_double2(h_row[rb * NB + r8 * 8 + q], h_row[rb * NB + r8 * 8 + q_hi]);
            }
            if (warp == 0 && lane == 0 && gb >= 1 && gb - 1 + ST < n_total) {
                // refill the stage of block gb-1 once every warp has released it
                const int nb = gb - 1 + ST, ns = nb % ST;
                mbar_wait(&x_empty[ns], (uint32_t)(((gb - 1) / ST) & 1));
                mbar_expect_tx(&x_full[ns], stage_bytes);
                tma_bulk_g2s(xs_ring + (size_t)ns * NB * xs, x + (size_t)(nb % n_blocks) * NB * xs, stage_bytes, &x_full[ns]);
            }
            if (KIND == kPassMomFp && rb == n_blocks - 3) {
                // the G^-1 of the warp's 8 chains is read right after this pass (u = G^-1 PM): pull it into L2 now so
                // that the dependent loads of that phase do not wait for HBM
                const int lines = (D * D * 8 + 127) / 128;
                for (int j = 0; j < 8; ++j) {
                    const int slot_j = __shfl_sync(kFull, slot, j * 4);
                    if (slot_j < 0) continue;
                    const char* gp = reinterpret_cast<const char*>(S.invg + slot_j * P.slot_invg + (size_t)(chain0 + j) * D * D);
                    for (int l = lane; l < lines; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + (size_t)l * 128));
                }
            }
            __syncwarp();
            mbar_wait(&x_full[stage], (uint32_t)((gb / ST) & 1));
            const double* xb = xs_ring + (size_t)stage * NB * xs;
            // stage 1 for the block's four 8-row groups at once: four independent DMMA chains (the FP64 DMMA has a
            // long dependent-issue latency; a single chain per warp left the pipe idle)
            double sv[4][2];
#pragma unroll
            for (int r8 = 0; r8 < 4; ++r8) sv[r8][0] = sv[r8][1] = 0.0;
            if (WITH_U) {
                const double* xrow = xb + (size_t)s_row * xs + q;
                auto s_stage = [&](auto ksc) {                       // k-step count as a compile-time constant, as in q_stage
                    constexpr int KS = decltype(ksc)::value;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        double b[4];
#pragma unroll
                        for (int r8 = 0; r8 < 4; ++r8) b[r8] = xrow[(size_t)(r8 * 8) * xs + ks * 4];
#pragma unroll
                        for (int r8 = 0; r8 < 4; ++r8) dmma884(sv[r8][0], sv[r8][1], ua[ks], b[r8]);
                    }
                };
                switch (k_steps) {
                    case 1: s_stage(std::integral_constant<int, 1>{}); break;
                    case 2: s_stage(std::integral_constant<int, 2>{}); break;
                    case 3: s_stage(std::integral_constant<int, 3>{}); break;
                    case 4: s_stage(std::integral_constant<int, 4>{}); break;
                    case 5: s_stage(std::integral_constant<int, 5>{}); break;
                    case 6: s_stage(std::integral_constant<int, 6>{}); break;
                    case 7: s_stage(std::integral_constant<int, 7>{}); break;
                    default: s_stage(std::integral_constant<int, 8>{}); break;
                }
            }
            // x_{n,D-1} of this lane's rows (q, q_hi of every group) is read twice, here and for column D - 1 of Q below:
            // keeping the eight values across the R computation costs the PAIR kind its last registers
            const double* xtq = xb + (D - 1);
            if (tail && WITH_U) {
#pragma unroll
                for (int r8 = 0; r8 < 4; ++r8) {
                    sv[r8][0] = fma(u_tail, xtq[(size_t)(r8 * 8 + q) * xs], sv[r8][0]);
                    sv[r8][1] = fma(u_tail, xtq[(size_t)(r8 * 8 + q_hi) * xs], sv[r8][1]);
                }
            }
            // R = c .* S .* S (and / or c .* h): this lane's C-fragment values (chain g; rows q, q_hi) ARE its A
            // fragments (chain g; k = q) of the two 4-row k-steps of every group
            double aq[4][2], at[4][2];
#pragma unroll
            for (int r8 = 0; r8 < 4; ++r8) {
                if (KIND != kPassTrace) {
                    aq[r8][0] = cwv[r8].x * sv[r8][0] * sv[r8][0];
                    aq[r8][1] = cwv[r8].y * sv[r8][1] * sv[r8][1];
                }
                if (WITH_H) {
                    at[r8][0] = cwv[r8].x * hv[r8].x;
                    at[r8][1] = cwv[r8].y * hv[r8].y;
                }
            }
            if (tail) {                                      // column D - 1 of Q
#pragma unroll
                for (int r8 = 0; r8 < 4; ++r8) {
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const double xv = xtq[(size_t)(r8 * 8 + (kk ? q_hi : q)) * xs];
                        acc_t = fma(KIND == kPassTrace ? at[r8][kk] : aq[r8][kk], xv, acc_t);
                        if (KIND == kPassPair) acc2_t = fma(at[r8][kk], xv, acc2_t);
                    }
                }
            }
            // stage 2: Q += R . X over the block's eight 4-row k-steps; d-tiles are the independent chains.  The tile
            // count is a compile-time constant inside q_stage (one warp-uniform switch per block): the B fragments of a
            // k-step are loaded together and nothing but DMMAs sits between them.  Columns >= D of the last tile read
            // whatever follows the row's D entries (padding, the label, the next row; the ring is followed by the
            // scratch area, so the reads stay inside shared memory): column n of the product depends on column n of B
            // alone and the accumulator columns >= D are never used.
            auto q_stage = [&](auto dtc) {
                constexpr int DT = decltype(dtc)::value;
                const double* xq[2] = {xb + (size_t)q * xs + g, xb + (size_t)q_hi * xs + g};
#pragma unroll
                for (int r8 = 0; r8 < 4; ++r8) {
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const double* xr = xq[kk] + (size_t)(r8 * 8) * xs;
                        double b[DT];
#pragma unroll
                        for (int dt = 0; dt < DT; ++dt) b[dt] = xr[dt * 8];
#pragma unroll
                        for (int dt = 0; dt < DT; ++dt) {
                            if (KIND == kPassTrace) dmma884(acc[dt][0], acc[dt][1], at[r8][kk], b[dt]);
                            else dmma884(acc[dt][0], acc[dt][1], aq[r8][kk], b[dt]);
                            if (KIND == kPassPair) dmma884(acc2[dt][0], acc2[dt][1], at[r8][kk], b[dt]);
                        }
                    }
                }
            };
            switch (d_tiles) {
                case 1: q_stage(std::integral_constant<int, 1>{}); break;
                case 2: q_stage(std::integral_constant<int, 2>{}); break;
                case 3: q_stage(std::integral_constant<int, 3>{}); break;
                default: q_stage(std::integral_constant<int, 4>{}); break;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&x_empty[stage]);
        }
        if (tail) {
            // the four lanes of a chain hold the partial sums of their rows; lane q = 0 owns column D - 1 = 8 d_tiles
            acc_t += __shfl_xor_sync(kFull, acc_t, 1);
            acc_t += __shfl_xor_sync(kFull, acc_t, 2);
            if (KIND == kPassPair) {
                acc2_t += __shfl_xor_sync(kFull, acc2_t, 1);
                acc2_t += __shfl_xor_sync(kFull, acc2_t, 2);
            }
#pragma unroll
            for (int dt = 1; dt < 4; ++dt)
                if (dt == d_tiles && q == 0) { acc[dt][0] = acc_t; acc2[dt][0] = acc2_t; }
        }

        if (KIND != kPassMomFp) {
            if (active) {
#pragma unroll
                for (int dt = 0; dt < 4; ++dt)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int d = dt * 8 + 2 * q + j;
                        if (d < D) {
                            if (KIND == kPassTrace) S.trace_tmp[(size_t)c * D + d] = acc[dt][j];
                            else S.quad_tmp[(size_t)c * D + d] = acc[dt][j];
                            if (KIND == kPassPair) S.trace_tmp[(size_t)c * D + d] = acc2[dt][j];
                        }
                    }
            }
        } else {
            // ---- PM = p + s eps/2 (grad - tr/2 + quad/2) for this lane's pairs                      rmhmc.py:108
#pragma unroll
            for (int dt = 0; dt < 4; ++dt) {
                const double2 pv = PARK ? *reinterpret_cast<const double2*>(pl_s + g * 32 + dt * 8 + 2 * q) : make_double2(pl[dt][0], pl[dt][1]);
                const double2 bv = PARK ? *reinterpret_cast<const double2*>(bl_s + g * 32 + dt * 8 + 2 * q) : make_double2(bl[dt][0], bl[dt][1]);
                *reinterpret_cast<double2*>(pm_s + g * 32 + dt * 8 + 2 * q) =
                    make_double2(pv.x + hstep * (bv.x + 0.5 * acc[dt][0]), pv.y + hstep * (bv.y + 0.5 * acc[dt][1]));
            }
            __syncwarp();
            // ---- u = G^-1 PM, one chain at a time: lane i owns u_i, G^-1 read by columns (coalesced)   rmhmc.py:103
            const bool last = fi + 1 == n_iter;
            for (int j = 0; j < 8; ++j) {
                const int cj = chain0 + j;
                const int slot_j = __shfl_sync(kFull, slot, j * 4);
                const double h_j = __shfl_sync(kFull, hstep, j * 4);
                if (slot_j < 0) continue;
                double y0 = 0.0, y1 = 0.0;
                const double* xv = pm_s + j * 32;
                if (lane < D) {
                    // (issuing all D loads of the column before the first FMA was measured slower: 8192 chains 0.39 -> 0.47 ms)
                    const double* col = S.invg + slot_j * P.slot_invg + (size_t)cj * D * D + lane;
                    int b = 0;
#pragma unroll 4
                    for (; b + 1 < D; b += 2) {
                        y0 = fma(col[(size_t)b * D], xv[b], y0);
                        y1 = fma(col[(size_t)(b + 1) * D], xv[b + 1], y1);
                    }
                    if (b < D) y0 = fma(col[(size_t)b * D], xv[b], y0);
                }
                const double u = y0 + y1;
                u_s[j * 32 + lane] = lane < D ? u : 0.0;
                if (last && lane < D) {
                    const size_t cd = (size_t)cj * D + lane;
                    S.mom[cd] = xv[lane];                                                    // rmhmc.py:110
                    S.u0[cd] = u;                                                            // rmhmc.py:113
                    S.theta_w[cd] = S.theta[slot_j * P.slot_theta + cd] + h_j * (u + u);     // rmhmc.py:116-122, first iterate
                }
            }
            __syncwarp();
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) ua[ks] = active ? u_s[g * 32 + ks * 4 + q] : 0.0;
            if (tail) u_tail = active ? u_s[g * 32 + D - 1] : 0.0;
            __syncwarp();
        }
    }
}
#endif  // __CUDACC__

}  // namespace rmhmc
