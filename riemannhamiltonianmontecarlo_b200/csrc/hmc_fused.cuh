// Euclidean HMC (hmc.py:38-89), many leapfrog steps per launch.
//
// The round-per-launch path (hmc_kernels.cuh: k_hmc_front | k_metric<MODE 2> | k_hmc_back) costs three launches and a
// round trip of every chain's state through HBM per leapfrog step, 150 launches per MCMC iteration at L = 100.  Here one
// WARP owns 8 chains (one DMMA m-tile) for n_rounds rounds and keeps their state -- position, momentum, gradient in the
// DMMA accumulator layout -- in registers:
//   per round:  [new iteration: p = z, RandomStep, H_current]            hmc.py:41-48,72
//               p += eps/2 grad(w);  w' = w + eps p                      hmc.py:52-58   (NaN momentum ends the trajectory, :56-57)
//               one pass over X:  f = X w' (DMMA), r = t - sigma(f), grad = X^T r (DMMA; the C fragment of f is the A fragment of r),
//                                 log-likelihood only for chains whose trajectory ends in this round
//               p += eps/2 grad(w');  end of trajectory: H, accept / reject, sample store, trace          hmc.py:60-84
// Chains free-run exactly as in the round-per-launch path (one round = one leapfrog step of every chain), the X row
// blocks stream through the same bulk-TMA / mbarrier ring as pass_kernel.cuh, and mid-trajectory state is written back at
// the end of the launch in the layout k_hmc_front / k_hmc_back expect, so the two paths can be mixed.  D <= 32.
#pragma once
#include "chain_kernels.cuh"
#include "metric_kernel.cuh"
#include "pass_kernel.cuh"

namespace rmhmc {

constexpr int kHfWarps = 8;

__host__ inline size_t hmc_fused_smem_bytes(int xs, int warps) {
    return (size_t)kPassStages * kPassRows * xs * 8 + (size_t)warps * 8 * 32 * 8 + 512 * 8 + 2 * kPassStages * 8;
}

#ifdef __CUDACC__
__device__ __forceinline__ double quad_sum(double v) {          // over the four lanes (q) that share a chain
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

template <int W, bool TAIL>
__global__ void __launch_bounds__(W * 32, 512 / (W * 32)) k_hmc_rounds(EngineParams P, ChainArrays S, const double* __restrict__ x,
                                                                        int xs, int n_rows, int n_rounds) {
    constexpr int NB = kPassRows, ST = kPassStages;
    constexpr unsigned kFullMask = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* xs_ring = reinterpret_cast<double*>(smem_raw);                    // [ST][NB][xs]
    double* scratch = xs_ring + (size_t)ST * NB * xs;                         // [W][8][32]  layout conversion of w'
    double* exp_tab = scratch + (size_t)W * 8 * 32;                           // [256]
    double* log_tab = exp_tab + 256;                                          // [128][2]
    uint64_t* x_full = reinterpret_cast<uint64_t*>(log_tab + 256);
    uint64_t* x_empty = x_full + ST;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int D = P.dim, tcol = xs - 1;
    const int chain0 = (blockIdx.x * W + warp) * 8;
    const int c = chain0 + g;
    const int n_blocks = P.n_rows_pad / NB;
    const int n_total = n_blocks * n_rounds;
    const uint32_t stage_bytes = (uint32_t)(NB * xs * 8);
    const double eps = P.step_size, alpha = P.alpha;

    for (int i = tid; i < 256; i += W * 32) exp_tab[i] = exp_table_entry(i);
    for (int i = tid; i < 128; i += W * 32) log_table_entry(i, log_tab[2 * i], log_tab[2 * i + 1]);
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

    // ---- this lane's chain: state in the accumulator layout (parameters 8 dt + 2 q + j)
    const bool valid = c < P.n_chains;
    long long it = valid ? S.iter[c] : 0;
    const bool runs = valid && it < P.it_stop;               // chains already at it_stop are left untouched
    int cur = valid ? S.cur[c] : 0, step = valid ? S.step[c] : 0, nsteps = valid ? S.nsteps[c] : 0;
    double hcur = valid ? S.hcur[c] : 0.0;
    double ljl = valid ? S.logjoint[cur * P.slot_scalar + c] : 0.0;         // log joint of the CURRENT state
    double w[4][2], p[4][2], gr[4][2];
    auto load_slot = [&](int slot, bool with_mom) {
#pragma unroll
        for (int dt = 0; dt < 4; ++dt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int d = dt * 8 + 2 * q + j;
                const bool live = valid && d < D;
                const size_t o = slot * P.slot_theta + (size_t)c * D + d;
                w[dt][j] = live ? S.theta[o] : 0.0;
                gr[dt][j] = live ? S.grad[o] : 0.0;
                p[dt][j] = (live && with_mom) ? S.mom[(size_t)c * D + d] : 0.0;
            }
    };
    load_slot(step == 0 ? cur : 1 - cur, step > 0);
    const int q_hi = 4 + (q ^ 2);                            // bank-conflict-free row pairing, see pass_kernel.cuh
    const int s_row = (g & 1) ? 4 + ((g >> 1) ^ 2) : (g >> 1);   // data row (within an 8-row group) behind S-stage column g: this
                                                                 // lane's two f values belong to rows q and q_hi, the rows it
                                                                 // feeds to the gradient contraction -- no C -> A shuffle
    double* wsm = scratch + (size_t)warp * 8 * 32;
    const double half_log = 0.5 * log(2.0 * 3.14159265358979323846 * alpha);
    // D = 8 k + 1: the last parameter goes through plain FMAs instead of a k-step and a d-tile of its own (pass_kernel.cuh)
    constexpr bool tail = TAIL;                          // the launcher passes (D & 7) == 1 && D > 8
    const int k_steps = tail ? D / 4 : (D + 3) / 4, d_tiles = tail ? D / 8 : (D + 7) / 8;

    int gb = 0;
    for (int r = 0; r < n_rounds; ++r) {
        const bool active = valid && it < P.it_stop;
        // ---- new iteration: momentum draw, RandomStep, H_current                                    hmc.py:41-48,72
        if (active && step == 0) {
            const size_t row = P.rng_mode == 0 ? (size_t)(it - P.tape_base) * P.n_chains + c : 0;
            double kin = 0.0;
#pragma unroll
            for (int dt = 0; dt < 4; ++dt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int d = dt * 8 + 2 * q + j;
                    double z = 0.0;
                    if (d < D) z = P.rng_mode == 0 ? P.tape_z[row * D + d] : philox_normal(P, c, it, (uint32_t)d);
                    p[dt][j] = z;
                    kin += z * z;
                }
            const double u_step = P.rng_mode == 0 ? P.tape_u_step[row] : philox_pair(P, c, it, 0x100u).u0;
            nsteps = (int)ceil(u_step * (double)P.n_leapfrog);
            hcur = kin;            // |p|^2 partial of this lane; reduced by all lanes below
        }
        {
            // H_current = -log joint + |p|^2 / 2 for the chains that have just drawn
            const bool drew = active && step == 0;
            const double kin = quad_sum(drew ? hcur : 0.0);
            if (drew) hcur = -ljl + 0.5 * kin;
        }
        const bool moving = active && nsteps > 0;
        // ---- first half-step and position update                                                    hmc.py:52-58
        bool bad = false;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (moving) {
                    p[dt][j] = p[dt][j] + eps / 2 * gr[dt][j];
                    bad |= isnan(p[dt][j]);
                }
            }
        {
            int b = bad ? 1 : 0;
            b |= __shfl_xor_sync(kFullMask, b, 1);
            b |= __shfl_xor_sync(kFullMask, b, 2);
            bad = b != 0;
        }
        const bool broken = moving && bad;                   // a NaN momentum ends the trajectory before the position moves
        // w' goes to shared memory (accumulator layout) and comes back as DMMA A fragments; the registers of w and of the
        // old gradient are free during the pass
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) {
            const double a0 = (moving && !broken) ? w[dt][0] + eps * p[dt][0] : w[dt][0];
            const double a1 = (moving && !broken) ? w[dt][1] + eps * p[dt][1] : w[dt][1];
            *reinterpret_cast<double2*>(wsm + g * 32 + dt * 8 + 2 * q) = make_double2(a0, a1);
        }
        __syncwarp();
        double ua[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) ua[ks] = wsm[g * 32 + ks * 4 + q];
        const double u_tail = tail ? wsm[g * 32 + D - 1] : 0.0;
        const bool ends_after = moving && !broken && step + 1 >= nsteps;      // needs the log-likelihood at w'

        // ---- one pass over the data: grad = X^T (t - sigma(X w')), log-likelihood
        double acc[4][2], ll = 0.0, acc_t = 0.0;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) acc[dt][0] = acc[dt][1] = 0.0;
        for (int rb = 0; rb < n_blocks; ++rb, ++gb) {
            const int stage = gb % ST;
            if (warp == 0 && lane == 0 && gb >= 1 && gb - 1 + ST < n_total) {
                const int nb = gb - 1 + ST, ns = nb % ST;
                mbar_wait(&x_empty[ns], (uint32_t)(((gb - 1) / ST) & 1));
                mbar_expect_tx(&x_full[ns], stage_bytes);
                tma_bulk_g2s(xs_ring + (size_t)ns * NB * xs, x + (size_t)(nb % n_blocks) * NB * xs, stage_bytes, &x_full[ns]);
            }
            __syncwarp();
            mbar_wait(&x_full[stage], (uint32_t)((gb / ST) & 1));
            const double* xb = xs_ring + (size_t)stage * NB * xs;
            double sv[4][2];
#pragma unroll
            for (int r8 = 0; r8 < 4; ++r8) sv[r8][0] = sv[r8][1] = 0.0;
            {
                const double* xrow = xb + (size_t)s_row * xs + q;
                auto s_stage = [&](auto ksc) {                       // k-step count as a compile-time constant (pass_kernel.cuh)
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
            const double* xtq = xb + (D - 1);
            if (tail) {
#pragma unroll
                for (int r8 = 0; r8 < 4; ++r8) {
                    sv[r8][0] = fma(u_tail, xtq[(size_t)(r8 * 8 + q) * xs], sv[r8][0]);
                    sv[r8][1] = fma(u_tail, xtq[(size_t)(r8 * 8 + q_hi) * xs], sv[r8][1]);
                }
            }
            double aq[4][2];
#pragma unroll
            for (int r8 = 0; r8 < 4; ++r8) {
                double rr[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const double fv = sv[r8][j];
                    const double e = fast_exp_nonpos(-fabs(fv), exp_tab);
                    const double qq = fast_rcp_1to2(1.0 + e);
                    const double pn = fv >= 0.0 ? qq : e * qq;
                    const int rloc = r8 * 8 + (j ? q_hi : q);
                    const double t = xb[(size_t)rloc * xs + tcol];
                    const bool ovf = fv > 709.782712893384;      // the reference's exp(f) overflows: NaN gradient, -inf log-likelihood
                    rr[j] = ovf ? __longlong_as_double(0x7ff8000000000000LL) : t - pn;                 // hmc.py:53,61
                    if (ends_after && rb * NB + rloc < n_rows) {
                        const double l1pe = ovf ? __longlong_as_double(0x7ff0000000000000LL) : fmax(fv, 0.0) + fast_log1p_01(e, log_tab);
                        ll += t * fv - l1pe;                                                          // hmc.py:65-66
                    }
                }
                aq[r8][0] = rr[0];
                aq[r8][1] = rr[1];
            }
            if (tail) {                                      // column D - 1 of X^T (t - p)
#pragma unroll
                for (int r8 = 0; r8 < 4; ++r8) {
                    acc_t = fma(aq[r8][0], xtq[(size_t)(r8 * 8 + q) * xs], acc_t);
                    acc_t = fma(aq[r8][1], xtq[(size_t)(r8 * 8 + q_hi) * xs], acc_t);
                }
            }
            // d-tile count as a compile-time constant; columns >= D of the last tile read what follows the row's D
            // entries (inside shared memory: the ring is followed by the per-warp scratch) and only reach accumulator
            // columns that are never used (pass_kernel.cuh)
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
                        for (int dt = 0; dt < DT; ++dt) dmma884(acc[dt][0], acc[dt][1], aq[r8][kk], b[dt]);
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
            acc_t += __shfl_xor_sync(kFullMask, acc_t, 1);
            acc_t += __shfl_xor_sync(kFullMask, acc_t, 2);
#pragma unroll
            for (int dt = 1; dt < 4; ++dt)
                if (dt == d_tiles && q == 0) acc[dt][0] = acc_t;
        }
        // ---- gradient of the log joint, log joint at w'                                              hmc.py:60-67
        double lp = 0.0;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) {
            const double2 wv = *reinterpret_cast<const double2*>(wsm + g * 32 + dt * 8 + 2 * q);
            w[dt][0] = wv.x;
            w[dt][1] = wv.y;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int d = dt * 8 + 2 * q + j;
                gr[dt][j] = d < D ? acc[dt][j] - w[dt][j] / alpha : 0.0;       // chains that did not move reload their state below
                if (d < D) lp += -half_log - w[dt][j] * w[dt][j] / (2.0 * alpha);
            }
        }
        __syncwarp();
        const double ljl_new = quad_sum(ends_after ? ll : 0.0) + quad_sum(ends_after ? lp : 0.0);
        if (moving && !broken) {
#pragma unroll
            for (int dt = 0; dt < 4; ++dt)
#pragma unroll
                for (int j = 0; j < 2; ++j) p[dt][j] = p[dt][j] + eps / 2 * gr[dt][j];                  // hmc.py:62
            ++step;
            if (q == 0) ++S.leapfrogs[c];
        }
        // ---- end of the trajectory: Hamiltonian, accept / reject, store                              hmc.py:64-84
        const bool ends = active && (nsteps <= 0 || broken || step >= nsteps);
        double kin = 0.0;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) kin += p[dt][0] * p[dt][0] + p[dt][1] * p[dt][1];
        kin = quad_sum(ends ? kin : 0.0);
        if (ends) {
            const bool moved = nsteps > 0;
            const double hprop = !moved ? hcur : -(broken ? ljl : ljl_new) + 0.5 * kin;               // hmc.py:69
            const double ratio = hcur - hprop;                                                        // hmc.py:75
            bool take = ratio > 0.0, used_u = false;
            if (!take) {
                used_u = true;
                const double ua_ = P.rng_mode == 0 ? P.tape_u_acc[(size_t)(it - P.tape_base) * P.n_chains + c]
                                                   : philox_pair(P, c, it, 0x102u).u0;
                take = ratio > log(ua_);
            }
            const bool fin_new = take && moved;
            const int out = 1 - cur;
            const bool tracing = P.tr_mom_end && it < P.tr_iters;
            const bool storing = P.samples && it > P.burn_in && it - P.burn_in < P.sample_cap;         // hmc.py:83-84
#pragma unroll
            for (int dt = 0; dt < 4; ++dt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int d = dt * 8 + 2 * q + j;
                    if (d >= D) continue;
                    const size_t cd = (size_t)c * D + d;
                    const double w_cur = S.theta[cur * P.slot_theta + cd];
                    if (tracing) {
                        const size_t o = ((size_t)c * P.tr_iters + it) * D + d;
                        P.tr_mom_end[o] = p[dt][j];
                        P.tr_theta_end[o] = moved ? w[dt][j] : w_cur;
                    }
                    if (storing) P.samples[((size_t)c * P.sample_cap + (it - P.burn_in)) * D + d] = fin_new ? w[dt][j] : w_cur;
                    if (fin_new) {
                        S.theta[out * P.slot_theta + cd] = w[dt][j];
                        S.grad[out * P.slot_theta + cd] = gr[dt][j];
                    } else {                                   // back to the current state for the next iteration
                        w[dt][j] = w_cur;
                        gr[dt][j] = S.grad[cur * P.slot_theta + cd];
                    }
                }
            if (q == 0) {
                if (tracing) {
                    P.tr_hcur[(size_t)c * P.tr_iters + it] = hcur;
                    P.tr_hprop[(size_t)c * P.tr_iters + it] = hprop;
                    P.tr_flags[(size_t)c * P.tr_iters + it] = (take ? 1 : 0) | (used_u ? 2 : 0) | (nsteps << 8);
                }
                if (fin_new) S.logjoint[out * P.slot_scalar + c] = ljl_new;
                if (take) ++S.accepted[c];
            }
            if (fin_new) { cur = out; ljl = ljl_new; }
            step = 0;
            it += 1;
        }
    }

    // ---- write the chain state back in the layout of the round-per-launch kernels
    if (runs) {
        if (q == 0) {
            S.iter[c] = it;
            S.cur[c] = cur;
            S.step[c] = step;
            S.nsteps[c] = nsteps;
            S.hcur[c] = hcur;
            S.dir[c] = 1;
        }
        if (step > 0) {
#pragma unroll
            for (int dt = 0; dt < 4; ++dt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int d = dt * 8 + 2 * q + j;
                    if (d >= D) continue;
                    const size_t cd = (size_t)c * D + d;
                    S.theta[(1 - cur) * P.slot_theta + cd] = w[dt][j];
                    S.grad[(1 - cur) * P.slot_theta + cd] = gr[dt][j];
                    S.mom[cd] = p[dt][j];
                }
        }
    }
}
#endif  // __CUDACC__

}  // namespace rmhmc
