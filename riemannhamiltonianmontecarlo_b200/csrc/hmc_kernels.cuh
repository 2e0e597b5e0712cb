// Euclidean HMC (hmc.py:38-89), per-chain stages.  Identity mass matrix, Stoermer-Verlet steps.
// Same asynchronous-rounds engine as RMHMC: one round = one leapfrog step (hmc.py:51-62) for every
// chain.  The gradient X^T(t - sigma(X w)) comes from the metric kernel in its gradient-only mode;
// the second gradient of a step is the first gradient of the next one, so it is evaluated once.
#pragma once
#include "chain_kernels.cuh"

namespace rmhmc {

#ifdef __CUDACC__
// stage 1: [new iteration: p = z, H_current] ; p += eps/2 grad(w) ; w += eps p
__global__ void __launch_bounds__(32) k_hmc_front(EngineParams P, ChainArrays S) {
    const int c = blockIdx.x, lane = threadIdx.x, D = P.dim;
    if (c >= P.n_chains) return;
    long long it = S.iter[c];
    if (it >= P.it_stop) return;
    const bool live = lane < D;
    const int cur = S.cur[c];
    const int step = S.step[c];
    double p = 0.0;
    int nsteps;
    if (step == 0) {
        double u_step;
        if (P.rng_mode == 0) {
            size_t row = (size_t)(it - P.tape_base) * P.n_chains + c;
            if (live) p = P.tape_z[row * D + lane];                  // (z I)^T, hmc.py:41
            u_step = P.tape_u_step[row];
        } else {
            if (live) p = philox_normal(P, c, it, (uint32_t)lane);
            u_step = philox_pair(P, c, it, 32u).u0;
        }
        nsteps = (int)ceil(u_step * (double)P.n_leapfrog);           // hmc.py:48
        double hcur = -S.logjoint[cur * P.slot_scalar + c] + 0.5 * warp_sum(live ? p * p : 0.0);   // hmc.py:72
        if (lane == 0) {
            S.hcur[c] = hcur;
            S.nsteps[c] = nsteps;
            S.dir[c] = 1;
        }
        if (nsteps <= 0) {
            if (live) S.mom[(size_t)c * D + lane] = p;
            return;
        }
    } else {
        nsteps = S.nsteps[c];
        if (live) p = S.mom[(size_t)c * D + lane];
    }
    const int in_slot = step == 0 ? cur : 1 - cur;
    double w = 0.0, grad = 0.0;
    if (live) {
        w = S.theta[in_slot * P.slot_theta + (size_t)c * D + lane];
        grad = S.grad[in_slot * P.slot_theta + (size_t)c * D + lane];
    }
    p += P.step_size / 2 * grad;                                     // hmc.py:52-54
    // hmc.py:56-57: a NaN momentum ends the trajectory before the position moves
    bool broken = __any_sync(0xffffffffu, live && isnan(p));
    if (broken) {
        if (lane == 0) S.dir[c] = -1;                                // "broken trajectory" flag
        if (live) {
            S.mom[(size_t)c * D + lane] = p;
            S.theta_w[(size_t)c * D + lane] = w;
        }
        return;
    }
    w += P.step_size * p;                                            // hmc.py:58 (InvMass = I)
    if (live) {
        S.mom[(size_t)c * D + lane] = p;
        S.theta_w[(size_t)c * D + lane] = w;
    }
}

// stage 2: p += eps/2 grad(w_new); end of trajectory: H, accept, store.  init != 0: fill slot cur.
__global__ void __launch_bounds__(32) k_hmc_back(EngineParams P, ChainArrays S, int init) {
    const int c = blockIdx.x, lane = threadIdx.x, D = P.dim;
    if (c >= P.n_chains) return;
    long long it = S.iter[c];
    if (!init && it >= P.it_stop) return;
    const bool live = lane < D;
    const int cur = S.cur[c];
    const int nsteps = init ? 1 : S.nsteps[c];
    const int out = init ? cur : 1 - cur;
    const bool broken = !init && S.dir[c] < 0;
    int step = init ? 0 : S.step[c];
    double p = (!init && live) ? S.mom[(size_t)c * D + lane] : 0.0;
    double hprop;
    if (nsteps > 0) {
        double th = live ? S.theta_w[(size_t)c * D + lane] : 0.0;
        double grad = live ? S.grad_tmp[(size_t)c * D + lane] - th / P.alpha : 0.0;     // hmc.py:61
        double lp = live ? -0.5 * log(2.0 * 3.14159265358979323846 * P.alpha) - th * th / (2.0 * P.alpha) : 0.0;
        double ljl = S.loglik_tmp[c] + warp_sum(lp);                                    // hmc.py:64-67
        if (live) {
            S.theta[out * P.slot_theta + (size_t)c * D + lane] = th;
            S.grad[out * P.slot_theta + (size_t)c * D + lane] = grad;
        }
        if (lane == 0) S.logjoint[out * P.slot_scalar + c] = ljl;
        if (init) return;
        if (!broken) {
            p += P.step_size / 2 * grad;                                                // hmc.py:62
            if (live) S.mom[(size_t)c * D + lane] = p;
            ++step;
            if (lane == 0) ++S.leapfrogs[c];
            if (step < nsteps) {
                if (lane == 0) S.step[c] = step;
                return;
            }
        }
        hprop = -ljl + 0.5 * warp_sum(live ? p * p : 0.0);                              // hmc.py:69
    } else {
        hprop = S.hcur[c];
    }
    double ratio = S.hcur[c] - hprop;                                                    // hmc.py:75
    bool take = ratio > 0.0, used_u = false;
    if (!take) {
        used_u = true;
        double ua = P.rng_mode == 0 ? P.tape_u_acc[(size_t)(it - P.tape_base) * P.n_chains + c]
                                    : philox_pair(P, c, it, 34u).u0;
        take = ratio > log(ua);
    }
    const int fin = (take && nsteps > 0) ? out : cur;
    if (P.tr_mom_end && it < P.tr_iters) {
        size_t o = ((size_t)c * P.tr_iters + it) * D + lane;
        if (live) {
            P.tr_mom_end[o] = p;
            P.tr_theta_end[o] = S.theta[(nsteps > 0 ? out : cur) * P.slot_theta + (size_t)c * D + lane];
        }
        if (lane == 0) {
            P.tr_hcur[(size_t)c * P.tr_iters + it] = S.hcur[c];
            P.tr_hprop[(size_t)c * P.tr_iters + it] = hprop;
            P.tr_flags[(size_t)c * P.tr_iters + it] = (take ? 1 : 0) | (used_u ? 2 : 0) | (nsteps << 8);
        }
    }
    if (P.samples && it > P.burn_in && it - P.burn_in < P.sample_cap && live)              // hmc.py:83-84
        P.samples[((size_t)c * P.sample_cap + (it - P.burn_in)) * D + lane] =
            S.theta[fin * P.slot_theta + (size_t)c * D + lane];
    if (lane == 0) {
        S.cur[c] = fin;
        if (take) ++S.accepted[c];
        S.step[c] = 0;
        S.iter[c] = it + 1;
    }
}
#endif

}  // namespace rmhmc
