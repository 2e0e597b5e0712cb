// Euclidean HMC (hmc.py:38-89), per-chain stages.  Identity mass matrix, Stoermer-Verlet steps.
// Same asynchronous-rounds engine as RMHMC: one round = one leapfrog step (hmc.py:51-62) for every
// chain.  The gradient X^T(t - sigma(X w)) comes from the metric kernel in its gradient-only mode;
// the second gradient of a step is the first gradient of the next one, so it is evaluated once.
#pragma once
#include "chain_kernels.cuh"

namespace rmhmc {

#ifdef __CUDACC__
// stage 1: [new iteration: p = z, H_current] ; p += eps/2 grad(w) ; w += eps p
// One warp per chain; lane l owns parameters l, l + 32, ... (D <= 128).
__global__ void __launch_bounds__(32) k_hmc_front(EngineParams P, ChainArrays S) {
    const int c = blockIdx.x, lane = threadIdx.x, D = P.dim;
    if (c >= P.n_chains) return;
    long long it = S.iter[c];
    if (it >= P.it_stop) return;
    const int cur = S.cur[c];
    const int step = S.step[c];
    double* mom = S.mom + (size_t)c * D;
    int nsteps;
    if (step == 0) {
        double u_step, kin = 0.0;
        size_t row = P.rng_mode == 0 ? (size_t)(it - P.tape_base) * P.n_chains + c : 0;
        for (int d = lane; d < D; d += 32) {
            double p = P.rng_mode == 0 ? P.tape_z[row * D + d] : philox_normal(P, c, it, (uint32_t)d);   // (z I)^T, hmc.py:41
            mom[d] = p;
            kin += p * p;
        }
        u_step = P.rng_mode == 0 ? P.tape_u_step[row] : philox_pair(P, c, it, 0x100u).u0;
        nsteps = (int)ceil(u_step * (double)P.n_leapfrog);           // hmc.py:48
        double hcur = -S.logjoint[cur * P.slot_scalar + c] + 0.5 * warp_sum(kin);   // hmc.py:72
        if (lane == 0) {
            S.hcur[c] = hcur;
            S.nsteps[c] = nsteps;
            S.dir[c] = 1;
        }
        if (nsteps <= 0) return;
    }
    const int in_slot = step == 0 ? cur : 1 - cur;
    const double* w_in = S.theta + in_slot * P.slot_theta + (size_t)c * D;
    const double* g_in = S.grad + in_slot * P.slot_theta + (size_t)c * D;
    bool bad = false;
    for (int d = lane; d < D; d += 32) {
        double p = mom[d] + P.step_size / 2 * g_in[d];               // hmc.py:52-54
        mom[d] = p;
        bad |= isnan(p);
    }
    // hmc.py:56-57: a NaN momentum ends the trajectory before the position moves
    const bool broken = __any_sync(0xffffffffu, bad);
    if (broken && lane == 0) S.dir[c] = -1;                          // "broken trajectory" flag
    for (int d = lane; d < D; d += 32)
        S.theta_w[(size_t)c * D + d] = broken ? w_in[d] : w_in[d] + P.step_size * mom[d];   // hmc.py:58 (InvMass = I)
}

// stage 2: p += eps/2 grad(w_new); end of trajectory: H, accept, store.  init != 0: fill slot cur.
__global__ void __launch_bounds__(32) k_hmc_back(EngineParams P, ChainArrays S, int init) {
    const int c = blockIdx.x, lane = threadIdx.x, D = P.dim;
    if (c >= P.n_chains) return;
    long long it = S.iter[c];
    if (!init && it >= P.it_stop) return;
    const int cur = S.cur[c];
    const int nsteps = init ? 1 : S.nsteps[c];
    const int out = init ? cur : 1 - cur;
    const bool broken = !init && S.dir[c] < 0;
    int step = init ? 0 : S.step[c];
    double* mom = S.mom + (size_t)c * D;
    double hprop;
    if (nsteps > 0) {
        double lp = 0.0;
        for (int d = lane; d < D; d += 32) {
            double th = S.theta_w[(size_t)c * D + d];
            S.theta[out * P.slot_theta + (size_t)c * D + d] = th;
            S.grad[out * P.slot_theta + (size_t)c * D + d] = S.grad_tmp[(size_t)c * D + d] - th / P.alpha;   // hmc.py:61
            lp += -0.5 * log(2.0 * 3.14159265358979323846 * P.alpha) - th * th / (2.0 * P.alpha);
        }
        double ljl = S.loglik_tmp[c] + warp_sum(lp);                                    // hmc.py:64-67
        if (lane == 0) S.logjoint[out * P.slot_scalar + c] = ljl;
        if (init) return;
        if (!broken) {
            for (int d = lane; d < D; d += 32)
                mom[d] += P.step_size / 2 * S.grad[out * P.slot_theta + (size_t)c * D + d];   // hmc.py:62
            ++step;
            if (lane == 0) ++S.leapfrogs[c];
            if (step < nsteps) {
                if (lane == 0) S.step[c] = step;
                return;
            }
        }
        double kin = 0.0;
        for (int d = lane; d < D; d += 32) kin += mom[d] * mom[d];
        hprop = -ljl + 0.5 * warp_sum(kin);                                             // hmc.py:69
    } else {
        hprop = S.hcur[c];
    }
    double ratio = S.hcur[c] - hprop;                                                    // hmc.py:75
    bool take = ratio > 0.0, used_u = false;
    if (!take) {
        used_u = true;
        double ua = P.rng_mode == 0 ? P.tape_u_acc[(size_t)(it - P.tape_base) * P.n_chains + c]
                                    : philox_pair(P, c, it, 0x102u).u0;
        take = ratio > log(ua);
    }
    const int fin = (take && nsteps > 0) ? out : cur;
    if (P.tr_mom_end && it < P.tr_iters) {
        for (int d = lane; d < D; d += 32) {
            size_t o = ((size_t)c * P.tr_iters + it) * D + d;
            P.tr_mom_end[o] = mom[d];
            P.tr_theta_end[o] = S.theta[(nsteps > 0 ? out : cur) * P.slot_theta + (size_t)c * D + d];
        }
        if (lane == 0) {
            P.tr_hcur[(size_t)c * P.tr_iters + it] = S.hcur[c];
            P.tr_hprop[(size_t)c * P.tr_iters + it] = hprop;
            P.tr_flags[(size_t)c * P.tr_iters + it] = (take ? 1 : 0) | (used_u ? 2 : 0) | (nsteps << 8);
        }
    }
    if (P.samples && it > P.burn_in && it - P.burn_in < P.sample_cap)                       // hmc.py:83-84
        for (int d = lane; d < D; d += 32)
            P.samples[((size_t)c * P.sample_cap + (it - P.burn_in)) * D + d] = S.theta[fin * P.slot_theta + (size_t)c * D + d];
    __syncwarp();
    if (lane == 0) {
        S.cur[c] = fin;
        if (take) ++S.accepted[c];
        S.step[c] = 0;
        S.iter[c] = it + 1;
    }
}
#endif

}  // namespace rmhmc
