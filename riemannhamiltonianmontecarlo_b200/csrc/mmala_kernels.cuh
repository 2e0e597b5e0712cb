// Manifold MALA and simplified manifold MALA (Girolami & Calderhead), per-chain stage.
//
// The reference repository has these samplers only as MATLAB originals:
//   code/authors_code/Bayes_Log_Reg/MCMC/BLR_mMALA.m:159-330, BLR_mMALA_Simp.m:170-290   (cited as M:line / S:line)
// One MCMC iteration = ONE evaluation of the metric quantities at the proposed point, so a "round" of the engine is
// one iteration for every chain (chains run in lock-step; there are no trajectories):
//   k_mmala_turn (front)   proposal  w' = Mean + (z chol(eps G^-1))',  Mean = w + drift(w)                M:227-233
//   k_metric<MODE 1>       G(w'), X^T (t - p), log-likelihood, c_n                                       M:236-253
//   k_chain_factor         L_G = chol(G), G^-1                                                            M:256
//   leverage GEMM, k_pass<TRACE>   tr(G^-1 dG_d)      (full mMALA only)                                   M:260-271
//   k_mmala_turn (back)    drift(w'), chol(eps G'^-1), both proposal densities, Ratio, accept, store      M:273-315
// Drift: the reference sums SecondTerm(:,d) = G^-1 dG_d G^-1 e_d over d and adds ThirdTerm = G^-1 tr (M:208-213,
// :227-229).  Because the stacked partials are a fully symmetric 3-tensor, sum_d SecondTerm(:,d) == ThirdTerm
// (checked to 6e-16 in the oracle), so
//   Mean = w + eps/2 G^-1 grad - eps sum_d SecondTerm + eps/2 ThirdTerm = w + eps/2 G^-1 (grad - tr)
// and tr_d = sum_n c_n x_nd (x_n^T G^-1 x_n) comes from the matrix-free passes (mf_kernels.cuh).  The simplified
// variant drops the two partials terms (S:207).
//
// Slot contents in this sampler (ChainArrays, two slots, accept = flip): theta, logjoint, lfac = L_G (dense lower),
// invg = R = chol(eps G^-1) as MATLAB's UPPER factor (R'R = eps G^-1, dense, zeros below the diagonal),
// logdet = sum log diag R, grad = the drift vector.
#pragma once
#include "chain_kernels.cuh"
#include "mf_kernels.cuh"

namespace rmhmc {

#ifdef __CUDACC__
// N = compile-time order of the warp Cholesky (chain_order(D)); one warp per chain, lane = parameter.
// do_back: finish the iteration whose builds at theta_w have just run (init != 0: only fill slot cur);
// do_front: propose for the next iteration.
template <int N>
__global__ void __launch_bounds__(32) k_mmala_turn(EngineParams P, ChainArrays S, int do_back, int do_front, int init,
                                                   int simplified) {
    __shared__ double xs[32], red[32], colbuf[64];
    const int c = blockIdx.x, lane = threadIdx.x, D = P.dim;
    if (c >= P.n_chains) return;
    long long it = S.iter[c];
    if (!init && it >= P.it_stop) return;
    const bool live = lane < D;
    const size_t cd = (size_t)c * D + lane;
    const double eps = P.step_size;
    // variant 2 = the IWLS proposal of code/iwls.py:28-45: N(w + G^-1 grad, G^-1) -- the simplified-mMALA proposal with
    // drift coefficient 1 and covariance G^-1 (cov . X^T W z = w + G^-1 (X^T (t - p) - w / alpha))
    const bool iwls = simplified == 2;
    const double drift_c = iwls ? 1.0 : eps / 2, cov_c = iwls ? 1.0 : eps;
    int cur = S.cur[c];

    if (do_back) {
        const int out = init ? cur : 1 - cur;
        double* invg_out = S.invg + out * P.slot_invg + (size_t)c * D * D;       // holds G^-1 (k_chain_factor), R afterwards
        const double* lg_out = S.lfac + out * P.slot_invg + (size_t)c * D * D;
        double th = 0.0, dr = 0.0;
        if (live) {
            th = S.theta_w[cd];
            dr = S.grad_tmp[cd] - th / P.alpha;                                                   // M:258 (gradient)
            if (simplified == 0) dr -= S.trace_tmp[cd];                                           // grad - tr
        }
        // drift = eps/2 G^-1 (grad - tr)                                                          M:258,:273-277
        const double drift = drift_c * mf_matvec<32>(invg_out, dr, xs, D, lane);
        // log prior, summed in parameter order (LogNormPDF.m)
        __syncwarp();
        xs[lane] = th;
        __syncwarp();
        const double half_log = 0.5 * log(2.0 * 3.14159265358979323846 * P.alpha);
        double lp = 0.0;
        for (int b = 0; b < D; ++b) lp += -half_log - xs[b] * xs[b] / (2.0 * P.alpha);
        const double ljl = S.loglik_tmp[c] + lp;                                                  // M:236-239
        // R = chol(eps G^-1), sum log diag R                                                      M:241,:279
        double row[N], dinv;
#pragma unroll
        for (int j = 0; j < N; ++j)
            row[j] = (j < D && lane >= j && live) ? cov_c * invg_out[(size_t)lane * D + j] : (j == lane ? 1.0 : 0.0);
        __syncwarp();
        double ldp;
        if (iwls) {
            // iwls.py:64,68: the density's log-determinant term is sum log diag chol(cov + 1e-6 I), the draw uses cov itself
            double row2[N], dinv2;
#pragma unroll
            for (int j = 0; j < N; ++j) row2[j] = row[j] + ((j == lane && live) ? 1e-6 : 0.0);
            ldp = chol_fixed<N>(row2, colbuf, lane, dinv2);
            __syncwarp();
            chol_fixed<N>(row, colbuf, lane, dinv);
        } else {
            ldp = chol_fixed<N>(row, colbuf, lane, dinv);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (j < D && live) invg_out[(size_t)j * D + lane] = lane >= j ? row[j] : 0.0;         // R[j][lane] = L[lane][j]
        if (live) {
            S.theta[out * P.slot_theta + cd] = th;
            S.grad[out * P.slot_theta + cd] = drift;
        }
        if (lane == 0) {
            S.logjoint[out * P.slot_scalar + c] = ljl;
            S.logdet[out * P.slot_scalar + c] = ldp;
        }
        if (init) return;

        // ProbOldGivenNew = -sum log diag R' - 0.5 (Mean' - w)^T (G'/eps) (Mean' - w), (.)^T G' (.) = |L_G'^T (.)|^2     M:279
        const double w_cur = live ? S.theta[cur * P.slot_theta + cd] : 0.0;
        const double diff = live ? (th + drift) - w_cur : 0.0;
        const double y = mf_matvec<32>(lg_out, diff, xs, D, lane);
        const double p_old_new = -ldp - 0.5 * mf_sum<32>(live ? y * y : 0.0, red, D, lane) / cov_c;
        const double ratio = ljl + p_old_new - S.logjoint[cur * P.slot_scalar + c] - S.hcur[c];   // M:282
        bool take = ratio > 0.0, used_u = false;
        if (!take) {                                    // the uniform is consumed only when Ratio > 0 is false (M:285)
            used_u = true;
            const double ua = P.rng_mode == 0 ? P.tape_u_acc[(size_t)(it - P.tape_base) * P.n_chains + c]
                                              : philox_pair(P, c, it, 0x102u).u0;
            take = ratio > log(ua);
        }
        const int fin = take ? out : cur;
        if (P.tr_theta_end && it < P.tr_iters) {
            if (live) P.tr_theta_end[((size_t)c * P.tr_iters + it) * D + lane] = th;
            if (lane == 0) {
                P.tr_hprop[(size_t)c * P.tr_iters + it] = ratio;
                P.tr_hcur[(size_t)c * P.tr_iters + it] = S.hcur[c];
                P.tr_flags[(size_t)c * P.tr_iters + it] = (take ? 1 : 0) | (used_u ? 2 : 0);
            }
        }
        if (P.samples && it >= P.burn_in && it - P.burn_in < P.sample_cap && live)                 // M:313-315
            P.samples[((size_t)c * P.sample_cap + (it - P.burn_in)) * D + lane] = take ? th : w_cur;
        __syncwarp();
        if (lane == 0) {
            S.cur[c] = fin;
            if (take) ++S.accepted[c];
            S.iter[c] = it + 1;
            ++S.leapfrogs[c];                           // metric evaluations
        }
        cur = fin;
        it += 1;
    }

    if (!do_front || it >= P.it_stop) return;
    // ---- proposal: w' = Mean + R^T z, ProbNewGivenOld                                              M:227-241
    double z = 0.0;
    if (live) z = P.rng_mode == 0 ? P.tape_z[((size_t)(it - P.tape_base) * P.n_chains + c) * D + lane]
                                  : philox_normal(P, c, it, (uint32_t)lane);
    const double* r_cur = S.invg + cur * P.slot_invg + (size_t)c * D * D;
    const double* lg_cur = S.lfac + cur * P.slot_invg + (size_t)c * D * D;
    const double step = mf_matvec<32>(r_cur, z, xs, D, lane);            // (z R)^T: y_i = sum_j R[j][i] z_j
    double w_new = 0.0, diff = 0.0;
    if (live) {
        const double mean = S.theta[cur * P.slot_theta + cd] + S.grad[cur * P.slot_theta + cd];
        w_new = mean + step;
        diff = mean - w_new;
        S.theta_w[cd] = w_new;
    }
    const double y = mf_matvec<32>(lg_cur, diff, xs, D, lane);
    const double p_new_old = -S.logdet[cur * P.slot_scalar + c] - 0.5 * mf_sum<32>(live ? y * y : 0.0, red, D, lane) / cov_c;
    if (lane == 0) {
        S.hcur[c] = p_new_old;
        S.nsteps[c] = 1;                                // k_chain_factor treats the chain as active
        S.step[c] = 0;
    }
}
#endif  // __CUDACC__

}  // namespace rmhmc
