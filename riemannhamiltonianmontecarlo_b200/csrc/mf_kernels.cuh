// Matrix-free metric partials: the per-chain side.
//
// The reference forms the D matrices dG/dw_d = X^T diag(c x_d) X, c_n = v_n (1 - 2 p_n), then
// InvGdG[d] = G^-1 dG_d (rmhmc.py:64-77, :142-156) and uses them in exactly two ways:
//   tr(G^-1 dG_d)                = sum_n c_n x_nd (x_n^T G^-1 x_n)              (rmhmc.py:77,156)
//   p^T G^-1 dG_d G^-1 p         = sum_n c_n x_nd (x_n . u)^2,  u = G^-1 p      (rmhmc.py:105-107,159-161)
// Both right-hand sides are passes over the data with O(N D^2) resp. O(N D) work per chain instead of
// the O(N D^3) of the tensor build, so in this mode the tensor T is never formed:
//   leverages  h = KR2(X) . q, q = packed G^-1 (off-diagonals doubled): one DMMA GEMM per leapfrog step (k_tbuild_pre)
//   trace pass tr_d   = sum_n c_n h_n x_nd
//   quad pass  quad_d = sum_n c_n (x_n . u)^2 x_nd         (F + 1 times per leapfrog step)
// (pass kernels: pass_kernel.cuh for D <= 32, k_metric<MODE 3/4> for 32 < D <= 128).
// What is left per chain is O(D^2): the mat-vecs with G^-1 and L, the fixed-point updates, the
// Hamiltonian and the accept -- the kernels below, one thread per parameter, NTHR = 32 (D <= 32, one
// warp per chain) or 128 (D <= 128).  Round schedule (capi.cu: rmhmc_rounds):
//   k_mf_turn(front) | momentum fixed point | (F-1) x { metric, solve } | closing metric |
//   factor (+ q, u) | leverage GEMM | trace + quad pass | k_mf_turn(back + front) | ...
// where the momentum fixed point is ONE launch (k_pass<MOMFP> / k_mom_fp: F x { quad pass, update }) unless the data
// are row-sharded, F <= 1 or D > 32: then F x { quad pass, [all-reduce,] k_mf_mom_iter }.
#pragma once
#include "chain_kernels.cuh"
#include "common.cuh"

namespace rmhmc {

#ifdef __CUDACC__
template <int NTHR>
__device__ __forceinline__ void mf_sync() {
    if (NTHR == 32) __syncwarp();
    else __syncthreads();
}

// sum over the chain's parameters of a per-thread term (0 for idle threads); fixed order
template <int NTHR>
__device__ __forceinline__ double mf_sum(double term, double* red, int D, int tid) {
    if (NTHR == 32) return warp_sum(term);
    __syncthreads();
    red[tid] = term;
    __syncthreads();
    double s = 0.0;
    for (int b = 0; b < D; ++b) s += red[b];
    return s;
}

// y_i = sum_b M[b][i] x_b for a symmetric D x D matrix in global memory (so that consecutive threads
// read consecutive addresses); x is broadcast through shared memory
template <int NTHR>
__device__ __forceinline__ double mf_matvec(const double* __restrict__ M, double xi, double* xs, int D, int tid) {
    mf_sync<NTHR>();
    xs[tid] = tid < D ? xi : 0.0;
    mf_sync<NTHR>();
    double y0 = 0.0, y1 = 0.0;
    if (tid < D) {
        const double* col = M + tid;
        int b = 0;
#pragma unroll 4
        for (; b + 1 < D; b += 2) {
            y0 = fma(col[(size_t)b * D], xs[b], y0);
            y1 = fma(col[(size_t)(b + 1) * D], xs[b + 1], y1);
        }
        if (b < D) y0 = fma(col[(size_t)b * D], xs[b], y0);
    }
    return y0 + y1;
}

// ---------------------------------------------------------------- per-round chain kernel
// do_back : finish the leapfrog step whose closing passes have just run (metric at theta_w ->
//           grad_tmp / loglik_tmp, factor -> L / G^-1 / log-det[out] and u = G^-1 p, trace pass ->
//           trace_tmp, quad pass -> quad_tmp): gradient, log joint, explicit momentum half-step
//           (R12-R14); if the trajectory is complete: Hamiltonian, accept/reject, store (R15-R18).
//           init != 0: only fill slot `cur`.
// do_front: start the next leapfrog step: [new iteration: momentum draw, H_current (R2-R6)], u = G^-1 p
//           for the first quadratic-form pass of the implicit momentum half-step (R7-R8).
template <int NTHR>
__global__ void __launch_bounds__(NTHR) k_mf_turn(EngineParams P, ChainArrays S, int do_back, int do_front, int init) {
    __shared__ double xs[NTHR], red[NTHR];
    const int c = blockIdx.x, tid = threadIdx.x, D = P.dim;
    if (c >= P.n_chains) return;
    long long it = S.iter[c];
    if (!init && it >= P.it_stop) return;
    const bool live = tid < D;
    const size_t cd = (size_t)c * D + tid;
    int cur = S.cur[c];
    int step = init ? 0 : S.step[c];

    if (do_back) {
        const int nsteps = init ? 1 : S.nsteps[c];
        const int out = init ? cur : 1 - cur;
        const int sgn = init ? 1 : S.dir[c];
        double hprop, p = 0.0;
        bool finished = true;
        if (nsteps > 0) {
            double th = 0.0, grad = 0.0, tr = 0.0;
            if (live) {
                th = S.theta_w[cd];
                grad = S.grad_tmp[cd] - th / P.alpha;                                            // rmhmc.py:140
                tr = S.trace_tmp[cd];                                                            // rmhmc.py:156
                p = init ? 0.0 : S.mom[cd];
            }
            // log prior, summed in parameter order by every thread (rmhmc.py:166, tools.py:10-14)
            mf_sync<NTHR>();
            xs[tid] = th;
            mf_sync<NTHR>();
            const double half_log = 0.5 * log(2.0 * 3.14159265358979323846 * P.alpha);
            double lp = 0.0;
            for (int b = 0; b < D; ++b) lp += -half_log - xs[b] * xs[b] / (2.0 * P.alpha);
            const double ljl = S.loglik_tmp[c] + lp;                                             // rmhmc.py:166-169
            const double logdet = S.logdet[out * P.slot_scalar + c];
            if (live) {
                S.theta[out * P.slot_theta + cd] = th;
                S.grad[out * P.slot_theta + cd] = grad;
                S.trace[out * P.slot_theta + cd] = tr;
            }
            if (tid == 0) S.logjoint[out * P.slot_scalar + c] = ljl;
            if (init) return;

            // ---- R14: explicit closing momentum half-step (quad_tmp = u^T dG_d u, u = G_new^-1 p)
            // LastTerm scaling: 1/2 (rmhmc.py:159-161), Student-t: (1+D)/2 / (1 + p' G^-1 p) with u = G_new^-1 p from the
            // factor kernel (BLR_RMHMC_StudentT.m:368-373)
            double qscale = 0.5;
            if (P.student_t) qscale = 0.5 * (1.0 + D) / (1.0 + mf_sum<NTHR>(live ? p * S.uvec[cd] : 0.0, red, D, tid));
            if (live) {
                p = p + (sgn * P.step_size / 2) * (grad - 0.5 * tr + qscale * S.quad_tmp[cd]);   // rmhmc.py:163
                S.mom[cd] = p;
                if (P.tr_theta_steps && it < P.tr_iters)
                    P.tr_theta_steps[(((size_t)c * P.tr_iters + it) * P.n_leapfrog + step) * D + tid] = th;
            }
            ++step;
            if (tid == 0) ++S.leapfrogs[c];
            if (step < nsteps) {
                if (tid == 0) S.step[c] = step;
                finished = false;
            } else {
                // ---- R15: proposed Hamiltonian
                const double u = mf_matvec<NTHR>(S.invg + out * P.slot_invg + (size_t)c * D * D, p, xs, D, tid);
                const double pgp = mf_sum<NTHR>(live ? p * u : 0.0, red, D, tid);
                hprop = -ljl + logdet + (P.student_t ? 0.5 * (1.0 + D) * log(1.0 + pgp) : 0.5 * pgp);   // rmhmc.py:172 / T:386
            }
        } else {
            // empty trajectory (RandomStep = 0): the proposal is the current state
            hprop = S.hcur[c];
            if (live) p = S.mom[cd];
        }
        if (finished) {
            // ---- R16/R17: accept / reject.  The uniform is consumed only when Ratio > 0 is false.
            double ratio = S.hcur[c] - hprop;
            bool take = ratio > 0.0, used_u = false;
            if (!take) {
                used_u = true;
                double ua = P.rng_mode == 0 ? P.tape_u_acc[(size_t)(it - P.tape_base) * P.n_chains + c]
                                            : philox_pair(P, c, it, 0x102u).u0;
                take = ratio > log(ua);
            }
            const int fin = (take && nsteps > 0) ? out : cur;
            if (P.tr_mom_end && it < P.tr_iters) {
                size_t o = ((size_t)c * P.tr_iters + it) * D + tid;
                if (live) {
                    P.tr_mom_end[o] = p;
                    P.tr_theta_end[o] = S.theta[(nsteps > 0 ? out : cur) * P.slot_theta + cd];
                }
                if (tid == 0) {
                    P.tr_hprop[(size_t)c * P.tr_iters + it] = hprop;
                    P.tr_flags[(size_t)c * P.tr_iters + it] =
                        (take ? 1 : 0) | (used_u ? 2 : 0) | (sgn > 0 ? 16 : 0) | (nsteps << 8);
                }
            }
            // ---- R18: store (row it - burn_in, only for it > burn_in)
            // (the MATLAB loop of the Student-t variant stores iteration BurnIn as well: BLR_RMHMC_StudentT.m:403-405)
            if (P.samples && (it > P.burn_in || (P.student_t && it == P.burn_in)) && it - P.burn_in < P.sample_cap && live)
                P.samples[((size_t)c * P.sample_cap + (it - P.burn_in)) * D + tid] = S.theta[fin * P.slot_theta + cd];
            mf_sync<NTHR>();       // everyone has read the pre-update state
            if (tid == 0) {
                S.cur[c] = fin;
                if (take) ++S.accepted[c];
                S.step[c] = 0;
                S.iter[c] = it + 1;
            }
            cur = fin;
            step = 0;
            it += 1;
        }
    }

    if (!do_front || it >= P.it_stop) return;
    const int in_slot = step == 0 ? cur : 1 - cur;
    const double* invg = S.invg + in_slot * P.slot_invg + (size_t)c * D * D;
    double p = 0.0, u_first;
    int nsteps;
    if (step == 0) {
        // ---- R4-R6: p = L^T z with the current position's Cholesky factor, H_current
        int sgn;
        if (P.ext_mom) {
            // leapfrog seam: the caller supplies p, RandomStep and TimeStep
            if (live) p = P.ext_mom[cd];
            nsteps = P.ext_nsteps[c];
            sgn = P.ext_dir[c];
        } else {
            double z = 0.0, u_step, z_dir, z_chi = 1.0;
            if (P.rng_mode == 0) {
                size_t row = (size_t)(it - P.tape_base) * P.n_chains + c;
                if (live) z = P.tape_z[row * D + tid];
                u_step = P.tape_u_step[row];
                z_dir = P.tape_z_dir[row];
                if (P.student_t) z_chi = P.tape_z_chi[row];
            } else {
                if (live) z = philox_normal(P, c, it, (uint32_t)tid);
                u_step = philox_pair(P, c, it, 0x100u).u0;
                z_dir = philox_normal(P, c, it, 0x101u);
                if (P.student_t) z_chi = philox_normal(P, c, it, 0x103u);
            }
            mf_sync<NTHR>();
            xs[tid] = live ? z : 0.0;
            mf_sync<NTHR>();
            const double* lf = S.lfac + in_slot * P.slot_invg + (size_t)c * D * D;
            if (P.student_t) {
                // mvtrnd(G, 1)' (BLR_RMHMC_StudentT.m:265): normals through the Cholesky factor of the CORRELATION matrix of G,
                // chol(corr) = diag(G)^-1/2 L, divided by sqrt(chi2_1 / 1); no renormalisation hack (that is rmhmc.py's)
                double lz = 0.0, gii = 0.0;
                if (live)
                    for (int k = 0; k <= tid; ++k) { const double l = lf[(size_t)tid * D + k]; lz = fma(l, xs[k], lz); gii = fma(l, l, gii); }
                if (live) p = lz / sqrt(gii) / fabs(z_chi);
            } else {
                if (live)
                    for (int i = tid; i < D; ++i) p = fma(lf[(size_t)i * D + tid], xs[i], p);    // (z L)^T = L^T z, rmhmc.py:80
                const double nrm = sqrt(mf_sum<NTHR>(live ? p * p : 0.0, red, D, tid));
                if (nrm > 100.0) {                                                                // rmhmc.py:81-85
                    p /= nrm * 25.0;
                    if (tid == 0) ++S.renorm_mom[c];
                }
            }
            nsteps = (int)ceil(u_step * (double)P.n_leapfrog);                                    // rmhmc.py:89
            sgn = z_dir > 0.5 ? 1 : -1;                                                           // rmhmc.py:90-93
        }
        const double u = mf_matvec<NTHR>(invg, p, xs, D, tid);
        const double pgp = mf_sum<NTHR>(live ? p * u : 0.0, red, D, tid);
        const double hcur = -S.logjoint[in_slot * P.slot_scalar + c] + S.logdet[in_slot * P.slot_scalar + c] +
                            (P.student_t ? 0.5 * (1.0 + D) * log(1.0 + pgp) : 0.5 * pgp);         // rmhmc.py:175-176 / T:392
        if (tid == 0) {
            if (P.student_t) S.pudot[c] = pgp;
            S.hcur[c] = hcur;
            S.nsteps[c] = nsteps;
            S.dir[c] = sgn;
            S.aslot[c] = in_slot;
        }
        if (P.tr_mom0 && it < P.tr_iters && live) P.tr_mom0[((size_t)c * P.tr_iters + it) * D + tid] = p;
        if (P.tr_hcur && it < P.tr_iters && tid == 0) P.tr_hcur[(size_t)c * P.tr_iters + it] = hcur;
        if (live) {
            S.mom[cd] = p;
            S.uvec[cd] = u;
        }
        if (nsteps <= 0) return;      // u_step == 0: empty trajectory; the next back half finishes the iteration
        u_first = u;
    } else {
        if (live) p = S.mom[cd];
        u_first = mf_matvec<NTHR>(invg, p, xs, D, tid);
        if (P.student_t) {
            const double pgp = mf_sum<NTHR>(live ? p * u_first : 0.0, red, D, tid);
            if (tid == 0) S.pudot[c] = pgp;
        }
        if (tid == 0) S.aslot[c] = in_slot;
        if (live) S.uvec[cd] = u_first;
    }
    if (P.n_fixed == 0) {             // no fixed-point iterations at all: the position does not move (but is clamped)
        const double th = live ? S.theta[in_slot * P.slot_theta + cd] : 0.0;
        const double nrm = sqrt(mf_sum<NTHR>(th * th, red, D, tid));                              // rmhmc.py:125-130
        double div = 1.0;
        if (nrm > 10.0 && !P.student_t) {
            div = nrm * 3.0;
            if (tid == 0) ++S.renorm_pos[c];
        }
        if (live) {
            S.u0[cd] = u_first;
            S.theta_w[cd] = div == 1.0 ? th : th / div;
        }
    }
}

// ---------------------------------------------------------------- implicit momentum half-step, one iterate
// PM <- p + s eps/2 (grad - 0.5 tr + 0.5 quad), u <- G^-1 PM                          (rmhmc.py:102-108)
// is_last: p <- PM (rmhmc.py:110), u0 <- u (rmhmc.py:113) and the first position iterate, whose
// metric is the one already held: theta_w <- theta + s eps/2 (u0 + u0)             (rmhmc.py:116-122)
template <int NTHR>
__global__ void __launch_bounds__(NTHR) k_mf_mom_iter(EngineParams P, ChainArrays S, int is_last) {
    __shared__ double xs[NTHR], red[NTHR];
    const int c = blockIdx.x, tid = threadIdx.x, D = P.dim;
    if (c >= P.n_chains) return;
    if (S.iter[c] >= P.it_stop || S.nsteps[c] <= 0) return;
    const bool live = tid < D;
    const size_t cd = (size_t)c * D + tid;
    const int cur = S.cur[c];
    const int in_slot = S.step[c] == 0 ? cur : 1 - cur;
    const double h = S.dir[c] * P.step_size / 2;
    double pm = 0.0, th = 0.0;
    // LastTerm scaling: 1/2, Student-t: (1+D)/2 / (1 + PM' G^-1 PM) of the iterate the quad pass used (T:293-297)
    const double qscale = P.student_t ? 0.5 * (1.0 + D) / (1.0 + S.pudot[c]) : 0.5;
    if (live) {
        const size_t so = in_slot * P.slot_theta + cd;
        pm = S.mom[cd] + h * (S.grad[so] - 0.5 * S.trace[so] + qscale * S.quad_tmp[cd]);          // rmhmc.py:108
        if (is_last) th = S.theta[so];
    }
    double u = mf_matvec<NTHR>(S.invg + in_slot * P.slot_invg + (size_t)c * D * D, pm, xs, D, tid);
    double pgp = 0.0;
    if (P.student_t) pgp = mf_sum<NTHR>(live ? pm * u : 0.0, red, D, tid);
    if (!is_last) {
        if (live) S.uvec[cd] = u;
        if (P.student_t && tid == 0) S.pudot[c] = pgp;
        return;
    }
    if (P.student_t) u = (1.0 + D) * u / (1.0 + pgp);          // T:309-326: both terms of the position update carry this weight
    double y = th + h * (u + u);
    double div = 1.0;
    if (P.n_fixed <= 1 && !P.student_t) {       // this iterate is already the step's final position
        const double nrm = sqrt(mf_sum<NTHR>(live ? y * y : 0.0, red, D, tid));                   // rmhmc.py:125-130
        if (nrm > 10.0) {
            div = nrm * 3.0;
            if (tid == 0) ++S.renorm_pos[c];
        }
    }
    if (live) {
        S.mom[cd] = pm;
        S.u0[cd] = u;
        S.theta_w[cd] = div == 1.0 ? y : y / div;
    }
}
#endif  // __CUDACC__

}  // namespace rmhmc
