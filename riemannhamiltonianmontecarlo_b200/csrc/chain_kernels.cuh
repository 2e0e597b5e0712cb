// Per-chain stages of the generalized leapfrog: one warp owns one chain, one lane one parameter
// (D <= 32).  Everything that is O(D^3) per chain lives here: Cholesky / inverse / log-det of G,
// tr(G^-1 dG_d), the quadratic forms of the implicit momentum update, the position solves, the
// Hamiltonian and the Metropolis accept.  The O(N D^2) and O(N D^3) contractions over the data are
// the two tensor-core kernels (metric_kernel.cuh, tbuild_kernel.cuh).
//
// Chains run asynchronously: every "round" advances each chain by one leapfrog step of whatever
// trajectory it is on (rmhmc.py:96-163); a chain that finishes its trajectory does its accept/reject
// (rmhmc.py:166-191) at the end of that round and starts the next iteration (rmhmc.py:51-93) at the
// beginning of the following one.  No lock-step over iterations, no compaction, every round is a
// full batch.
#pragma once
#include "common.cuh"

namespace rmhmc {

struct EngineParams {
    int n_chains, dim, ds;             // ds = smem stride of D x D matrices (odd)
    int p2, p2p, p3, p3p, n_rows_pad;
    int n_leapfrog, n_fixed;
    double step_size, alpha;
    long long it_stop;                 // chains idle once they have completed this many iterations
    long long burn_in;
    long long sample_cap;              // rows of the per-chain sample buffer
    // randomness: tape (host supplied draws, parity runs) or counter-based Philox
    int rng_mode;                      // 0 = tape, 1 = philox
    const double* tape_z;              // [W][C][D]
    const double* tape_u_step;         // [W][C]
    const double* tape_z_dir;          // [W][C]
    const double* tape_u_acc;          // [W][C]
    long long tape_base;               // iteration index of tape row 0
    unsigned long long seed;
    long long chain_offset;            // global id of local chain 0 (multi-GPU sharding)
    double* samples;                   // [C][cap][D] or null
    // optional per-step trace for the parity tests (null in production)
    double* tr_theta_steps;            // [C][TI][L][D]
    double* tr_mom_end;                // [C][TI][D]
    double* tr_theta_end;              // [C][TI][D]
    double* tr_mom0;                   // [C][TI][D]
    double* tr_hcur;                   // [C][TI]
    double* tr_hprop;                  // [C][TI]
    int* tr_flags;                     // [C][TI] bit0 accepted, bit1 uniform consumed, bits 8.. nsteps, bit 4 dir>0
    long long tr_iters;                // TI
    const unsigned short* qidx;        // [P2][32]: packed-triple index of (pair, lane d)
    const unsigned char* pair_a;       // [P2]
    const unsigned char* pair_b;       // [P2]
    size_t slot_theta, slot_scalar, slot_gp, slot_invg, slot_t;   // doubles between slot 0 and slot 1
};

__host__ inline size_t chain_smem_bytes(int dim, int p2, int p3p, bool with_t) {
    int ds = dim | 1;
    size_t b = (size_t)2 * dim * ds * 8 + (size_t)p2 * 8 + 4 * 32 * 8 + 8;   // +8: T is 16-byte aligned
    if (with_t) b += (size_t)p3p * 8;
    return b;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ double u01_open(uint32_t hi, uint32_t lo) {   // (0,1), 53 bits
    unsigned long long k = (((unsigned long long)hi << 32) | lo) >> 11;
    return ((double)k + 0.5) * (1.0 / 9007199254740992.0);
}
struct Draw { double u0, u1; };
__device__ __forceinline__ Draw philox_pair(const EngineParams& P, long long chain, long long it, uint32_t what) {
    unsigned long long cid = (unsigned long long)(P.chain_offset + chain);
    uint4 ctr = make_uint4((uint32_t)cid, (uint32_t)(cid >> 32) ^ (what << 8), (uint32_t)it, (uint32_t)((unsigned long long)it >> 32));
    uint4 r = philox4x32(ctr, make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32)));
    return Draw{u01_open(r.x, r.y), u01_open(r.z, r.w)};
}
__device__ __forceinline__ double philox_normal(const EngineParams& P, long long chain, long long it, uint32_t what) {
    Draw d = philox_pair(P, chain, it, what);
    return sqrt(-2.0 * log(d.u0)) * cospi(2.0 * d.u1);
}

// ---------------------------------------------------------------- warp linear algebra (D <= 32)
// dense symmetric matrix from the packed upper triangle
__device__ __forceinline__ void unpack_sym(const double* __restrict__ gp, double* A, int D, int DS, int lane) {
    for (int idx = lane; idx < D * D; idx += 32) {
        int i = idx / D, j = idx - i * D;
        int lo = i < j ? i : j, hi = i < j ? j : i;
        A[i * DS + j] = gp[pair_index(lo, hi, D)];
    }
    __syncwarp();
}

// In-place lower Cholesky factor (A = L L^T; strict upper part left untouched).  Returns
// sum_k log L_kk = 0.5 log|A| (rmhmc.py:171,175).  A non-PD matrix yields NaNs, which the accept
// test then rejects (the reference would raise LinAlgError; it never happens since G >= I/alpha).
__device__ __forceinline__ double chol_warp(double* A, int D, int DS, int lane) {
    for (int k = 0; k < D; ++k) {
        double lkk = sqrt(A[k * DS + k]);
        double lik = 0.0;
        bool below = lane > k && lane < D;
        if (below) lik = A[lane * DS + k] / lkk;
        __syncwarp();
        if (lane == k) A[k * DS + k] = lkk;
        if (below) A[lane * DS + k] = lik;
        __syncwarp();
        if (below)
            for (int j = k + 1; j <= lane; ++j) A[lane * DS + j] -= lik * A[j * DS + k];
        __syncwarp();
    }
    double l = lane < D ? log(A[lane * DS + lane]) : 0.0;
    return warp_sum(l);
}

// Solve L L^T x = b; lane i holds b_i on entry and x_i on exit.
__device__ __forceinline__ double chol_solve_warp(const double* L, int D, int DS, int lane, double b) {
    double dinv = lane < D ? 1.0 / L[lane * DS + lane] : 0.0;
    for (int k = 0; k < D; ++k) {                       // forward: L y = b
        double yk = __shfl_sync(0xffffffffu, b, k) * __shfl_sync(0xffffffffu, dinv, k);
        if (lane == k) b = yk;
        if (lane > k && lane < D) b -= L[lane * DS + k] * yk;
    }
    for (int k = D - 1; k >= 0; --k) {                  // backward: L^T x = y
        double xk = __shfl_sync(0xffffffffu, b, k) * __shfl_sync(0xffffffffu, dinv, k);
        if (lane == k) b = xk;
        if (lane < k) b -= L[k * DS + lane] * xk;
    }
    return b;
}

// B = (L L^T)^-1, one right-hand side (column) per lane.
__device__ __forceinline__ void chol_inverse_warp(const double* L, double* B, int D, int DS, int lane) {
    const int j = lane < D ? lane : 0;      // idle lanes shadow column 0 without storing
    const bool live = lane < D;
    for (int i = 0; i < D; ++i) {           // forward sweep: L Y = I
        double s = (i == j) ? 1.0 : 0.0;
        for (int k = 0; k < i; ++k) s -= L[i * DS + k] * B[k * DS + j];
        s /= L[i * DS + i];
        __syncwarp();
        if (live) B[i * DS + j] = s;
        __syncwarp();
    }
    for (int i = D - 1; i >= 0; --i) {      // backward sweep: L^T X = Y
        double s = B[i * DS + j];
        for (int k = i + 1; k < D; ++k) s -= L[k * DS + i] * B[k * DS + j];
        s /= L[i * DS + i];
        __syncwarp();
        if (live) B[i * DS + j] = s;
        __syncwarp();
    }
}

// y_i = sum_j M[i][j] x_j, x given per lane (staged through xv in smem)
__device__ __forceinline__ double matvec_warp(const double* M, double* xv, int D, int DS, int lane, double x) {
    __syncwarp();
    if (lane < D) xv[lane] = x;
    __syncwarp();
    double y = 0.0;
    if (lane < D)
        for (int j = 0; j < D; ++j) y += M[lane * DS + j] * xv[j];
    return y;
}

// out_d = sum_{a<=b} Q[(a,b)] T[d,a,b] for lane d, T packed in smem, Q packed weights (off-diagonal
// pairs already carry their factor 2).
__device__ __forceinline__ double tensor_contract(const double* Tsm, const double* Q, const unsigned short* __restrict__ qidx,
                                                  int P2, int lane) {
    double s0 = 0.0, s1 = 0.0;
    int pr = 0;
    for (; pr + 1 < P2; pr += 2) {
        s0 += Q[pr] * Tsm[__ldg(qidx + pr * 32 + lane)];
        s1 += Q[pr + 1] * Tsm[__ldg(qidx + (pr + 1) * 32 + lane)];
    }
    if (pr < P2) s0 += Q[pr] * Tsm[__ldg(qidx + pr * 32 + lane)];
    return s0 + s1;
}

// LastTerm_d = 0.5 u^T dG_d u (rmhmc.py:105-107 with u = G^-1 p; G^-1 symmetric)
__device__ __forceinline__ double last_term(const EngineParams& P, const double* Tsm, double* Q, double* uv,
                                            int lane, double u) {
    __syncwarp();
    if (lane < P.dim) uv[lane] = u;
    __syncwarp();
    for (int pr = lane; pr < P.p2; pr += 32) {
        int pa = P.pair_a[pr], pb = P.pair_b[pr];
        double w = uv[pa] * uv[pb];
        Q[pr] = pa == pb ? w : 2.0 * w;
    }
    __syncwarp();
    return 0.5 * tensor_contract(Tsm, Q, P.qidx, P.p2, lane);
}

struct ChainSmem {
    double *A, *B, *Q, *v0, *v1, *v2, *v3, *T;
};
__device__ __forceinline__ ChainSmem carve_chain_smem(unsigned char* raw, const EngineParams& P, bool with_t) {
    ChainSmem s;
    double* p = reinterpret_cast<double*>(raw);
    s.A = p; p += P.dim * P.ds;
    s.B = p; p += P.dim * P.ds;
    s.Q = p; p += P.p2;
    s.v0 = p; p += 32; s.v1 = p; p += 32; s.v2 = p; p += 32; s.v3 = p; p += 32;
    if ((p - reinterpret_cast<double*>(raw)) & 1) ++p;      // double2 loads into T
    s.T = with_t ? p : nullptr;
    return s;
}

__device__ __forceinline__ void load_t_smem(double* Tsm, const double* __restrict__ src, int p3p, int lane) {
    const double2* s2 = reinterpret_cast<const double2*>(src);
    double2* d2 = reinterpret_cast<double2*>(Tsm);
    for (int i = lane; i < p3p / 2; i += 32) d2[i] = s2[i];
    __syncwarp();
}

__device__ __forceinline__ double clamp_position(double w, int lane, int dim, int* counter) {
    // rmhmc.py:125-130: if ||w|| > 10: w /= ||w|| * 3
    double n2 = warp_sum(lane < dim ? w * w : 0.0);
    double nrm = sqrt(n2);
    if (nrm > 10.0) {
        w /= nrm * 3.0;
        if (lane == 0) ++*counter;
    }
    return w;
}

// ---------------------------------------------------------------- round stage 1
// [new iteration: momentum draw, H_current] + implicit momentum half-step + first position iterate
__global__ void __launch_bounds__(32) k_chain_front(EngineParams P, ChainArrays S) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = blockIdx.x, lane = threadIdx.x, D = P.dim, DS = P.ds;
    if (c >= P.n_chains) return;
    long long it = S.iter[c];
    if (it >= P.it_stop) return;
    ChainSmem sm = carve_chain_smem(smem_raw, P, true);
    const int cur = S.cur[c];
    int step = S.step[c];
    const bool live = lane < D;
    double p = 0.0;
    int sgn, nsteps;

    if (step == 0) {
        // ---- R2/R4-R6: factor G at the current position, draw p = L^T z, H_current
        unpack_sym(S.gp + cur * P.slot_gp + (size_t)c * P.p2p, sm.A, D, DS, lane);
        chol_warp(sm.A, D, DS, lane);
        double z = 0.0, u_step, z_dir;
        if (P.rng_mode == 0) {
            size_t row = (size_t)(it - P.tape_base) * P.n_chains + c;
            if (live) z = P.tape_z[row * D + lane];
            u_step = P.tape_u_step[row];
            z_dir = P.tape_z_dir[row];
        } else {
            if (live) z = philox_normal(P, c, it, (uint32_t)lane);
            u_step = philox_pair(P, c, it, 32u).u0;
            z_dir = philox_normal(P, c, it, 33u);
        }
        __syncwarp();
        if (live) sm.v0[lane] = z;
        __syncwarp();
        if (live)
            for (int i = lane; i < D; ++i) p += sm.A[i * DS + lane] * sm.v0[i];   // (z L)^T = L^T z, rmhmc.py:80
        double nrm = sqrt(warp_sum(p * p));
        if (nrm > 100.0) {                                                          // rmhmc.py:81-85
            p /= nrm * 25.0;
            if (lane == 0) ++S.renorm_mom[c];
        }
        nsteps = (int)ceil(u_step * (double)P.n_leapfrog);                          // rmhmc.py:89
        sgn = z_dir > 0.5 ? 1 : -1;                                                 // rmhmc.py:90-93
        // InvG of the current slot -> B
        const double* ig = S.invg + cur * P.slot_invg + (size_t)c * D * D;
        for (int idx = lane; idx < D * D; idx += 32) sm.B[(idx / D) * DS + (idx % D)] = ig[idx];
        __syncwarp();
        double u = matvec_warp(sm.B, sm.v1, D, DS, lane, p);
        double kin = 0.5 * warp_sum(live ? p * u : 0.0);
        double hcur = -S.logjoint[cur * P.slot_scalar + c] + S.logdet[cur * P.slot_scalar + c] + kin;  // rmhmc.py:175-176
        if (lane == 0) {
            S.hcur[c] = hcur;
            S.nsteps[c] = nsteps;
            S.dir[c] = sgn;
        }
        if (P.tr_mom0 && it < P.tr_iters && live) P.tr_mom0[((size_t)c * P.tr_iters + it) * D + lane] = p;
        if (P.tr_hcur && it < P.tr_iters && lane == 0) P.tr_hcur[(size_t)c * P.tr_iters + it] = hcur;
        if (nsteps <= 0) {            // u_step == 0: empty trajectory; k_chain_back finishes the iteration
            if (live) S.mom[(size_t)c * D + lane] = p;
            return;
        }
    } else {
        nsteps = S.nsteps[c];
        sgn = S.dir[c];
        if (live) p = S.mom[(size_t)c * D + lane];
    }

    // ---- R7/R8: implicit momentum half-step with the metric quantities of the step's start point
    const int in_slot = step == 0 ? cur : 1 - cur;
    if (step != 0) {
        const double* ig = S.invg + in_slot * P.slot_invg + (size_t)c * D * D;
        for (int idx = lane; idx < D * D; idx += 32) sm.B[(idx / D) * DS + (idx % D)] = ig[idx];
        __syncwarp();
    }
    load_t_smem(sm.T, S.tpack + in_slot * P.slot_t + (size_t)c * P.p3p, P.p3p, lane);
    double grad = 0.0, tr = 0.0, w = 0.0;
    if (live) {
        grad = S.grad[in_slot * P.slot_theta + (size_t)c * D + lane];
        tr = S.trace[in_slot * P.slot_theta + (size_t)c * D + lane];
        w = S.theta[in_slot * P.slot_theta + (size_t)c * D + lane];
    }
    const double h = sgn * P.step_size / 2;
    const double base = grad - 0.5 * tr;
    double pm = p;
    for (int fi = 0; fi < P.n_fixed; ++fi) {
        double u = matvec_warp(sm.B, sm.v1, D, DS, lane, pm);
        double last = last_term(P, sm.T, sm.Q, sm.v2, lane, u);
        pm = p + h * (base + last);
    }
    p = pm;
    // ---- R9 and the first position iterate (its metric is the one we already hold)
    double u0 = matvec_warp(sm.B, sm.v1, D, DS, lane, p);
    double pw = w + h * (u0 + u0);
    if (P.n_fixed <= 1) {
        if (P.n_fixed == 0) pw = w;
        pw = clamp_position(pw, lane, D, &S.renorm_pos[c]);
    }
    if (live) {
        S.mom[(size_t)c * D + lane] = p;
        S.u0[(size_t)c * D + lane] = u0;
        S.theta_w[(size_t)c * D + lane] = pw;
    }
}

// ---------------------------------------------------------------- round stage 2 (x (F-1))
// position fixed-point iterate: solve G(theta_w) u = p, theta_w <- theta + s eps/2 (u0 + u)
__global__ void __launch_bounds__(32) k_chain_solve(EngineParams P, ChainArrays S, int is_last) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = blockIdx.x, lane = threadIdx.x, D = P.dim, DS = P.ds;
    if (c >= P.n_chains) return;
    if (S.iter[c] >= P.it_stop || S.nsteps[c] <= 0) return;
    ChainSmem sm = carve_chain_smem(smem_raw, P, false);
    const bool live = lane < D;
    const int cur = S.cur[c];
    const int in_slot = S.step[c] == 0 ? cur : 1 - cur;
    unpack_sym(S.g_tmp + (size_t)c * P.p2p, sm.A, D, DS, lane);
    chol_warp(sm.A, D, DS, lane);
    double p = live ? S.mom[(size_t)c * D + lane] : 0.0;
    double u = chol_solve_warp(sm.A, D, DS, lane, p);                   // rmhmc.py:121
    double w = live ? S.theta[in_slot * P.slot_theta + (size_t)c * D + lane] : 0.0;
    double u0 = live ? S.u0[(size_t)c * D + lane] : 0.0;
    double pw = w + (S.dir[c] * P.step_size / 2) * (u0 + u);            // rmhmc.py:122
    if (is_last) pw = clamp_position(pw, lane, D, &S.renorm_pos[c]);
    if (live) S.theta_w[(size_t)c * D + lane] = pw;
}

// ---------------------------------------------------------------- round stage 3
// metric quantities at the new position, explicit closing momentum half-step, and -- when the
// trajectory is complete -- Hamiltonian, accept/reject, sample store.  With init != 0 it only
// fills slot `cur` from the builds at theta_w (sampler start-up).
__global__ void __launch_bounds__(32) k_chain_back(EngineParams P, ChainArrays S, int init) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = blockIdx.x, lane = threadIdx.x, D = P.dim, DS = P.ds;
    if (c >= P.n_chains) return;
    long long it = S.iter[c];
    if (!init && it >= P.it_stop) return;
    ChainSmem sm = carve_chain_smem(smem_raw, P, true);
    const bool live = lane < D;
    const int cur = S.cur[c];
    const int nsteps = init ? 1 : S.nsteps[c];
    const int out = init ? cur : 1 - cur;
    double hprop = 0.0, p = 0.0;
    int step = init ? 0 : S.step[c];

    if (nsteps > 0) {
        // ---- R12/R13: factor G(theta_w), inverse, log-det, traces; R7': gradient; log joint
        const double* gsrc = S.g_tmp + (size_t)c * P.p2p;
        double* gdst = S.gp + out * P.slot_gp + (size_t)c * P.p2p;
        for (int i = lane; i < P.p2p; i += 32) gdst[i] = gsrc[i];
        unpack_sym(gsrc, sm.A, D, DS, lane);
        double logdet = chol_warp(sm.A, D, DS, lane);
        chol_inverse_warp(sm.A, sm.B, D, DS, lane);
        double* ig = S.invg + out * P.slot_invg + (size_t)c * D * D;
        for (int idx = lane; idx < D * D; idx += 32) ig[idx] = sm.B[(idx / D) * DS + (idx % D)];
        load_t_smem(sm.T, S.tpack + out * P.slot_t + (size_t)c * P.p3p, P.p3p, lane);
        for (int pr = lane; pr < P.p2; pr += 32) {
            int pa = P.pair_a[pr], pb = P.pair_b[pr];
            double w = sm.B[pa * DS + pb];
            sm.Q[pr] = pa == pb ? w : 2.0 * w;
        }
        __syncwarp();
        double tr = tensor_contract(sm.T, sm.Q, P.qidx, P.p2, lane);    // tr(G^-1 dG_d), rmhmc.py:156
        double th = live ? S.theta_w[(size_t)c * D + lane] : 0.0;
        double grad = live ? S.grad_tmp[(size_t)c * D + lane] - th / P.alpha : 0.0;   // rmhmc.py:140
        double lp = live ? -0.5 * log(2.0 * 3.14159265358979323846 * P.alpha) - th * th / (2.0 * P.alpha) : 0.0;
        double ljl = S.loglik_tmp[c] + warp_sum(lp);                    // rmhmc.py:166-169, tools.py:10-14
        if (live) {
            S.theta[out * P.slot_theta + (size_t)c * D + lane] = th;
            S.grad[out * P.slot_theta + (size_t)c * D + lane] = grad;
            S.trace[out * P.slot_theta + (size_t)c * D + lane] = tr;
        }
        if (lane == 0) {
            S.logjoint[out * P.slot_scalar + c] = ljl;
            S.logdet[out * P.slot_scalar + c] = logdet;
        }
        if (init) return;

        // ---- R14: explicit closing momentum half-step
        p = live ? S.mom[(size_t)c * D + lane] : 0.0;
        double u = matvec_warp(sm.B, sm.v1, D, DS, lane, p);
        double last = last_term(P, sm.T, sm.Q, sm.v2, lane, u);
        p += (S.dir[c] * P.step_size / 2) * (grad - 0.5 * tr + last);
        if (live) S.mom[(size_t)c * D + lane] = p;
        if (P.tr_theta_steps && it < P.tr_iters && live)
            P.tr_theta_steps[(((size_t)c * P.tr_iters + it) * P.n_leapfrog + step) * D + lane] = th;
        ++step;
        if (lane == 0) ++S.leapfrogs[c];
        if (step < nsteps) {
            if (lane == 0) S.step[c] = step;
            return;
        }
        // ---- R15: proposed Hamiltonian
        double u2 = matvec_warp(sm.B, sm.v1, D, DS, lane, p);
        hprop = -ljl + logdet + 0.5 * warp_sum(live ? p * u2 : 0.0);
    } else {
        // empty trajectory (RandomStep = 0): the proposal is the current state
        hprop = S.hcur[c];
        p = live ? S.mom[(size_t)c * D + lane] : 0.0;
    }

    // ---- R16/R17: accept / reject.  The uniform is consumed only when Ratio > 0 is false.
    double ratio = S.hcur[c] - hprop;
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
            P.tr_hprop[(size_t)c * P.tr_iters + it] = hprop;
            P.tr_flags[(size_t)c * P.tr_iters + it] =
                (take ? 1 : 0) | (used_u ? 2 : 0) | (S.dir[c] > 0 ? 16 : 0) | (nsteps << 8);
        }
    }
    // ---- R18: store (row it - burn_in, only for it > burn_in)
    if (P.samples && it > P.burn_in && it - P.burn_in < P.sample_cap && live)
        P.samples[((size_t)c * P.sample_cap + (it - P.burn_in)) * D + lane] =
            S.theta[fin * P.slot_theta + (size_t)c * D + lane];
    if (lane == 0) {
        S.cur[c] = fin;
        if (take) ++S.accepted[c];
        S.step[c] = 0;
        S.iter[c] = it + 1;
    }
}
#endif  // __CUDACC__

}  // namespace rmhmc
