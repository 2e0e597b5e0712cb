// Per-chain stages of the generalized leapfrog: one warp owns one chain, one lane one parameter
// (D <= 32).  Everything that is O(D^3) per chain lives here: Cholesky / inverse / log-det of G,
// tr(G^-1 dG_d), the quadratic forms of the implicit momentum update, the position solves, the
// Hamiltonian and the Metropolis accept.  The O(N D^2) and O(N D^3) contractions over the data are
// the tensor-core kernels (metric_kernel.cuh, tbuild_kernel.cuh, pass_kernel.cuh).  k_chain_turn and the
// tensor contractions below belong to the TENSOR partials mode; the MATRIX_FREE mode (default) uses
// mf_kernels.cuh for the per-chain halves and shares k_chain_factor / k_chain_solve.
//
// Chains run asynchronously: every "round" advances each chain by one leapfrog step of whatever
// trajectory it is on (rmhmc.py:96-163); a chain that finishes its trajectory does its accept/reject
// (rmhmc.py:166-191) at the end of that round and starts the next iteration (rmhmc.py:51-93) at the
// beginning of the following one.  No lock-step over iterations, no compaction, every round is a
// full batch.
#pragma once
#include "chain_big.cuh"
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
    // Student-t kinetic energy (BLR_RMHMC_StudentT.m; matrix-free partials, D <= 32): K = (1+D)/2 log(1 + p' G^-1 p),
    // momentum p = D^-1/2 L z / |z_chi| (mvtrnd(G, 1): correlation-scaled normals over sqrt(chi2_1))
    int student_t;
    const double* tape_z_chi;          // [W][C] normal whose square is the chi-square draw
    long long tape_base;               // iteration index of tape row 0
    unsigned long long seed;
    long long chain_offset;            // global id of local chain 0 (multi-GPU sharding)
    // leapfrog seam (rmhmc_leapfrog): momentum, trajectory length and direction supplied by the caller
    const double* ext_mom;             // [C][D] or null
    const int* ext_nsteps;             // [C]
    const int* ext_dir;                // [C]
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
    const unsigned short* tidx;        // [D][P2]: packed-triple index of (d, pair) (D <= 32)
    const unsigned int* tidx32;        // same, 32-bit (D > 32)
    const unsigned char* pair_a;       // [P2]
    const unsigned char* pair_b;       // [P2]
    size_t slot_theta, slot_scalar, slot_invg, slot_t;   // doubles between slot 0 and slot 1
    // matrix-free partials
    int matrix_free;                   // 1: no packed T; traces / quadratic forms by passes over the data
    int p2k;                           // P2 padded to the K tile of the leverage GEMM (32)
    size_t slot_cw;                    // doubles between the two c_n slots
};

__host__ inline size_t factor_smem_bytes(int order) { return ((size_t)2 * order * (order | 1) + 64) * 8; }
__host__ inline size_t solve_smem_bytes(int order) { return ((size_t)order * (order | 1) + 64) * 8; }

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
// One lane owns one row.  Matrices live in registers (statically indexed, loops unrolled to DMAX
// with warp-uniform guards on the runtime D) and are mirrored in shared memory where another lane's
// column is needed.  Cross-lane traffic is warp shuffles; no __syncthreads anywhere.
constexpr unsigned kFull = 0xffffffffu;

// row[j] = G[lane][j], j <= lane, from the packed upper triangle (coalesced: consecutive lanes read
// consecutive packed entries for a fixed j)
template <int DMAX>
__device__ __forceinline__ void load_packed_rows(const double* __restrict__ gp, double (&row)[DMAX], int D, int lane) {
#pragma unroll
    for (int j = 0; j < DMAX; ++j)
        row[j] = (j < D && lane >= j && lane < D) ? gp[j * D - j * (j - 1) / 2 + (lane - j)] : 0.0;
}

// In-register lower Cholesky factor (row[j] = L[lane][j], j <= lane).  Returns
// sum_k log L_kk = 0.5 log|G| (rmhmc.py:171,175); dinv = 1 / L[lane][lane].  A non-PD matrix yields
// NaNs, which the accept test then rejects (the reference would raise LinAlgError; it cannot happen
// since G >= I/alpha).
// 1/sqrt(a): rsqrt.approx.ftz.f64 (20-bit seed) + two Newton steps (~1 ulp); NaN for a < 0 like sqrt
__device__ __forceinline__ double fast_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double h = 0.5 * a;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
}

template <int DMAX>
__device__ __forceinline__ double chol_regs(double (&row)[DMAX], int D, int lane, double& dinv) {
    double diag = 1.0, dinv_acc = 0.0;
#pragma unroll
    for (int k = 0; k < DMAX; ++k) {
        if (k < D) {
            double akk = __shfl_sync(kFull, row[k], k);
            double inv = fast_rsqrt(akk);
            double lkk = akk * inv;
            double lik = (lane == k) ? lkk : row[k] * inv;
            if (lane == k) { diag = lkk; dinv_acc = inv; }
            row[k] = lik;
#pragma unroll
            for (int j = k + 1; j < DMAX; ++j) {
                if (j < D) {
                    double ljk = __shfl_sync(kFull, lik, j);
                    if (lane >= j) row[j] = fma(-lik, ljk, row[j]);
                }
            }
        }
    }
    dinv = dinv_acc;
    return warp_sum(lane < D ? log(diag) : 0.0);
}

// Same factorisation, but column k is broadcast through a 32-double shared buffer (one LDS per
// trailing element instead of two SHFL): fewer instructions in the instruction-bound per-chain kernels.
template <int DMAX>
__device__ __forceinline__ double chol_regs_sm(double (&row)[DMAX], double* colbuf, int D, int lane, double& dinv) {
    double diag = 1.0, dinv_acc = 0.0;
#pragma unroll
    for (int k = 0; k < DMAX; ++k) {
        if (k < D) {
            double akk = __shfl_sync(kFull, row[k], k);
            double inv = fast_rsqrt(akk);
            double lkk = akk * inv;
            double lik = (lane == k) ? lkk : row[k] * inv;
            if (lane == k) { diag = lkk; dinv_acc = inv; }
            row[k] = lik;
            __syncwarp();
            colbuf[lane] = lik;
            __syncwarp();
#pragma unroll
            for (int j = k + 1; j < DMAX; ++j) {
                if (j < D) {
                    double ljk = colbuf[j];
                    if (lane >= j) row[j] = fma(-lik, ljk, row[j]);
                }
            }
        }
    }
    dinv = dinv_acc;
    return warp_sum(lane < D ? log(diag) : 0.0);
}

template <int DMAX>
__device__ __forceinline__ void store_rows(double* Lsm, const double (&row)[DMAX], int D, int DS, int lane) {
#pragma unroll
    for (int j = 0; j < DMAX; ++j)
        if (j < D && lane < D) Lsm[lane * DS + j] = (j <= lane) ? row[j] : 0.0;
    __syncwarp();
}

// Solve L L^T x = b; lane i holds b_i on entry and x_i on exit (rmhmc.py:113,121).
template <int DMAX>
__device__ __forceinline__ double chol_solve_regs(const double (&row)[DMAX], const double* Lsm, double dinv, int D,
                                                  int DS, int lane, double b) {
#pragma unroll
    for (int k = 0; k < DMAX; ++k) {                    // forward: L y = b
        if (k < D) {
            double yk = __shfl_sync(kFull, b, k) * __shfl_sync(kFull, dinv, k);
            if (lane == k) b = yk;
            else if (lane > k) b = fma(-row[k], yk, b);
        }
    }
#pragma unroll
    for (int k = DMAX - 1; k >= 0; --k) {               // backward: L^T x = y
        if (k < D) {
            double xk = __shfl_sync(kFull, b, k) * __shfl_sync(kFull, dinv, k);
            double lkj = lane < k ? Lsm[k * DS + lane] : 0.0;
            if (lane == k) b = xk;
            else if (lane < k) b = fma(-lkj, xk, b);
        }
    }
    return b;
}

// ig[b] = (L L^T)^-1 [lane][b].  Lane j first builds column j of M = L^-1 (Msm mirrors it), then
// G^-1 = M^T M.
template <int DMAX>
__device__ __forceinline__ void chol_inverse_regs(const double* Lsm, double* Msm, double dinv, double (&ig)[DMAX],
                                                  int D, int DS, int lane) {
    double m[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) {
        m[i] = 0.0;
        if (i < D) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int k = 0; k < i; ++k) {
                if (k & 1) s1 = fma(Lsm[i * DS + k], m[k], s1);
                else s0 = fma(Lsm[i * DS + k], m[k], s0);
            }
            double di = __shfl_sync(kFull, dinv, i);
            m[i] = (i == lane) ? di : (i > lane ? -(s0 + s1) * di : 0.0);
        }
    }
#pragma unroll
    for (int i = 0; i < DMAX; ++i)
        if (i < D && lane < D) Msm[i * DS + lane] = m[i];
    __syncwarp();
#pragma unroll
    for (int b = 0; b < DMAX; ++b) {
        double s0 = 0.0, s1 = 0.0;
        if (b < D) {
#pragma unroll
            for (int k = b; k < DMAX; ++k) {
                if (k < D) {
                    if (k & 1) s1 = fma(m[k], Msm[k * DS + b], s1);
                    else s0 = fma(m[k], Msm[k * DS + b], s0);
                }
            }
        }
        ig[b] = s0 + s1;
    }
    __syncwarp();
}

// ---- variants working on a PACKED lower factor in shared memory (L[i][j], j <= i, stored at
// pair_index(j, i)): used by the epilogues fused into the metric kernel, where the chain's packed
// metric tile is overwritten in place by its factor.
template <int DMAX>
__device__ __forceinline__ void store_rows_packed(double* Lp, const double (&row)[DMAX], int D, int lane) {
#pragma unroll
    for (int j = 0; j < DMAX; ++j)
        if (j < D && lane < D && j <= lane) Lp[pair_index(j, lane, D)] = row[j];
    __syncwarp();
}

template <int DMAX>
__device__ __forceinline__ double chol_solve_packed(const double (&row)[DMAX], const double* Lp, double dinv, int D,
                                                    int lane, double b) {
#pragma unroll
    for (int k = 0; k < DMAX; ++k) {                    // forward: L y = b
        if (k < D) {
            double yk = __shfl_sync(kFull, b, k) * __shfl_sync(kFull, dinv, k);
            if (lane == k) b = yk;
            else if (lane > k) b = fma(-row[k], yk, b);
        }
    }
    const int lbase = lane * D - lane * (lane - 1) / 2 - lane;      // pair_index(lane, k) = lbase + k
#pragma unroll
    for (int k = DMAX - 1; k >= 0; --k) {               // backward: L^T x = y
        if (k < D) {
            double xk = __shfl_sync(kFull, b, k) * __shfl_sync(kFull, dinv, k);
            double lkj = lane < k ? Lp[lbase + k] : 0.0;
            if (lane == k) b = xk;
            else if (lane < k) b = fma(-lkj, xk, b);
        }
    }
    return b;
}

// ig[b] = (L L^T)^-1 [lane][b] with L packed in Lp; Mp is a packed scratch of the same size.
template <int DMAX>
__device__ __forceinline__ void chol_inverse_packed(const double* Lp, double* Mp, double dinv, double (&ig)[DMAX],
                                                    int D, int lane) {
    double m[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) {
        m[i] = 0.0;
        if (i < D) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int k = 0; k < i; ++k) {
                double lik = Lp[k * D - k * (k - 1) / 2 + (i - k)];      // L[i][k], same address for all lanes
                if (k & 1) s1 = fma(lik, m[k], s1);
                else s0 = fma(lik, m[k], s0);
            }
            double di = __shfl_sync(kFull, dinv, i);
            m[i] = (i == lane) ? di : (i > lane ? -(s0 + s1) * di : 0.0);
        }
    }
    const int lbase = lane * D - lane * (lane - 1) / 2 - lane;
#pragma unroll
    for (int i = 0; i < DMAX; ++i)
        if (i < D && lane < D && i >= lane) Mp[lbase + i] = m[i];          // M[i][lane]
    __syncwarp();
#pragma unroll
    for (int b = 0; b < DMAX; ++b) {
        double s0 = 0.0, s1 = 0.0;
        if (b < D) {
            const int bbase = b * D - b * (b - 1) / 2 - b;
#pragma unroll
            for (int k = b; k < DMAX; ++k) {
                if (k < D) {
                    if (k & 1) s1 = fma(m[k], Mp[bbase + k], s1);          // M[k][b]
                    else s0 = fma(m[k], Mp[bbase + k], s0);
                }
            }
        }
        ig[b] = s0 + s1;
    }
    __syncwarp();
}


// ---------------------------------------------------------------- fixed-order variants (engine kernels)
// The per-chain kernels are instruction-issue bound (ncu: 5.4 k warp instructions per 25 x 25 solve, more than
// half of them loop guards and lane predicates).  These variants take the order N at compile time
// (N >= D; lanes / columns >= D carry the identity, i.e. the matrix factored is blockdiag(G, I), whose
// factor, inverse and log-det restricted to the leading D x D block are those of G) and update the unused
// strict upper triangle along with the rest instead of predicating it away -- it never feeds a used entry.
template <int N>
__device__ __forceinline__ void load_packed_rows_pad(const double* __restrict__ gp, double (&row)[N], int D, int lane) {
#pragma unroll
    for (int j = 0; j < N; ++j)
        row[j] = (j < D && lane >= j && lane < D) ? gp[j * D - j * (j - 1) / 2 + (lane - j)] : (j == lane ? 1.0 : 0.0);
}

// colbuf: 2 x 32 doubles of shared memory (column k is broadcast through buffer k & 1: one __syncwarp per column)
template <int N>
__device__ __forceinline__ double chol_fixed(double (&row)[N], double* colbuf, int lane, double& dinv) {
    double diag = 1.0, dinv_acc = 1.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double akk = __shfl_sync(kFull, row[k], k);
        const double inv = fast_rsqrt(akk);
        const double lkk = akk * inv;
        const double lik = (lane == k) ? lkk : row[k] * inv;
        if (lane == k) { diag = lkk; dinv_acc = inv; }
        row[k] = lik;
        double* cb = colbuf + (k & 1) * 32;
        cb[lane] = lik;
        __syncwarp();
        // column k broadcast two entries per shared load (16-byte aligned pairs; k is a compile-time constant)
        if (((k + 1) & 1) && k + 1 < N) row[k + 1] = fma(-lik, cb[k + 1], row[k + 1]);
#pragma unroll
        for (int j = (k + 2) & ~1; j + 1 < N; j += 2) {
            const double2 cj = *reinterpret_cast<const double2*>(cb + j);
            row[j] = fma(-lik, cj.x, row[j]);
            row[j + 1] = fma(-lik, cj.y, row[j + 1]);
        }
        if ((N & 1) && ((k + 2) & ~1) <= N - 1) row[N - 1] = fma(-lik, cb[N - 1], row[N - 1]);
    }
    dinv = dinv_acc;
    return warp_sum(lane < N ? log(diag) : 0.0);
}

template <int N>
__device__ __forceinline__ void store_rows_fixed(double* Lsm, const double (&row)[N], int lane) {
    constexpr int NS = N | 1;
    if (lane < N) {
#pragma unroll
        for (int j = 0; j < N; ++j) Lsm[lane * NS + j] = (j <= lane) ? row[j] : 0.0;
    }
    __syncwarp();
}

// Solve L L^T x = b; lane i holds b_i on entry and x_i on exit (rmhmc.py:113,121)
template <int N>
__device__ __forceinline__ double chol_solve_fixed(const double (&row)[N], const double* Lsm, double dinv, int lane, double b) {
    constexpr int NS = N | 1;
#pragma unroll
    for (int k = 0; k < N; ++k) {                       // forward: L y = b
        const double yk = __shfl_sync(kFull, b * dinv, k);
        b = lane == k ? yk : (lane > k ? fma(-row[k], yk, b) : b);
    }
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {                  // backward: L^T x = y
        const double xk = __shfl_sync(kFull, b * dinv, k);
        const double lkj = Lsm[k * NS + (lane < N ? lane : 0)];
        b = lane == k ? xk : (lane < k ? fma(-lkj, xk, b) : b);
    }
    return b;
}

// ig[b] = (L L^T)^-1 [lane][b]: lane j builds column j of M = L^-1 (mirrored in Msm), then G^-1 = M^T M
template <int N>
__device__ __forceinline__ void chol_inverse_fixed(const double* Lsm, double* Msm, double dinv, double (&ig)[N], int lane) {
    constexpr int NS = N | 1;
    double m[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) {
            if (k & 1) s1 = fma(Lsm[i * NS + k], m[k], s1);
            else s0 = fma(Lsm[i * NS + k], m[k], s0);
        }
        const double di = __shfl_sync(kFull, dinv, i);
        m[i] = (i == lane) ? di : (i > lane ? -(s0 + s1) * di : 0.0);
    }
    if (lane < N) {
#pragma unroll
        for (int i = 0; i < N; ++i) Msm[i * NS + lane] = m[i];
    }
    __syncwarp();
#pragma unroll
    for (int b = 0; b < N; ++b) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = b; k < N; ++k) {
            if (k & 1) s1 = fma(m[k], Msm[k * NS + b], s1);
            else s0 = fma(m[k], Msm[k * NS + b], s0);
        }
        ig[b] = s0 + s1;
    }
    __syncwarp();
}

__host__ __device__ inline int chain_order(int dim) { return dim <= 8 ? 8 : (dim <= 16 ? 16 : (dim <= 25 ? 25 : 32)); }

template <int DMAX>
__device__ __forceinline__ double matvec_regs(const double (&ig)[DMAX], int D, double x) {
    double y0 = 0.0, y1 = 0.0;
#pragma unroll
    for (int b = 0; b < DMAX; ++b) {
        if (b < D) {
            double xb = __shfl_sync(kFull, x, b);
            if (b & 1) y1 = fma(ig[b], xb, y1);
            else y0 = fma(ig[b], xb, y0);
        }
    }
    return y0 + y1;
}

// (pair a, pair b, weight) of the packed pairs this lane owns in the tensor contractions
template <int NCH>
struct PairRegs {
    int pa[NCH], pb[NCH];
    double w[NCH];
};
template <int NCH>
__device__ __forceinline__ void load_pairs(const EngineParams& P, PairRegs<NCH>& pr, int lane) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        int p = lane + 32 * ch;
        bool ok = p < P.p2;
        pr.pa[ch] = ok ? P.pair_a[p] : 0;
        pr.pb[ch] = ok ? P.pair_b[p] : 0;
        pr.w[ch] = ok ? (pr.pa[ch] == pr.pb[ch] ? 1.0 : 2.0) : 0.0;
    }
}

// out[d] = sum over packed pairs p of q[p] * T[d, a_p, b_p], for every d.  Lanes run over pairs (so a
// warp's reads of the packed T are mostly consecutive), the CTA's warps split the d range, and each
// per-d sum is reduced with a fixed xor butterfly.  q[ch] belongs to pair lane + 32 ch and already
// carries the factor 2 of off-diagonal pairs.  Ends with a CTA barrier.
template <int NCH>
__device__ __forceinline__ void tensor_contract(const double* Tsm, const double (&q)[NCH],
                                                const unsigned short* __restrict__ tidx, double* out, int D, int P2,
                                                int warp, int n_warps, int lane) {
    for (int d = warp; d < D; d += n_warps) {
        const unsigned short* row = tidx + d * P2 + lane;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            if (lane + 32 * ch < P2) {
                double t = Tsm[__ldg(row + 32 * ch)];
                if (ch & 1) a1 = fma(q[ch], t, a1);
                else a0 = fma(q[ch], t, a0);
            }
        }
        double acc = warp_sum(a0 + a1);
        if (lane == 0) out[d] = acc;
    }
    __syncthreads();
}

// out_d = u^T dG_d u (caller halves it: LastTerm, rmhmc.py:105-107 with u = G^-1 p; G^-1 symmetric)
template <int NCH>
__device__ __forceinline__ void quad_terms(const EngineParams& P, const PairRegs<NCH>& pr, const double* Tsm,
                                           const double* u, double* out, int warp, int n_warps, int lane) {
    // every warp holds u in its lanes and picks u_a, u_b by shuffle (no shared-memory gathers)
    const double ul = lane < P.dim ? u[lane] : 0.0;
    double q[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
        q[ch] = pr.w[ch] * __shfl_sync(kFull, ul, pr.pa[ch]) * __shfl_sync(kFull, ul, pr.pb[ch]);
    tensor_contract<NCH>(Tsm, q, P.tidx, out, P.dim, P.p2, warp, n_warps, lane);
}

// out_d = tr(G^-1 dG_d) = <G^-1, dG_d> (rmhmc.py:76-77,155-156)
template <int NCH>
__device__ __forceinline__ void trace_terms(const EngineParams& P, const PairRegs<NCH>& pr, const double* Tsm,
                                            const double* IGsm, double* out, int warp, int n_warps, int lane) {
    double q[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) q[ch] = pr.w[ch] * IGsm[pr.pa[ch] * P.ds + pr.pb[ch]];
    tensor_contract<NCH>(Tsm, q, P.tidx, out, P.dim, P.p2, warp, n_warps, lane);
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

// ---------------------------------------------------------------- factor kernel (one warp per chain)
// Cholesky factor, inverse and log-det of the metric built at theta_w (rmhmc.py:138,171 / :58-60).
// Writes L, G^-1 and 0.5 log|G| of the proposal slot (of slot `cur` when init != 0).
template <int N>
__global__ void __launch_bounds__(32) k_chain_factor(EngineParams P, ChainArrays S, int init) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NS = N | 1;
    const int c = blockIdx.x, lane = threadIdx.x, D = P.dim;
    if (c >= P.n_chains) return;
    if (!init && (S.iter[c] >= P.it_stop || S.nsteps[c] <= 0)) return;
    double* colbuf = reinterpret_cast<double*>(smem_raw);
    double* Lsm = colbuf + 64;
    double* Msm = Lsm + N * NS;
    const int out = init ? S.cur[c] : 1 - S.cur[c];
    double dinv, ig[N];
    {
        double lrow[N];
        load_packed_rows_pad<N>(S.g_tmp + (size_t)c * P.p2p, lrow, D, lane);
        double logdet = chol_fixed<N>(lrow, colbuf, lane, dinv);
        store_rows_fixed<N>(Lsm, lrow, lane);
        if (lane == 0) S.logdet[out * P.slot_scalar + c] = logdet;
    }
    double* ld = S.lfac + out * P.slot_invg + (size_t)c * D * D;
    for (int idx = lane; idx < D * D; idx += 32) ld[idx] = Lsm[(idx / D) * NS + (idx % D)];
    chol_inverse_fixed<N>(Lsm, Msm, dinv, ig, lane);
    double* igd = S.invg + out * P.slot_invg + (size_t)c * D * D;
#pragma unroll
    for (int b = 0; b < N; ++b)
        if (b < D && lane < D) igd[lane * D + b] = ig[b];
    if (P.matrix_free) {
        // q = packed G^-1 with doubled off-diagonals (A operand of the leverage GEMM h = KR2(X) q); lane = column b
        // of the symmetric inverse, so that for a fixed row a consecutive lanes write consecutive packed entries
        double* qp = S.qpack + (size_t)c * P.p2k;
#pragma unroll
        for (int a = 0; a < N; ++a)
            if (a < D && lane >= a && lane < D) qp[a * D - a * (a - 1) / 2 + (lane - a)] = lane == a ? ig[a] : 2.0 * ig[a];
        if (lane == 0) S.aslot[c] = out;
        if (!init) {
            // u = G_new^-1 p for the quadratic form of the explicit momentum half-step (rmhmc.py:158-161)
            const double pm = lane < D ? S.mom[(size_t)c * D + lane] : 0.0;
            const double u = matvec_regs<N>(ig, N, pm);
            if (lane < D) S.uvec[(size_t)c * D + lane] = u;
        }
    }
}

// ---------------------------------------------------------------- the per-round chain kernel
// One CTA of kTurnThreads threads per chain.
// do_back : finish the leapfrog step whose closing builds (metric at theta_w -> grad_tmp/loglik_tmp,
//           partials -> T[out], k_chain_factor -> L/G^-1/log-det[out]) have just run: traces, gradient,
//           log joint, explicit momentum half-step (R12-R14); if the trajectory is complete:
//           Hamiltonian, accept/reject, sample store (R15-R18).  init != 0: only fill slot `cur`.
// do_front: start the next leapfrog step: [new iteration: momentum draw, H_current (R2-R6)],
//           implicit momentum half-step (R7-R8), u0 and the first position iterate (R9-R10).
// Both halves run back to back in the same CTA so that T and G^-1 of the step's end point -- which
// is the next step's start point unless the proposal was rejected -- stay in shared memory.
// BIG (32 < D <= 128): 256 threads, T is read from global memory through 32-bit indices and the pair
// weights live in shared memory instead of registers.
constexpr int kTurnThreads = 128;

__host__ inline size_t turn_smem_bytes(int dim, int p2, int p3p, bool big) {
    size_t vec = big ? kMaxDimBig : 32;
    return ((size_t)dim * (dim | 1) + 8 * vec) * 8 + 8 + (big ? (size_t)p2 : (size_t)p3p) * 8;
}

template <int NCH, bool BIG>
__global__ void __launch_bounds__(BIG ? kBigThreads : kTurnThreads)
k_chain_turn(EngineParams P, ChainArrays S, int do_back, int do_front, int init) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NTHR = BIG ? kBigThreads : kTurnThreads;
    constexpr int NW = NTHR / 32;
    constexpr int VEC = BIG ? kMaxDimBig : 32;
    const int c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, D = P.dim, DS = P.ds;
    if (c >= P.n_chains) return;
    long long it = S.iter[c];
    if (!init && it >= P.it_stop) return;
    double* IG = reinterpret_cast<double*>(smem_raw);
    double* v_p = IG + D * DS;        // momentum
    double* v_u = v_p + VEC;          // G^-1 x
    double* v_out = v_u + VEC;        // contraction results
    double* v_grad = v_out + VEC;
    double* v_tr = v_grad + VEC;
    double* v_th = v_tr + VEC;
    double* v_x = v_th + VEC;         // scratch vector (z, fixed-point iterate)
    double* v_y = v_x + VEC;
    double* Tsm = v_y + VEC;          // packed T (D <= 32) or the pair weights q (BIG)
    if ((Tsm - IG) & 1) ++Tsm;
    const double* Tg = nullptr;       // BIG: packed T of the loaded slot, in global memory
    const bool live = tid < D;
    PairRegs<NCH> pr;
    if (!BIG) load_pairs<NCH>(P, pr, lane);

    int cur = S.cur[c];
    int step = init ? 0 : S.step[c];
    int have_slot = -1;               // slot whose T / G^-1 / grad / trace / theta are in shared memory

    auto load_slot_mats = [&](int slot) {
        if (BIG) {
            Tg = S.tpack + slot * P.slot_t + (size_t)c * P.p3p;
        } else {
            const double2* ts = reinterpret_cast<const double2*>(S.tpack + slot * P.slot_t + (size_t)c * P.p3p);
            double2* td = reinterpret_cast<double2*>(Tsm);
            for (int i = tid; i < P.p3p / 2; i += NTHR) td[i] = ts[i];
        }
        const double* ig = S.invg + slot * P.slot_invg + (size_t)c * D * D;
        for (int idx = tid; idx < D * D; idx += NTHR) IG[(idx / D) * DS + (idx % D)] = ig[idx];
    };
    // out_d = u^T dG_d u  /  out_d = tr(G^-1 dG_d); both end with a CTA barrier
    auto quad = [&](const double* u, double* out) {
        if (BIG) {
            for (int p = tid; p < P.p2; p += NTHR) {
                int pa = P.pair_a[p], pb = P.pair_b[p];
                Tsm[p] = (pa == pb ? 1.0 : 2.0) * u[pa] * u[pb];
            }
            __syncthreads();
            tensor_contract_big(Tg, Tsm, P.tidx32, out, D, P.p2);
        } else {
            quad_terms<NCH>(P, pr, Tsm, u, out, warp, NW, lane);
        }
    };
    auto trace = [&](double* out) {
        if (BIG) {
            for (int p = tid; p < P.p2; p += NTHR) {
                int pa = P.pair_a[p], pb = P.pair_b[p];
                Tsm[p] = (pa == pb ? 1.0 : 2.0) * IG[pa * DS + pb];
            }
            __syncthreads();
            tensor_contract_big(Tg, Tsm, P.tidx32, out, D, P.p2);
        } else {
            trace_terms<NCH>(P, pr, Tsm, IG, out, warp, NW, lane);
        }
    };
    // y = G^-1 x for vectors in shared memory; ends with a CTA barrier
    auto matvec = [&](const double* x, double* y) {
        if (live) {
            double y0 = 0.0, y1 = 0.0;
            const double* r = IG + tid * DS;
            int b = 0;
            for (; b + 1 < D; b += 2) { y0 = fma(r[b], x[b], y0); y1 = fma(r[b + 1], x[b + 1], y1); }
            if (b < D) y0 = fma(r[b], x[b], y0);
            y[tid] = y0 + y1;
        }
        __syncthreads();
    };
    auto dot = [&](const double* x, const double* y) {     // every thread gets the same value
        if (BIG) {
            double s = 0.0;
            for (int b = 0; b < D; ++b) s = fma(x[b], y[b], s);
            return s;
        }
        return warp_sum(lane < D ? x[lane] * y[lane] : 0.0);    // D <= 32: one lane per term, fixed butterfly
    };

    if (do_back) {
        const int nsteps = init ? 1 : S.nsteps[c];
        const int out = init ? cur : 1 - cur;
        const int sgn = init ? 1 : S.dir[c];
        double hprop;
        bool finished = true;
        if (nsteps > 0) {
            // ---- R12/R13: traces with the freshly factored metric; R7': gradient; log joint
            load_slot_mats(out);
            if (live) {
                double th = S.theta_w[(size_t)c * D + tid];
                v_th[tid] = th;
                v_grad[tid] = S.grad_tmp[(size_t)c * D + tid] - th / P.alpha;                    // rmhmc.py:140
                v_p[tid] = init ? 0.0 : S.mom[(size_t)c * D + tid];
            }
            __syncthreads();
            trace(v_tr);
            const double half_log = 0.5 * log(2.0 * 3.14159265358979323846 * P.alpha);
            double lp = 0.0;
            for (int b = 0; b < D; ++b) lp += -half_log - v_th[b] * v_th[b] / (2.0 * P.alpha);
            const double ljl = S.loglik_tmp[c] + lp;                        // rmhmc.py:166-169, tools.py:10-14
            const double logdet = S.logdet[out * P.slot_scalar + c];
            if (live) {
                S.theta[out * P.slot_theta + (size_t)c * D + tid] = v_th[tid];
                S.grad[out * P.slot_theta + (size_t)c * D + tid] = v_grad[tid];
                S.trace[out * P.slot_theta + (size_t)c * D + tid] = v_tr[tid];
            }
            if (tid == 0) S.logjoint[out * P.slot_scalar + c] = ljl;
            if (init) return;
            have_slot = out;

            // ---- R14: explicit closing momentum half-step
            matvec(v_p, v_u);
            quad(v_u, v_out);
            if (live) {
                double p = v_p[tid] + (sgn * P.step_size / 2) * (v_grad[tid] - 0.5 * v_tr[tid] + 0.5 * v_out[tid]);
                v_p[tid] = p;
                S.mom[(size_t)c * D + tid] = p;
                if (P.tr_theta_steps && it < P.tr_iters)
                    P.tr_theta_steps[(((size_t)c * P.tr_iters + it) * P.n_leapfrog + step) * D + tid] = v_th[tid];
            }
            __syncthreads();
            ++step;
            if (tid == 0) ++S.leapfrogs[c];
            if (step < nsteps) {
                if (tid == 0) S.step[c] = step;
                finished = false;
            } else {
                // ---- R15: proposed Hamiltonian
                matvec(v_p, v_u);
                hprop = -ljl + logdet + 0.5 * dot(v_p, v_u);
            }
        } else {
            // empty trajectory (RandomStep = 0): the proposal is the current state
            hprop = S.hcur[c];
            if (live) v_p[tid] = S.mom[(size_t)c * D + tid];
            __syncthreads();
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
                    P.tr_mom_end[o] = v_p[tid];
                    P.tr_theta_end[o] = S.theta[(nsteps > 0 ? out : cur) * P.slot_theta + (size_t)c * D + tid];
                }
                if (tid == 0) {
                    P.tr_hprop[(size_t)c * P.tr_iters + it] = hprop;
                    P.tr_flags[(size_t)c * P.tr_iters + it] =
                        (take ? 1 : 0) | (used_u ? 2 : 0) | (sgn > 0 ? 16 : 0) | (nsteps << 8);
                }
            }
            // ---- R18: store (row it - burn_in, only for it > burn_in)
            if (P.samples && it > P.burn_in && it - P.burn_in < P.sample_cap && live)
                P.samples[((size_t)c * P.sample_cap + (it - P.burn_in)) * D + tid] =
                    S.theta[fin * P.slot_theta + (size_t)c * D + tid];
            __syncthreads();       // everyone has read the pre-update state
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
    if (have_slot != in_slot) {
        load_slot_mats(in_slot);
        if (live) {
            v_grad[tid] = S.grad[in_slot * P.slot_theta + (size_t)c * D + tid];
            v_tr[tid] = S.trace[in_slot * P.slot_theta + (size_t)c * D + tid];
            v_th[tid] = S.theta[in_slot * P.slot_theta + (size_t)c * D + tid];
        }
        __syncthreads();
    }
    int sgn, nsteps;
    if (step == 0) {
        // ---- R4-R6: p = L^T z with the current position's Cholesky factor, H_current
        if (P.ext_mom) {
            // leapfrog seam: the caller supplies p, RandomStep and TimeStep
            if (live) v_p[tid] = P.ext_mom[(size_t)c * D + tid];
            nsteps = P.ext_nsteps[c];
            sgn = P.ext_dir[c];
            __syncthreads();
        } else {
            double z = 0.0, u_step, z_dir;
            if (P.rng_mode == 0) {
                size_t row = (size_t)(it - P.tape_base) * P.n_chains + c;
                if (live) z = P.tape_z[row * D + tid];
                u_step = P.tape_u_step[row];
                z_dir = P.tape_z_dir[row];
            } else {
                if (live) z = philox_normal(P, c, it, (uint32_t)tid);
                u_step = philox_pair(P, c, it, 0x100u).u0;
                z_dir = philox_normal(P, c, it, 0x101u);
            }
            if (live) v_x[tid] = z;
            __syncthreads();
            if (live) {
                const double* lf = S.lfac + in_slot * P.slot_invg + (size_t)c * D * D;
                double p = 0.0;
                for (int i = tid; i < D; ++i) p = fma(lf[i * D + tid], v_x[i], p);     // (z L)^T = L^T z, rmhmc.py:80
                v_p[tid] = p;
            }
            __syncthreads();
            double nrm = sqrt(dot(v_p, v_p));
            if (nrm > 100.0) {                                                          // rmhmc.py:81-85
                __syncthreads();
                if (live) v_p[tid] /= nrm * 25.0;
                if (tid == 0) ++S.renorm_mom[c];
                __syncthreads();
            }
            nsteps = (int)ceil(u_step * (double)P.n_leapfrog);                          // rmhmc.py:89
            sgn = z_dir > 0.5 ? 1 : -1;                                                 // rmhmc.py:90-93
        }
        matvec(v_p, v_u);
        double hcur = -S.logjoint[in_slot * P.slot_scalar + c] + S.logdet[in_slot * P.slot_scalar + c] +
                      0.5 * dot(v_p, v_u);                                          // rmhmc.py:175-176
        if (tid == 0) {
            S.hcur[c] = hcur;
            S.nsteps[c] = nsteps;
            S.dir[c] = sgn;
        }
        if (P.tr_mom0 && it < P.tr_iters && live) P.tr_mom0[((size_t)c * P.tr_iters + it) * D + tid] = v_p[tid];
        if (P.tr_hcur && it < P.tr_iters && tid == 0) P.tr_hcur[(size_t)c * P.tr_iters + it] = hcur;
        if (nsteps <= 0) {        // u_step == 0: empty trajectory; the next back half finishes the iteration
            if (live) S.mom[(size_t)c * D + tid] = v_p[tid];
            return;
        }
    } else {
        nsteps = S.nsteps[c];
        sgn = S.dir[c];
        if (live) v_p[tid] = S.mom[(size_t)c * D + tid];
        __syncthreads();
    }

    // ---- R7/R8: implicit momentum half-step with the metric quantities of the step's start point
    const double h = sgn * P.step_size / 2;
    if (live) v_x[tid] = v_p[tid];                  // v_x: fixed-point iterate PM
    __syncthreads();
    for (int fi = 0; fi < P.n_fixed; ++fi) {
        matvec(v_x, v_u);
        quad(v_u, v_out);
        if (live) v_x[tid] = v_p[tid] + h * (v_grad[tid] - 0.5 * v_tr[tid] + 0.5 * v_out[tid]);
        __syncthreads();
    }
    // ---- R9 and the first position iterate (its metric is the one we already hold)
    matvec(v_x, v_u);
    if (live) {
        double u0 = v_u[tid];
        v_y[tid] = P.n_fixed == 0 ? v_th[tid] : v_th[tid] + h * (u0 + u0);
    }
    __syncthreads();
    double div = 1.0;
    if (P.n_fixed <= 1) {                       // this iterate is already the step's final position
        double nrm = sqrt(dot(v_y, v_y));       // rmhmc.py:125-130
        if (nrm > 10.0) {
            div = nrm * 3.0;
            if (tid == 0) ++S.renorm_pos[c];
        }
    }
    if (live) {
        S.mom[(size_t)c * D + tid] = v_x[tid];
        S.u0[(size_t)c * D + tid] = v_u[tid];
        S.theta_w[(size_t)c * D + tid] = div == 1.0 ? v_y[tid] : v_y[tid] / div;
    }
}

#ifndef RMHMC_SOLVE_CTAS
#define RMHMC_SOLVE_CTAS 16
#endif
constexpr int kSolveCtas = RMHMC_SOLVE_CTAS;
// ---------------------------------------------------------------- position fixed-point iterate (x (F-1))
// solve G(theta_w) u = p, theta_w <- theta + s eps/2 (u0 + u)   (rmhmc.py:116-122)
template <int N>
__global__ void __launch_bounds__(32, kSolveCtas) k_chain_solve(EngineParams P, ChainArrays S, int is_last) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = blockIdx.x, lane = threadIdx.x, D = P.dim;
    if (c >= P.n_chains) return;
    if (S.iter[c] >= P.it_stop || S.nsteps[c] <= 0) return;
    double* colbuf = reinterpret_cast<double*>(smem_raw);
    double* Lsm = colbuf + 64;
    const bool live = lane < D;
    const int cur = S.cur[c];
    const int in_slot = S.step[c] == 0 ? cur : 1 - cur;
    double lrow[N], dinv;
    load_packed_rows_pad<N>(S.g_tmp + (size_t)c * P.p2p, lrow, D, lane);
    double p = live ? S.mom[(size_t)c * D + lane] : 0.0;
    double w = live ? S.theta[in_slot * P.slot_theta + (size_t)c * D + lane] : 0.0;
    double u0 = live ? S.u0[(size_t)c * D + lane] : 0.0;
    chol_fixed<N>(lrow, colbuf, lane, dinv);
    store_rows_fixed<N>(Lsm, lrow, lane);
    double u = chol_solve_fixed<N>(lrow, Lsm, dinv, lane, p);                        // rmhmc.py:121
    if (P.student_t) u = (1.0 + D) * u / (1.0 + warp_sum(live ? p * u : 0.0));       // BLR_RMHMC_StudentT.m:326 (u0 is stored scaled)
    double pw = w + (S.dir[c] * P.step_size / 2) * (u0 + u);                        // rmhmc.py:122
    if (is_last && !P.student_t) pw = clamp_position(pw, lane, D, &S.renorm_pos[c]);
    if (live) S.theta_w[(size_t)c * D + lane] = pw;
}

// ---------------------------------------------------------------- D > 32: CTA-per-chain variants
__host__ inline size_t big_mat_smem_bytes(int dim, int n_mats) { return ((size_t)n_mats * dim * (dim | 1) + kMaxDimBig) * 8; }

__global__ void __launch_bounds__(kBigThreads) k_chain_factor_big(EngineParams P, ChainArrays S, int init) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = blockIdx.x, tid = threadIdx.x, D = P.dim, DS = P.ds;
    if (c >= P.n_chains) return;
    if (!init && (S.iter[c] >= P.it_stop || S.nsteps[c] <= 0)) return;
    double* A = reinterpret_cast<double*>(smem_raw);
    double* B = A + D * DS;
    const int out = init ? S.cur[c] : 1 - S.cur[c];
    unpack_sym_cta(S.g_tmp + (size_t)c * P.p2p, A, D, DS);
    double logdet = chol_cta(A, D, DS);
    if (tid == 0) S.logdet[out * P.slot_scalar + c] = logdet;
    double* ld = S.lfac + out * P.slot_invg + (size_t)c * D * D;
    for (int idx = tid; idx < D * D; idx += kBigThreads) ld[idx] = A[(idx / D) * DS + (idx % D)];
    chol_inverse_cta(A, B, D, DS);
    double* igd = S.invg + out * P.slot_invg + (size_t)c * D * D;
    for (int idx = tid; idx < D * D; idx += kBigThreads) igd[idx] = B[(idx / D) * DS + (idx % D)];
    if (P.matrix_free) {
        double* qp = S.qpack + (size_t)c * P.p2k;
        for (int p = tid; p < P.p2; p += kBigThreads) {
            int pa = P.pair_a[p], pb = P.pair_b[p];
            qp[p] = (pa == pb ? 1.0 : 2.0) * B[pa * DS + pb];
        }
        if (tid == 0) S.aslot[c] = out;
        if (!init && tid < D) {
            const double* mom = S.mom + (size_t)c * D;
            double y0 = 0.0, y1 = 0.0;
            int b = 0;
            for (; b + 1 < D; b += 2) { y0 = fma(B[b * DS + tid], mom[b], y0); y1 = fma(B[(b + 1) * DS + tid], mom[b + 1], y1); }
            if (b < D) y0 = fma(B[b * DS + tid], mom[b], y0);
            S.uvec[(size_t)c * D + tid] = y0 + y1;
        }
    }
}

__global__ void __launch_bounds__(kBigThreads) k_chain_solve_big(EngineParams P, ChainArrays S, int is_last) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = blockIdx.x, tid = threadIdx.x, D = P.dim, DS = P.ds;
    if (c >= P.n_chains) return;
    if (S.iter[c] >= P.it_stop || S.nsteps[c] <= 0) return;
    double* A = reinterpret_cast<double*>(smem_raw);
    double* b = A + D * DS;
    const int cur = S.cur[c];
    const int in_slot = S.step[c] == 0 ? cur : 1 - cur;
    if (tid < D) b[tid] = S.mom[(size_t)c * D + tid];
    unpack_sym_cta(S.g_tmp + (size_t)c * P.p2p, A, D, DS);
    chol_cta(A, D, DS);
    chol_solve_cta(A, b, D, DS);                                                     // rmhmc.py:121
    double pw = 0.0;
    if (tid < D) {
        double w = S.theta[in_slot * P.slot_theta + (size_t)c * D + tid];
        pw = w + (S.dir[c] * P.step_size / 2) * (S.u0[(size_t)c * D + tid] + b[tid]);    // rmhmc.py:122
    }
    __syncthreads();
    if (tid < D) b[tid] = pw;
    __syncthreads();
    if (is_last) {                                                                   // rmhmc.py:125-130
        double n2 = 0.0;
        for (int d = 0; d < D; ++d) n2 = fma(b[d], b[d], n2);
        double nrm = sqrt(n2);
        if (nrm > 10.0) {
            pw /= nrm * 3.0;
            if (tid == 0) ++S.renorm_pos[c];
        }
    }
    if (tid < D) S.theta_w[(size_t)c * D + tid] = pw;
}

#endif  // __CUDACC__

}  // namespace rmhmc
