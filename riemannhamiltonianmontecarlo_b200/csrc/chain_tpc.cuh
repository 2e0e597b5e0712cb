// Thread-per-chain variants of the per-chain Cholesky stages (D <= 32), used for throughput batches:
//   k_chain_solve_tpc   position fixed-point iterate: solve G(theta_w) u = p, theta_w <- theta + s eps/2 (u0 + u)
//                       (rmhmc.py:116-122, the ||w|| > 10 hack of :125-130 on the last iterate)
//   k_chain_factor_tpc  L = chol(G), G^-1, 0.5 log|G| of the new metric (rmhmc.py:138,171 / :58-60), the packed G^-1 with
//                       doubled off-diagonals for the leverage GEMM and u = G^-1 p for rmhmc.py:158-161
// Same inputs and outputs as k_chain_solve / k_chain_factor (chain_kernels.cuh), which remain the variants for small
// batches (one warp per chain: short latency, but ~2400 warp instructions per 25 x 25 solve).  Here one CTA = one warp
// = 32 chains, one THREAD per chain; the 32 packed metrics sit in shared memory element-major (tpc_core.h), filled by
// warp-cooperative 8-byte cp.async (lane = packed element, so global reads stay coalesced) and drained the same way for
// the dense outputs.  ~350 warp instructions per chain and solve.
#pragma once
#include "chain_kernels.cuh"
#include "tpc_core.h"

namespace rmhmc {

#ifndef RMHMC_TPC_TILE
#define RMHMC_TPC_TILE 8
#endif
constexpr int kTpcTile = RMHMC_TPC_TILE;          // rows per register tile (x 2 columns in the factorisation)
constexpr int kTpcMaxFillIters = (kMaxDimWarp * (kMaxDimWarp + 1) / 2 + 31) / 32;      // 17
constexpr int kTpcMaxDenseIters = kMaxDimWarp * kMaxDimWarp / 32;                      // 32

// augmented triangle + TILE pad + one D-vector (solution) + one D-vector (true diagonal of L), 33 doubles each
__host__ inline size_t tpc_smem_bytes(int dim) {
    return ((size_t)tpc_elems(dim, dim + 1) + kTpcTile + 2 * (size_t)dim) * kTpcStride * 8;
}

#ifdef __CUDACC__
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// soff[it] = shared-memory element offset (x kTpcStride) of packed element e = lane + 32 it of the D-triangle inside the
// augmented layout: (i, j) at tpc_col(j, D) + i - j  ->  tpc_col(j, D + 1) + i - j = e + j.  diag_bits: bit it set when
// e is a diagonal element.
__device__ __forceinline__ void tpc_fill_table(int (&soff)[kTpcMaxFillIters], unsigned& diag_bits, int D, int lane) {
    int j = 0, cnext = D;                           // cnext = tpc_col(j + 1, D)
    diag_bits = 0;
#pragma unroll
    for (int it = 0; it < kTpcMaxFillIters; ++it) {
        const int e = lane + 32 * it;
        while (j < D - 1 && e >= cnext) { ++j; cnext += D - j; }
        soff[it] = (e + j) * kTpcStride;
        if (e == cnext - (D - j)) diag_bits |= 1u << it;
    }
}

// packed metrics of the warp's chains (bit t of mask: chain c0 + t takes part) -> shared memory.  Lane = packed element
// (coalesced 256-byte global reads), 8-byte cp.async straight into the element-major layout; two instructions per copy.
__device__ __forceinline__ void tpc_fill(double* sm, const double* __restrict__ g_tmp, const int (&soff)[kTpcMaxFillIters],
                                         long long c0, unsigned mask, int P2, int p2p, int lane) {
    uint32_t saddr[kTpcMaxFillIters];
    const uint32_t base = smem_u32(sm);
#pragma unroll
    for (int it = 0; it < kTpcMaxFillIters; ++it) saddr[it] = base + (uint32_t)soff[it] * 8u;
    const int n_full = P2 >> 5;                     // iterations in which every lane has an element
    const bool tail = lane < (P2 & 31);
    const char* gp = reinterpret_cast<const char*>(g_tmp + (size_t)c0 * p2p + lane);
    const size_t gstride = (size_t)p2p * 8;
#pragma unroll 8
    for (int t = 0; t < 32; ++t, gp += gstride) {
        if (!((mask >> t) & 1u)) continue;
#pragma unroll
        for (int it = 0; it < kTpcMaxFillIters; ++it)
            if (it < n_full || (it == n_full && tail))
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(saddr[it] + 8u * t), "l"(gp + 256 * it));
    }
}

template <int TILE>
__global__ void __launch_bounds__(32) k_chain_solve_tpc(EngineParams P, ChainArrays S, int is_last) {
    extern __shared__ __align__(16) double tpc_sm[];
    constexpr int ST = kTpcStride;
    const int lane = threadIdx.x, D = P.dim, DL = D + 1;
    const long long c0 = (long long)blockIdx.x * 32, c = c0 + lane;
    const bool active = c < P.n_chains && S.iter[c] < P.it_stop && S.nsteps[c] > 0;
    const unsigned mask = __ballot_sync(kFull, active);
    if (!mask) return;
    double* A = tpc_sm + lane;                                       // element e of this thread's chain at A[e * ST]
    double* X = A + (size_t)(tpc_elems(D, DL) + TILE) * ST;          // solution vector
    {
        int soff[kTpcMaxFillIters];
        unsigned diag_bits;
        tpc_fill_table(soff, diag_bits, D, lane);
        tpc_fill(tpc_sm, S.g_tmp, soff, c0, mask, P.p2, P.p2p, lane);
    }
    if (active) {                                                    // right-hand side p as row D         rmhmc.py:121
        const double* pm = S.mom + (size_t)c * D;
        for (int j = 0; j < D; ++j) A[(tpc_col(j, DL) + D - j) * ST] = pm[j];
    }
    cp_async_wait_all();
    __syncwarp();
    if (!active) return;
    tpc_cholesky2<TILE>(A, D, DL);
    tpc_backsolve(A, X, D);
    const size_t cd = (size_t)c * D;
    if (P.student_t) {                                               // BLR_RMHMC_StudentT.m:326 (u0 is stored scaled)
        double pu = 0.0;
        for (int d = 0; d < D; ++d) pu = fma(S.mom[cd + d], X[d * ST], pu);
        const double sc = (1.0 + D) / (1.0 + pu);
        for (int d = 0; d < D; ++d) X[d * ST] *= sc;
    }
    const int cur = S.cur[c];
    const int in_slot = S.step[c] == 0 ? cur : 1 - cur;
    const double* th = S.theta + in_slot * P.slot_theta + cd;
    const double h = S.dir[c] * P.step_size / 2;
    double n2 = 0.0;
    for (int d = 0; d < D; ++d) {
        const double pw = th[d] + h * (S.u0[cd + d] + X[d * ST]);                            // rmhmc.py:122
        X[d * ST] = pw;
        n2 = fma(pw, pw, n2);
    }
    double div = 1.0;
    if (is_last && !P.student_t) {                                   // rmhmc.py:125-130
        const double nrm = sqrt(n2);
        if (nrm > 10.0) { div = nrm * 3.0; ++S.renorm_pos[c]; }
    }
    double* tw = S.theta_w + cd;
    if (div == 1.0) for (int d = 0; d < D; ++d) tw[d] = X[d * ST];
    else for (int d = 0; d < D; ++d) tw[d] = X[d * ST] / div;
}

template <int TILE>
__global__ void __launch_bounds__(32) k_chain_factor_tpc(EngineParams P, ChainArrays S, int init) {
    extern __shared__ __align__(16) double tpc_sm[];
    constexpr int ST = kTpcStride;
    const int lane = threadIdx.x, D = P.dim, DL = D + 1, DD = D * D;
    const long long c0 = (long long)blockIdx.x * 32, c = c0 + lane;
    const bool active = c < P.n_chains && (init || (S.iter[c] < P.it_stop && S.nsteps[c] > 0));
    const unsigned mask = __ballot_sync(kFull, active);
    if (!mask) return;
    double* A = tpc_sm + lane;
    double* X = A + (size_t)(tpc_elems(D, DL) + TILE) * ST;          // u = G^-1 p
    double* Dg = X + (size_t)D * ST;                                 // true diagonal of L
    int soff[kTpcMaxFillIters];
    unsigned diag_bits;
    tpc_fill_table(soff, diag_bits, D, lane);
    tpc_fill(tpc_sm, S.g_tmp, soff, c0, mask, P.p2, P.p2p, lane);
    const bool with_u = P.matrix_free && !init;
    const int out = active ? (init ? S.cur[c] : 1 - S.cur[c]) : 0;
    if (active) {
        const double* pm = S.mom + (size_t)c * D;
        for (int j = 0; j < D; ++j) A[(tpc_col(j, DL) + D - j) * ST] = with_u ? pm[j] : 0.0;
    }
    cp_async_wait_all();
    __syncwarp();
    if (active) {
        const double logdet = tpc_cholesky2<TILE>(A, D, DL, Dg);
        S.logdet[out * P.slot_scalar + c] = logdet;
        if (with_u) {                                                // u = G_new^-1 p                    rmhmc.py:158-161
            tpc_backsolve(A, X, D);
            double* uv = S.uvec + (size_t)c * D;
            for (int d = 0; d < D; ++d) uv[d] = X[d * ST];
        }
    }
    __syncwarp();
    // ---- dense row-major L (zero upper triangle; the diagonal comes from Dg, the triangle keeps 1 / L[j][j] for the
    // inversion), one chain at a time, lane = consecutive entries
    {
        int loff[kTpcMaxDenseIters];                                 // shared offset of entry idx = lane + 32 it; -1: zero
#pragma unroll
        for (int it = 0; it < kTpcMaxDenseIters; ++it) {
            const int idx = lane + 32 * it, r = idx / D, cc = idx - r * D;
            loff[it] = (idx < DD && cc <= r) ? (cc < r ? tpc_col(cc, DL) + r - cc : tpc_elems(D, DL) + TILE + D + r) * ST : -1;
        }
        for (int t = 0; t < 32; ++t) {
            if (!((mask >> t) & 1u)) continue;
            const int out_t = __shfl_sync(kFull, out, t);
            double* ld = S.lfac + out_t * P.slot_invg + (size_t)(c0 + t) * DD + lane;
            const double* src = tpc_sm + t;
#pragma unroll
            for (int it = 0; it < kTpcMaxDenseIters; ++it)
                if (32 * it < DD && lane + 32 * it < DD) ld[32 * it] = loff[it] >= 0 ? src[loff[it]] : 0.0;
        }
    }
    __syncwarp();
    if (active) {
        tpc_invert_lower<TILE>(A, D, DL);
        tpc_mtm_lower<TILE>(A, D, DL);
        if (P.matrix_free) S.aslot[c] = out;
    }
    __syncwarp();
    // ---- dense symmetric G^-1
    {
        int goff[kTpcMaxDenseIters];
#pragma unroll
        for (int it = 0; it < kTpcMaxDenseIters; ++it) {
            const int idx = lane + 32 * it, r = idx / D, cc = idx - r * D;
            const int lo = r < cc ? r : cc, hi = r < cc ? cc : r;
            goff[it] = idx < DD ? (tpc_col(lo, DL) + hi - lo) * ST : 0;
        }
        for (int t = 0; t < 32; ++t) {
            if (!((mask >> t) & 1u)) continue;
            const int out_t = __shfl_sync(kFull, out, t);
            double* igd = S.invg + out_t * P.slot_invg + (size_t)(c0 + t) * DD + lane;
            const double* src = tpc_sm + t;
#pragma unroll
            for (int it = 0; it < kTpcMaxDenseIters; ++it)
                if (32 * it < DD && lane + 32 * it < DD) igd[32 * it] = src[goff[it]];
        }
    }
    // ---- q = packed G^-1 with doubled off-diagonals (A operand of the leverage GEMM h = KR2(X) q)
    if (P.matrix_free) {
        for (int t = 0; t < 32; ++t) {
            if (!((mask >> t) & 1u)) continue;
            double* qp = S.qpack + (size_t)(c0 + t) * P.p2k + lane;
            const double* src = tpc_sm + t;
#pragma unroll
            for (int it = 0; it < kTpcMaxFillIters; ++it)
                if (32 * it < P.p2 && lane + 32 * it < P.p2) {
                    const double v = src[soff[it]];
                    qp[32 * it] = ((diag_bits >> it) & 1u) ? v : v + v;
                }
        }
    }
}
#endif  // __CUDACC__

}  // namespace rmhmc
