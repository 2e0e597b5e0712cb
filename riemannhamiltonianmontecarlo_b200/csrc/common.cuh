// Shared device/host helpers for the batched RMHMC kernels (sm_100a).
//
// Vocabulary (follows the reference, /root/reference/code/rmhmc.py):
//   chain      one independent Markov chain (the reference runs exactly one)
//   round      one generalized-leapfrog step (rmhmc.py:96-163) executed for every chain
//   G, Gp      Fisher metric X^T diag(v) X + I/alpha (rmhmc.py:57); Gp = packed upper triangle
//   T          all metric partials dG/dw_d stacked (rmhmc.py:64-75) -- fully symmetric in (d,a,b),
//              stored packed over i<=j<=k
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace rmhmc {

constexpr int kMaxDimWarp = 32;   // per-chain kernels map one lane to one parameter

// ---------------------------------------------------------------- packed index math
__host__ __device__ inline int pad_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline int num_pairs(int d) { return d * (d + 1) / 2; }
__host__ __device__ inline int num_triples(int d) { return d * (d + 1) * (d + 2) / 6; }

// (a <= b) -> index into the packed upper triangle, row-major by a
__host__ __device__ inline int pair_index(int a, int b, int d) {
    return a * d - a * (a - 1) / 2 + (b - a);
}
// first packed-triple index of slab i = { (i,j,k) : i <= j <= k }
__host__ __device__ inline int slab_offset(int i, int d) {
    // sum_{m<i} n_m (n_m+1)/2 with n_m = d-m  ==  T(d) - T(d-i), T(n) = n(n+1)(n+2)/6
    return num_triples(d) - num_triples(d - i);
}
// (i <= j <= k) -> packed triple index
__host__ __device__ inline int triple_index(int i, int j, int k, int d) {
    int n = d - i, jj = j - i, kk = k - i;
    return slab_offset(i, d) + jj * n - jj * (jj - 1) / 2 + (kk - jj);
}
// any order
__host__ __device__ inline int triple_index_any(int a, int b, int c, int d) {
    int lo = a < b ? a : b, hi = a < b ? b : a;
    int i, j, k;
    if (c <= lo) { i = c; j = lo; k = hi; }
    else if (c <= hi) { i = lo; j = c; k = hi; }
    else { i = lo; j = hi; k = c; }
    return triple_index(i, j, k, d);
}

// smem/global row stride (in doubles) of the staged design matrix: smallest 4*odd >= D+1.
// 4*odd keeps the 8x4 / 4x8 DMMA fragment loads bank-conflict free; the +1 leaves the last
// column free for the label t_n, which rides along with every staged row.
__host__ __device__ inline int x_stride(int d) {
    int s = pad_up(d + 1, 4);
    if ((s / 4) % 2 == 0) s += 4;
    return s;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------- FP64 tensor core
// D(8x8) += A(8x4, row) * B(4x8, col).  Fragment ownership (lane = 4*g + q):
//   a = A[g][q], b = B[q][g], c0 = C[g][2q], c1 = C[g][2q+1].   SASS: DMMA.8x8x4
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- async copies
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" :: "n"(N));
}

// mbarrier + 1-D bulk TMA (cp.async.bulk, SASS UBLKCP) for contiguous design-matrix row blocks
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#endif  // __CUDACC__

// ---------------------------------------------------------------- per-chain sampler state
// Structure-of-arrays over chains; two slots (current / proposal) so that accept is a flag flip.
struct ChainArrays {
    // slot-indexed: ptr + slot * slot_stride
    double* theta;     // [2][C][D]     position
    double* logjoint;  // [2][C]        log-likelihood + log-prior at theta
    double* lfac;      // [2][C][D*D]   lower Cholesky factor of G(theta) (dense, row-major)
    double* invg;      // [2][C][D*D]   G^-1 (dense, symmetric)
    double* logdet;    // [2][C]        sum log diag chol(G) = 0.5 log|G|
    double* tpack;     // [2][C][P3p]   packed partials tensor T(theta)
    double* trace;     // [2][C][D]     tr(G^-1 dG_d)
    double* grad;      // [2][C][D]     gradient of the log joint
    // per chain, single copy
    double* mom;       // [C][D]   momentum being integrated
    double* theta_w;   // [C][D]   working position (fixed-point iterate; input of the metric builds)
    double* u0;        // [C][D]   G(theta)^-1 p of the current leapfrog step (rmhmc.py:113)
    double* hcur;      // [C]      Hamiltonian at the start of the iteration
    double* pudot;     // [C]      PM . G^-1 PM of the momentum iterate whose u = G^-1 PM is in uvec (Student-t kinetic energy only)
    double* g_tmp;     // [C][P2p] metric at theta_w (output of a metric build)
    double* grad_tmp;  // [C][D]   X^T (t - p) at theta_w (closing build only)
    double* loglik_tmp;// [C]      log-likelihood at theta_w (closing build only)
    double* cbuf;      // [C][Np]  c_n = v_n (1 - 2 p_n) at theta_w (closing build only; materialised-partials mode)
    // matrix-free partials (mf_kernels.cuh): tr(G^-1 dG_d) and u^T dG_d u straight from the data
    double* cw;        // [2][Cpad][Np]  c_n of each slot's position (replaces the packed tensor T)
    double* hbuf;      // [Cpad][Np]     leverages h_n = x_n^T G^-1 x_n of the newest metric
    double* qpack;     // [Cpad][P2k]    packed G^-1, off-diagonal pairs doubled (A operand of the leverage GEMM)
    double* uvec;      // [C][D]         G^-1 PM: input of the next quadratic-form pass
    double* quad_tmp;  // [C][D]         sum_n c_n (x_n.u)^2 x_nd = u^T dG_d u
    double* trace_tmp; // [C][D]         sum_n c_n h_n x_nd = tr(G^-1 dG_d)   (directly after quad_tmp)
    int* aslot;        // [C]  slot whose c_n the next pass over the data reads
    int* cur;          // [C]  which slot holds the current state
    int* step;         // [C]  leapfrog steps done in the running trajectory
    int* nsteps;       // [C]  RandomStep of the running trajectory
    int* dir;          // [C]  TimeStep (+1/-1)
    long long* iter;   // [C]  MCMC iterations completed
    long long* accepted;       // [C]
    long long* leapfrogs;      // [C] leapfrog steps executed
    int* renorm_mom;   // [C]  momentum clamp events (rmhmc.py:81-85)
    int* renorm_pos;   // [C]  position clamp events (rmhmc.py:125-130)
};

}  // namespace rmhmc
