// Per-chain linear algebra for 32 < D <= 128: one CTA per chain, matrices in shared memory, one
// thread per row/column.  These paths serve the large-D configurations (BASELINE.json configs[2]:
// D = 100, configs[4]: D = 64), where a round is dominated by the O(N D^3) partials contraction
// (tens of TFLOP per round) and the O(D^3) per-chain work is noise; they are written for
// correctness and small code, not speed.
#pragma once
#include "common.cuh"

namespace rmhmc {

constexpr int kBigThreads = 256;
constexpr int kMaxDimBig = 128;

#ifdef __CUDACC__
// dense symmetric A (row stride DS) from the packed upper triangle; CTA-wide
__device__ __forceinline__ void unpack_sym_cta(const double* __restrict__ gp, double* A, int D, int DS) {
    for (int idx = threadIdx.x; idx < D * D; idx += blockDim.x) {
        int i = idx / D, j = idx - i * D;
        int lo = i < j ? i : j, hi = i < j ? j : i;
        A[i * DS + j] = gp[pair_index(lo, hi, D)];
    }
    __syncthreads();
}

// in-place lower Cholesky factor (strict upper part is zeroed); returns sum log L_kk to every thread
__device__ __forceinline__ double chol_cta(double* A, int D, int DS) {
    const int tid = threadIdx.x;
    for (int k = 0; k < D; ++k) {
        double lkk = sqrt(A[k * DS + k]);
        __syncthreads();
        if (tid == k) A[k * DS + k] = lkk;
        if (tid > k && tid < D) A[tid * DS + k] /= lkk;
        __syncthreads();
        if (tid > k && tid < D) {
            double lik = A[tid * DS + k];
            for (int j = k + 1; j <= tid; ++j) A[tid * DS + j] -= lik * A[j * DS + k];
        }
        __syncthreads();
    }
    double ld = 0.0;
    for (int k = 0; k < D; ++k) ld += log(A[k * DS + k]);
    if (tid < D)
        for (int j = tid + 1; j < D; ++j) A[tid * DS + j] = 0.0;
    __syncthreads();
    return ld;
}

// solve L L^T x = b in place on the shared vector b[0..D)
__device__ __forceinline__ void chol_solve_cta(const double* L, double* b, int D, int DS) {
    const int tid = threadIdx.x;
    for (int k = 0; k < D; ++k) {                       // forward
        if (tid == k) b[k] /= L[k * DS + k];
        __syncthreads();
        if (tid > k && tid < D) b[tid] -= L[tid * DS + k] * b[k];
        __syncthreads();
    }
    for (int k = D - 1; k >= 0; --k) {                  // backward
        if (tid == k) b[k] /= L[k * DS + k];
        __syncthreads();
        if (tid < k) b[tid] -= L[k * DS + tid] * b[k];
        __syncthreads();
    }
}

// B = (L L^T)^-1, thread j owns column j
__device__ __forceinline__ void chol_inverse_cta(const double* L, double* B, int D, int DS) {
    const int j = threadIdx.x;
    if (j < D) {
        for (int i = 0; i < D; ++i) {                   // L Y = I
            double s = (i == j) ? 1.0 : 0.0;
            for (int k = (j < i ? j : i); k < i; ++k) s -= L[i * DS + k] * B[k * DS + j];
            B[i * DS + j] = i < j ? 0.0 : s / L[i * DS + i];
        }
        for (int i = D - 1; i >= 0; --i) {              // L^T X = Y
            double s = B[i * DS + j];
            for (int k = i + 1; k < D; ++k) s -= L[k * DS + i] * B[k * DS + j];
            B[i * DS + j] = s / L[i * DS + i];
        }
    }
    __syncthreads();
}

// out[d] = sum_p q[p] * T[tidx[d * P2 + p]] with T in global memory; warps over d, lanes over p
__device__ __forceinline__ void tensor_contract_big(const double* __restrict__ T, const double* q,
                                                    const unsigned int* __restrict__ tidx, double* out, int D, int P2) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int d = warp; d < D; d += nw) {
        const unsigned int* row = tidx + (size_t)d * P2;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int p = lane;
        for (; p + 96 < P2; p += 128) {
            a0 = fma(q[p], T[__ldg(row + p)], a0);
            a1 = fma(q[p + 32], T[__ldg(row + p + 32)], a1);
            a2 = fma(q[p + 64], T[__ldg(row + p + 64)], a2);
            a3 = fma(q[p + 96], T[__ldg(row + p + 96)], a3);
        }
        for (; p < P2; p += 32) a0 = fma(q[p], T[__ldg(row + p)], a0);
        double acc = warp_sum((a0 + a1) + (a2 + a3));
        if (lane == 0) out[d] = acc;
    }
    __syncthreads();
}
#endif

}  // namespace rmhmc
