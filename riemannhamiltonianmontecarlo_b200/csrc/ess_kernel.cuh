// Batched effective sample size with the reference's exact semantics (tools.py:21-74):
// circular autocorrelation with period nFFT = nextpow2(S)+1 (a port quirk),
// Geyer's initial monotone sequence, clip at 1.  The reference goes through an FFT of every
// series; here each (chain, parameter) series is handled by one CTA that evaluates lags directly
// and stops at the first non-positive monotone Gamma pair -- nothing beyond it enters the
// estimate.  The circular aliasing of the reference is reproduced term by term:
//   acf_circ[k] = lin[k] + lin[nFFT - k]   (lin[j] = sum_t y_t y_{t+j}, 0 for j >= S).
#pragma once
#include "common.cuh"

namespace rmhmc {

constexpr int kEssThreads = 128;

#ifdef __CUDACC__
__device__ __forceinline__ double block_sum_128(double v, double* scratch) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    return scratch[0] + scratch[1] + scratch[2] + scratch[3];
}

__device__ __forceinline__ double lin_lag(const double* y, int S, int lag, double* scratch) {
    double s = 0.0;
    if (lag < S)
        for (int t = threadIdx.x; t + lag < S; t += kEssThreads) s += y[t] * y[t + lag];
    return block_sum_128(s, scratch);
}

// samples: series (c, d) element s at samples[c*chain_stride + s*row_stride + d]
// With starts/counts (ragged mode) chain c uses rows [starts[c], starts[c] + counts[c]) with
// max_lag = count - 1; S then is only the shared-memory capacity.
__global__ void __launch_bounds__(kEssThreads) k_ess(const double* __restrict__ samples, size_t chain_stride,
                                                      size_t row_stride, int S, int max_lag, int n_fft,
                                                      double* __restrict__ ess_out, int D,
                                                      const long long* __restrict__ starts,
                                                      const long long* __restrict__ counts,
                                                      double* __restrict__ gscratch = nullptr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // the centred series lives in shared memory (<= 24000 samples) or, for longer series, in a global scratch row
    double* y = gscratch ? gscratch + ((size_t)blockIdx.x * gridDim.y + blockIdx.y) * (size_t)S : reinterpret_cast<double*>(smem_raw);
    __shared__ double scratch[4];
    const int c = blockIdx.x, d = blockIdx.y;
    const double* src = samples + (size_t)c * chain_stride + d;
    if (counts) {
        long long n = counts[c];
        if (n > S) n = S;
        if (n < 2) {
            if (threadIdx.x == 0) ess_out[(size_t)c * D + d] = 0.0;
            return;
        }
        src += (size_t)starts[c] * row_stride;
        S = (int)n;
        max_lag = S - 1;
        n_fft = 1;
        while (n_fft < S) n_fft *= 2;
        n_fft += 1;
    }
    double s = 0.0;
    for (int t = threadIdx.x; t < S; t += kEssThreads) {
        double v = src[(size_t)t * row_stride];
        y[t] = v;
        s += v;
    }
    double mean = block_sum_128(s, scratch) / S;
    for (int t = threadIdx.x; t < S; t += kEssThreads) y[t] -= mean;
    __syncthreads();
    const double a0 = lin_lag(y, S, 0, scratch);        // circular lag 0: alias lag n_fft >= S is empty
    const int half = (max_lag + 1) / 2;
    double run_min = 0.0, sum_pos = 0.0;
    for (int j = 0; j < half; ++j) {
        double gam = 0.0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int k = 2 * j + e;
            double a = (k == 0) ? a0 : lin_lag(y, S, k, scratch);
            if (k > 0 && n_fft - k < S) a += lin_lag(y, S, n_fft - k, scratch);
            gam += a / a0;
        }
        run_min = (j == 0) ? gam : fmin(gam, run_min);
        if (!(run_min > 0.0)) break;       // monotone sequence: nothing after this is positive
        sum_pos += run_min;
    }
    if (threadIdx.x == 0) {
        double mono = -1.0 + 2.0 * sum_pos;              // -rho_0 + 2 sum Gamma  (rho_0 = 1)
        if (mono < 1.0) mono = 1.0;
        // a frozen series has zero variance: the reference divides 0/0 and returns NaN
        ess_out[(size_t)c * D + d] = (a0 > 0.0) ? (double)S / mono : __longlong_as_double(0x7ff8000000000000LL);
    }
}
// ---- tools.ac (tools.py:21-30): normalised circular autocorrelation, lags 0..n_lag, one series per
// blockIdx.y, one lag per blockIdx.x.  acf[series][lag] is written UN-normalised; k_acf_normalise
// divides by lag 0 afterwards.
__global__ void __launch_bounds__(kEssThreads) k_acf_raw(const double* __restrict__ series, int S, int n_fft,
                                                         int n_lag, double* __restrict__ acf) {
    __shared__ double scratch[4];
    const int k = blockIdx.x;
    const double* x = series + (size_t)blockIdx.y * S;
    double s = 0.0;
    for (int t = threadIdx.x; t < S; t += kEssThreads) s += x[t];
    const double mean = block_sum_128(s, scratch) / S;
    double a = 0.0;
    if (k < S)
        for (int t = threadIdx.x; t + k < S; t += kEssThreads) a += (x[t] - mean) * (x[t + k] - mean);
    const int k2 = n_fft - k;                            // circular alias (nFFT = nextpow2(S) + 1)
    if (k > 0 && k2 < S)
        for (int t = threadIdx.x; t + k2 < S; t += kEssThreads) a += (x[t] - mean) * (x[t + k2] - mean);
    a = block_sum_128(a, scratch);
    if (threadIdx.x == 0) acf[(size_t)blockIdx.y * (n_lag + 1) + k] = a;
}
__global__ void k_acf_normalise(double* acf, int n_lag, int n_series) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_series * (n_lag + 1)) return;
    int sidx = i / (n_lag + 1), k = i - sidx * (n_lag + 1);
    double a0 = acf[(size_t)sidx * (n_lag + 1)];
    if (k > 0) acf[i] /= a0;
}
__global__ void k_acf_lag0(double* acf, int n_lag, int n_series) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_series) acf[(size_t)s * (n_lag + 1)] = acf[(size_t)s * (n_lag + 1)] / acf[(size_t)s * (n_lag + 1)];
}
// Gelman-Rubin potential scale reduction per parameter over C chains of S samples (not part of the reference;
// SURVEY.md 8c specifies the classic estimator: W = mean_c var_c (ddof = 1), B/S = var_c(mean_c) (ddof = 1),
// Rhat = sqrt(((S-1)/S W + B/S) / W)).  One CTA per parameter; thread i takes chains i, i + 128, ...; sums in fixed order.
// moments_out (optional, [3][D]): sum_c mean_c, sum_c mean_c^2, sum_c var_c -- the sufficient statistics that ranks
// holding different chains add up (rmhmc_stats_gather) before k_rhat_finish.
__global__ void __launch_bounds__(kEssThreads) k_rhat(const double* __restrict__ samples, size_t chain_stride, size_t row_stride,
                                                       int C, int S, double* __restrict__ rhat_out,
                                                       double* __restrict__ moments_out = nullptr) {
    __shared__ double sm[3][kEssThreads];
    const int d = blockIdx.x, tid = threadIdx.x;
    double s_mean = 0.0, s_mean2 = 0.0, s_var = 0.0;
    for (int c = tid; c < C; c += kEssThreads) {
        const double* y = samples + (size_t)c * chain_stride + d;
        double m = 0.0;
        for (int t = 0; t < S; ++t) m += y[(size_t)t * row_stride];
        m /= S;
        double v = 0.0;
        for (int t = 0; t < S; ++t) { double e = y[(size_t)t * row_stride] - m; v += e * e; }
        v /= (S - 1);
        s_mean += m; s_mean2 += m * m; s_var += v;
    }
    sm[0][tid] = s_mean; sm[1][tid] = s_mean2; sm[2][tid] = s_var;
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0, w = 0.0;
        for (int i = 0; i < kEssThreads; ++i) { a += sm[0][i]; b += sm[1][i]; w += sm[2][i]; }
        if (moments_out) { moments_out[d] = a; moments_out[gridDim.x + d] = b; moments_out[2 * gridDim.x + d] = w; }
        w /= C;
        const double b_over_s = (b - a * a / C) / (C - 1);
        if (rhat_out) rhat_out[d] = sqrt(((double)(S - 1) / S * w + b_over_s) / w);
    }
}
// Rhat from moments summed over all ranks: moments [3][D], c_total chains of S samples each
__global__ void k_rhat_finish(const double* __restrict__ moments, int D, double c_total, int S, double* __restrict__ rhat_out) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const double a = moments[d], b = moments[D + d], w = moments[2 * D + d] / c_total;
    const double b_over_s = (b - a * a / c_total) / (c_total - 1.0);
    rhat_out[d] = sqrt(((double)(S - 1) / S * w + b_over_s) / w);
}
// out[d] = sum_c ess[c][d] with NaN (a chain frozen over the whole window: 0/0 in tools.py:27) counted as 0; fixed order
__global__ void __launch_bounds__(kEssThreads) k_ess_colsum(const double* __restrict__ ess, long long C, int D, double* __restrict__ out) {
    __shared__ double sm[kEssThreads];
    const int d = blockIdx.x, tid = threadIdx.x;
    double s = 0.0;
    for (long long c = tid; c < C; c += kEssThreads) { const double v = ess[(size_t)c * D + d]; s += (v == v) ? v : 0.0; }
    sm[tid] = s;
    __syncthreads();
    if (tid == 0) { double a = 0.0; for (int i = 0; i < kEssThreads; ++i) a += sm[i]; out[d] = a; }
}
#endif

}  // namespace rmhmc
