// INT8-slice metric build on the Blackwell tensor cores (tcgen05.mma.kind::i8, TMEM int32 accumulators).
//
// The Fisher metric of every chain, G = X^T diag(v) X + I/alpha (rmhmc.py:51-57, :116-119, :134-137), is the
// chain-batched contraction
//     G[chains x pairs] = V[chains x rows] . KR2(X)[rows x pairs],   KR2(X)[n, (a,b)] = x_na x_nb  (a <= b).
// tcgen05 has no FP64 kind, so both operands are split into S balanced base-256 digits (an Ozaki scheme):
//     v_cn   = sA        * sum_i a_i 256^-i  (a_i in [-128, 127], i = 0..S-1; sA fixed: 0 <= v <= 1/4)
//     k_n,ab = sB[ab]    * sum_j b_j 256^-j  (sB per packed column)
// Every digit product is exact in the int32 TMEM accumulators (|a b| <= 2^14, K <= 16384 rows, <= S products per
// accumulator), products of equal weight i + j = w share accumulator `class w`, and the classes w < S are recombined
// exactly in int64 and rounded ONCE to FP64 in the epilogue.  The dropped classes (w >= S) and the operand truncation
// are ~2^-(8S-1) relative to the operand scales: measured max relative error of G 1-4e-12 for S = 5 (15 MMAs per K
// step) and 1e-14 for S = 6 (21 MMAs) on the German- / Australian-shaped data (scripts/ozaki_feasibility.py).
//
// Two launches per build:
//   k_i8_vslice   f = X theta, p = sigma(f), v = p (1 - p) in FP64 on the CUDA cores (thread = chain, X broadcast from
//                 shared memory), digits of v -> A planes [S][Cpad][Kp] int8 (K-major).  CLOSING: also X^T (t - p),
//                 log-likelihood and c_n = v (1 - 2p)  (rmhmc.py:140, :148-149, :167-168).
//   k_i8_gemm     one CTA = 128 chains x NC packed columns: TMA (tensor maps, SWIZZLE_64B) -> 3-stage shared-memory
//                 ring -> tcgen05.mma (one elected thread) -> S accumulators of NC columns in TMEM -> tcgen05.ld ->
//                 int64 recombination -> FP64 scale (+ I/alpha on the diagonal pairs) -> packed G.
// The B planes [S][columns][Kp] (digits of KR2(X)^T, 2 MB German-shaped) are formed once per data set.
#pragma once
#include <type_traits>

#include "metric_kernel.cuh"
#include "umma_common.cuh"

namespace rmhmc {

constexpr double kI8ScaleA = 0.26;        // v / sA <= 0.962: inside the balanced-digit range (|y| < 0.996)
constexpr int kI8TileM = 128;             // chains per GEMM CTA (= TMEM lanes)
constexpr int kI8BlockK = 64;             // bytes of K per pipeline stage (one SWIZZLE_64B row)
// GEMM tile shape.  WIDE (default): one CTA per SM, 128 chains x 96 columns, five accumulators fill 480 of the 512 TMEM
// columns, 12 epilogue warps.  NARROW (-DRMHMC_I8_NARROW=1): 128 x 48 tiles, 240 TMEM columns, 2-stage ring -> TWO CTAs per
// SM whose prologue / epilogue overlap the other's MMAs, N = 240 MMAs (all five KR2(X) digits of a K step in one
// instruction), 336 instead of 384 padded columns -- but 7 instead of 4 reads of the V digits from L2.  Measured
// (profiles/r02/i8_selftest_v7_n{0,1}.log, German-shaped): WIDE 0.292 / 0.146 / 0.083 / 0.043 ms at 65536 / 32768 / 16384 /
// 8192 chains, NARROW 0.316 / 0.165 / 0.091 / 0.051 ms (L2 -> SM operand traffic 3.2 GB instead of 2.35 GB per launch);
// NARROW only wins when WIDE cannot fill the SMs (Australian-shaped 4096 chains: 0.0159 vs 0.0186 ms).
#ifndef RMHMC_I8_NARROW
#define RMHMC_I8_NARROW 0
#endif
constexpr int kI8MaxRows = 16384;         // S * K * 2^14 < 2^31
constexpr int kI8VsThreads = 128;         // chains per k_i8_vslice CTA
constexpr int kI8VsRows = 32;             // rows per staged X block

template <int S> struct I8Shape {
    static_assert(S == 5 || S == 6, "5 or 6 digits");
    static constexpr bool NARROW = RMHMC_I8_NARROW != 0;
    static constexpr int NC = NARROW ? (S == 5 ? 48 : 32) : (S == 5 ? 96 : 80);      // packed columns per CTA: S * NC <= TMEM columns
    static constexpr int STAGES = NARROW ? 2 : (S == 5 ? 3 : 2);
    static constexpr int TMEM_COLS = NARROW ? 256 : 512;
    static constexpr int EPI_WARPS = NARROW ? 4 : 12;      // warp 0: TMA producer, warp 1: TMEM allocation + MMA issue, then the epilogue
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static constexpr int CTAS_PER_SM = NARROW ? 2 : 1;
    static constexpr int BITS = 8 * S - 1;                 // operand = rint(y * 2^BITS), |y| < 1
    static constexpr uint32_t A_SLICE = kI8TileM * kI8BlockK, B_SLICE = NC * kI8BlockK;
    static constexpr uint32_t STAGE_BYTES = S * (A_SLICE + B_SLICE);
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 128 /* barriers */ + NC * 16 /* column scales */;
};

// 2^52 + 2^51 + sum_k 128 * 256^k: adding it to rint-able y * 2^BITS leaves, in the low mantissa bytes, the digits
// biased by +128 (the carries of the balanced representation are done by the adder)
__host__ __device__ constexpr double i8_magic(int s) {
    double b = 0.0;
    for (int k = 0; k < s; ++k) b = b * 256.0 + 128.0;
    return 6755399441055744.0 + b;
}

struct I8GemmArgs {
    double* g_out;              // [C][P2p] packed metric
    const double2* colinfo;     // [columns] {sA sB[col] 2^(8(S-1) - 2 BITS), 1 if diagonal pair else 0}; {0, 0} for padding
    double alpha_inv;
    const double* rowscale;     // [C] per-chain scale of the A operand, or null (1): the leverage GEMM's q digits
    int n_chains, p2, p2p, k_blocks;      // p2 valid output columns, p2p = row stride of g_out (columns < p2p are stored)
    int a_rows, b_rows;         // rows per digit plane in the A / B tensor maps
    int debug_class;            // >= 0: write accumulator `class` alone (self-test)
    // split K (more than kI8MaxRows rows: the int32 accumulators hold S * 16384 * 2^14): blockIdx.z takes k-blocks
    // [z kb_split, (z + 1) kb_split) and writes its scaled partial sum to g_out + z split_stride; 0 = no split
    int kb_split;
    size_t split_stride;
};

struct I8VsArgs {
    const double* x;            // [Np][XS]
    const double* theta;        // [C][D]
    signed char* a8;            // [S][a_rows][kp]
    size_t plane_stride;        // a_rows * kp
    int kp;
    int n_chains, n_rows, n_rows_pad, dim, xs;
    // closing build
    double* grad_out; double* loglik_out; double* cbuf;
    const int* cw_cur; int cw_flip; size_t cw_slot;
};

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------ B planes
// max_n |x_na x_nb| per packed column (one CTA per column)
__global__ void k_i8_colmax(const double* __restrict__ x, const uchar2* __restrict__ pair_tab, double* __restrict__ colmax,
                            int n_rows_pad, int xs) {
    const int col = blockIdx.x;
    const uchar2 ab = pair_tab[col];
    double m = 0.0;
    for (int n = threadIdx.x; n < n_rows_pad; n += blockDim.x) m = fmax(m, fabs(x[(size_t)n * xs + ab.x] * x[(size_t)n * xs + ab.y]));
    __shared__ double red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
        colmax[col] = m;
    }
}
// digits of KR2(X)^T: b8[s][col][n], zero for col >= p2 and n >= n_rows_pad; colinfo[col]
template <int S>
__global__ void k_i8_form_b(const double* __restrict__ x, const uchar2* __restrict__ pair_tab, const double* __restrict__ colmax,
                            signed char* __restrict__ b8, double2* __restrict__ colinfo, int n_rows_pad, int xs, int p2,
                            int b_rows, int kp) {
    constexpr int BITS = I8Shape<S>::BITS;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)b_rows * kp) return;
    const int col = (int)(i / kp), n = (int)(i - (long long)col * kp);
    long long I = 0;
    double sb = 1.0;
    if (col < p2) {
        const uchar2 ab = pair_tab[col];
        const double m = colmax[col];
        sb = m > 0.0 ? m / 0.99 : 1.0;
        if (n < n_rows_pad) {
            const double k = x[(size_t)n * xs + ab.x] * x[(size_t)n * xs + ab.y];
            I = __double2ll_rn(k / sb * exp2((double)BITS));
        }
        if (n == 0) colinfo[col] = make_double2(kI8ScaleA * sb * exp2((double)(8 * (S - 1) - 2 * BITS)), ab.x == ab.y ? 1.0 : 0.0);
    } else if (n == 0) {
        colinfo[col] = make_double2(0.0, 0.0);
    }
    long long bias = 0;
#pragma unroll
    for (int k = 0; k < S; ++k) bias = bias * 256 + 128;
    const unsigned long long u = (unsigned long long)(I + bias);
#pragma unroll
    for (int s = 0; s < S; ++s)      // digit s has weight 256^(S-1-s)
        b8[((size_t)s * b_rows + col) * kp + n] = (signed char)(((u >> (8 * (S - 1 - s))) & 0xFF) ^ 0x80);
}

// ------------------------------------------------------------------------------------------------ leverage GEMM operands
// h[c][n] = x_n^T G_c^-1 x_n = sum_pairs q_c[pair] KR2(X)[n][pair]  (rmhmc.py:76-77 via the matrix-free identity): the same
// digit GEMM with (A, B, K, columns) = (q digits, KR2(X) digits by data row, packed pairs, data rows).
// max over the packed pairs of |x_na x_nb| per data row (one warp per row)
__global__ void k_i8_rowmax(const double* __restrict__ x, const uchar2* __restrict__ pair_tab, double* __restrict__ rowmax,
                            int n_rows_pad, int xs, int p2) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= n_rows_pad) return;
    double m = 0.0;
    for (int k = lane; k < p2; k += 32) { const uchar2 ab = pair_tab[k]; m = fmax(m, fabs(x[(size_t)n * xs + ab.x] * x[(size_t)n * xs + ab.y])); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) rowmax[n] = m;
}
// bl8[s][n][k]: digits of KR2(X)[n][k] / sBL[n] (zero for n >= n_rows_pad, k >= p2); colinfo_l[n] = {sBL[n] 2^(8(S-1) - 2 BITS), 0}
template <int S>
__global__ void k_i8_form_bl(const double* __restrict__ x, const uchar2* __restrict__ pair_tab, const double* __restrict__ rowmax,
                             signed char* __restrict__ bl8, double2* __restrict__ colinfo_l, int n_rows_pad, int xs, int p2,
                             int bl_rows, int kpl) {
    constexpr int BITS = I8Shape<S>::BITS;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)bl_rows * kpl) return;
    const int n = (int)(i / kpl), k = (int)(i - (long long)n * kpl);
    long long I = 0;
    if (n < n_rows_pad) {
        const double m = rowmax[n], sb = m > 0.0 ? m / 0.99 : 1.0;
        if (k < p2) {
            const uchar2 ab = pair_tab[k];
            I = __double2ll_rn(x[(size_t)n * xs + ab.x] * x[(size_t)n * xs + ab.y] / sb * exp2((double)BITS));
        }
        if (k == 0) colinfo_l[n] = make_double2(sb * exp2((double)(8 * (S - 1) - 2 * BITS)), 0.0);
    } else if (k == 0) {
        colinfo_l[n] = make_double2(0.0, 0.0);
    }
    long long bias = 0;
#pragma unroll
    for (int j = 0; j < S; ++j) bias = bias * 256 + 128;
    const unsigned long long u = (unsigned long long)(I + bias);
#pragma unroll
    for (int s = 0; s < S; ++s) bl8[((size_t)s * bl_rows + n) * kpl + k] = (signed char)(((u >> (8 * (S - 1 - s))) & 0xFF) ^ 0x80);
}
// digits of the packed inverse metric q_c (chain_kernels.cuh: qpack, off-diagonal pairs doubled) scaled by the chain's
// max |q| / 0.99 -> aq8[s][c][k], qscale[c].  One warp per chain; lane l < kpl / 16 converts 16 consecutive entries.
template <int S>
__global__ void k_i8_qdigits(const double* __restrict__ qpack, int p2, int p2k, signed char* __restrict__ aq8, size_t plane_stride,
                             int kpl, double* __restrict__ qscale, int n_chains) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= n_chains) return;
    const double* q = qpack + (size_t)c * p2k;
    double m = 0.0;
    for (int k = lane; k < p2; k += 32) m = fmax(m, fabs(q[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    const double sa = m > 0.0 && m < 1e300 ? m / 0.99 : 1.0;     // NaN / inf metric (diverged chain): digits are meaningless, h becomes NaN via qscale
    if (lane == 0) qscale[c] = (m == m && m < 1e300) ? sa : __longlong_as_double(0x7ff8000000000000LL);
    const double factor = (1.0 / sa) * (double)(1ull << 31) * (double)(1ull << (I8Shape<S>::BITS - 31));
    for (int k0 = lane * 16; k0 < kpl; k0 += 32 * 16) {
        unsigned lo[16], hi[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const double v = k0 + j < p2 ? q[k0 + j] : 0.0;
            const double vc = v == v ? fmin(fmax(v, -sa), sa) : 0.0;
            const double t = fma(vc, factor, i8_magic(S));
            lo[j] = (unsigned)__double2loint(t);
            hi[j] = (unsigned)__double2hiint(t);
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int byte = S - 1 - s;
            const unsigned b = byte & 3;
            unsigned w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned v0 = byte < 4 ? lo[4 * j] : hi[4 * j], v1 = byte < 4 ? lo[4 * j + 1] : hi[4 * j + 1];
                const unsigned v2 = byte < 4 ? lo[4 * j + 2] : hi[4 * j + 2], v3 = byte < 4 ? lo[4 * j + 3] : hi[4 * j + 3];
                w[j] = __byte_perm(__byte_perm(v0, v1, b | ((4 + b) << 4)), __byte_perm(v2, v3, b | ((4 + b) << 4)), 0x5410) ^ 0x80808080u;
            }
            *reinterpret_cast<uint4*>(aq8 + (size_t)s * plane_stride + (size_t)c * kpl + k0) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// biased digit bytes of one v: returns the S bytes u_{S-1} (least significant) .. u_0 in lo (bytes 0..3) and hi
template <int S>
__device__ __forceinline__ void i8_digits(double v, unsigned& lo, unsigned& hi) {
    const double t = fma(v, (1.0 / kI8ScaleA) * (double)(1ull << 31) * (double)(1ull << (I8Shape<S>::BITS - 31)), i8_magic(S));
    lo = (unsigned)__double2loint(t);
    hi = (unsigned)__double2hiint(t);
}

// digits of v = p (1 - p) already in HBM (32 < D: k_metric<MODE 5 / 6> writes it): vbuf[c][n_rows_pad] -> a8[s][c][kp], fixed
// scale kI8ScaleA.  One thread converts 16 consecutive rows (128 bytes in, S x 16 bytes out).
template <int S>
__global__ void k_i8_vdigits(const double* __restrict__ vbuf, int n_rows_pad, signed char* __restrict__ a8, size_t plane_stride,
                             int kp, long long n_chains) {
    const int per_chain = kp / 16;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chains * per_chain) return;
    const long long c = i / per_chain;
    const int k0 = (int)(i - c * per_chain) * 16;
    const double* v = vbuf + (size_t)c * n_rows_pad;
    unsigned lo[16], hi[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) i8_digits<S>(k0 + j < n_rows_pad ? v[k0 + j] : 0.0, lo[j], hi[j]);
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int byte = S - 1 - s;
        const unsigned b = byte & 3;
        unsigned w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned v0 = byte < 4 ? lo[4 * j] : hi[4 * j], v1 = byte < 4 ? lo[4 * j + 1] : hi[4 * j + 1];
            const unsigned v2 = byte < 4 ? lo[4 * j + 2] : hi[4 * j + 2], v3 = byte < 4 ? lo[4 * j + 3] : hi[4 * j + 3];
            w[j] = __byte_perm(__byte_perm(v0, v1, b | ((4 + b) << 4)), __byte_perm(v2, v3, b | ((4 + b) << 4)), 0x5410) ^ 0x80808080u;
        }
        *reinterpret_cast<uint4*>(a8 + (size_t)s * plane_stride + (size_t)c * kp + k0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ------------------------------------------------------------------------------------------------ A planes

// thread = chain; the CTA's 128 chains sweep the design matrix in 32-row blocks (cp.async double buffer, X rows are
// broadcast reads).  DP = parameters rounded up to the unroll bound (even).  blockIdx.y splits the rows (iterate builds).
template <int S, int DP, bool CLOSING>
__global__ void __launch_bounds__(kI8VsThreads) k_i8_vslice(I8VsArgs a) {
    constexpr int NB = kI8VsRows, GR = CLOSING ? 8 : 16;      // rows per register group
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int xs = a.xs;
    double* xbuf = reinterpret_cast<double*>(smem_raw);                   // [2][NB][xs]
    double* exp_tab = xbuf + 2 * (size_t)NB * xs;                         // [256]
    double* log_tab = exp_tab + 256;                                      // [128][2]
    const int tid = threadIdx.x;
    const int c = blockIdx.x * kI8VsThreads + tid;
    const bool live = c < a.n_chains;
    const int n_blocks_all = a.n_rows_pad / NB;
    const int rb_begin = (int)((long long)n_blocks_all * blockIdx.y / gridDim.y);
    const int n_blocks = (int)((long long)n_blocks_all * (blockIdx.y + 1) / gridDim.y) - rb_begin;

    exp_tab[tid] = exp_table_entry(tid);
    exp_tab[tid + 128] = exp_table_entry(tid + 128);
    if (CLOSING) log_table_entry(tid, log_tab[2 * tid], log_tab[2 * tid + 1]);

    double th[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) th[d] = (live && d < a.dim) ? a.theta[(size_t)c * a.dim + d] : 0.0;
    double grad[CLOSING ? DP : 1];
    double ll = 0.0;
    if (CLOSING) {
#pragma unroll
        for (int d = 0; d < DP; ++d) grad[d] = 0.0;
    }
    const int tcol = xs - 1;
    double* crow = nullptr;
    if (CLOSING && live) {
        const size_t slot = a.cw_cur ? (size_t)(a.cw_cur[c] ^ a.cw_flip) * a.cw_slot : 0;
        crow = a.cbuf + slot + (size_t)c * a.n_rows_pad;
    }
    signed char* arow = a.a8 + (size_t)c * a.kp;

    const int chunks = NB * xs / 2;       // 16-byte chunks per X block
    auto stage = [&](int rb, int buf) {
        const double* src = a.x + (size_t)(rb_begin + rb) * NB * xs;
        double* dst = xbuf + (size_t)buf * NB * xs;
        for (int i = tid; i < chunks; i += kI8VsThreads) cp_async16(dst + 2 * i, src + 2 * i);
        cp_async_commit();
    };
    stage(0, 0);
    for (int rb = 0; rb < n_blocks; ++rb) {
        if (rb + 1 < n_blocks) { stage(rb + 1, (rb + 1) & 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const double* xb = xbuf + (size_t)(rb & 1) * NB * xs;
#pragma unroll 1
        for (int r0 = 0; r0 < NB; r0 += GR) {
            double f[GR];
#pragma unroll
            for (int r = 0; r < GR; ++r) f[r] = 0.0;
#pragma unroll
            for (int dp = 0; dp < DP / 2; ++dp) {
                if (2 * dp < a.dim) {
#pragma unroll
                    for (int r = 0; r < GR; ++r) {
                        const double2 xv = *reinterpret_cast<const double2*>(xb + (size_t)(r0 + r) * xs + 2 * dp);
                        f[r] = fma(xv.x, th[2 * dp], f[r]);
                        f[r] = fma(xv.y, th[2 * dp + 1], f[r]);
                    }
                }
            }
            unsigned lo[GR], hi[GR];
            double rr[CLOSING ? GR : 1], cc[CLOSING ? GR : 1];
#pragma unroll
            for (int r4 = 0; r4 < GR; r4 += 4) {
                double ev[4], qq[4], eq[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) ev[i] = fast_exp_nonpos(-fabs(f[r4 + i]), exp_tab);
#pragma unroll
                for (int i = 0; i < 4; ++i) qq[i] = fast_rcp_1to2(1.0 + ev[i]);
#pragma unroll
                for (int i = 0; i < 4; ++i) eq[i] = ev[i] * qq[i];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = r4 + i;
                    const double vv = eq[i] * qq[i];                 // p (1 - p)
                    i8_digits<S>(vv, lo[r], hi[r]);
                    if (CLOSING) {
                        const double fv = f[r];
                        const bool pos = fv >= 0.0;
                        const double p = pos ? qq[i] : eq[i];
                        const double om = pos ? eq[i] : qq[i];       // 1 - p
                        const double t = xb[(size_t)(r0 + r) * xs + tcol];
                        const bool ovf = fv > 709.782712893384;      // the reference's exp(f) overflows: NaN gradient, -inf log-likelihood
                        rr[r] = ovf ? __longlong_as_double(0x7ff8000000000000LL) : t - p;
                        cc[r] = vv * (om - p);
                        const int row = (rb_begin + rb) * NB + r0 + r;
                        if (row < a.n_rows) {
                            const double l1pe = ovf ? __longlong_as_double(0x7ff0000000000000LL) : fmax(fv, 0.0) + fast_log1p_01(ev[i], log_tab);
                            ll += t * fv - l1pe;
                        }
                    }
                }
            }
            if (CLOSING) {
#pragma unroll
                for (int dp = 0; dp < DP / 2; ++dp) {
                    if (2 * dp < a.dim) {
#pragma unroll
                        for (int r = 0; r < GR; ++r) {
                            const double2 xv = *reinterpret_cast<const double2*>(xb + (size_t)(r0 + r) * xs + 2 * dp);
                            grad[2 * dp] = fma(rr[r], xv.x, grad[2 * dp]);
                            grad[2 * dp + 1] = fma(rr[r], xv.y, grad[2 * dp + 1]);
                        }
                    }
                }
            }
            if (live) {
                const int row0 = (rb_begin + rb) * NB + r0;
                // digit s (weight 256^(S-1-s)) is byte S-1-s of (hi:lo); bytes of 4 consecutive rows -> one word
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const int byte = S - 1 - s;
                    unsigned w[GR / 4];
#pragma unroll
                    for (int j = 0; j < GR / 4; ++j) {
                        unsigned v0, v1, v2, v3;
                        if (byte < 4) { v0 = lo[4 * j]; v1 = lo[4 * j + 1]; v2 = lo[4 * j + 2]; v3 = lo[4 * j + 3]; }
                        else { v0 = hi[4 * j]; v1 = hi[4 * j + 1]; v2 = hi[4 * j + 2]; v3 = hi[4 * j + 3]; }
                        const unsigned b = byte & 3;
                        const unsigned p01 = __byte_perm(v0, v1, b | ((4 + b) << 4));            // bytes: v0[b], v1[b]
                        const unsigned p23 = __byte_perm(v2, v3, b | ((4 + b) << 4));
                        w[j] = __byte_perm(p01, p23, 0x5410) ^ 0x80808080u;
                    }
                    signed char* dst = arow + (size_t)s * a.plane_stride + row0;
                    if constexpr (GR == 16) *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                    else *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
                }
                if (CLOSING) {
#pragma unroll
                    for (int r = 0; r < GR; r += 2) *reinterpret_cast<double2*>(crow + row0 + r) = make_double2(cc[r], cc[r + 1]);
                }
            }
        }
        __syncthreads();
    }
    if (CLOSING && live) {
#pragma unroll
        for (int d = 0; d < DP; ++d)
            if (d < a.dim) a.grad_out[(size_t)c * a.dim + d] = grad[d];
        a.loglik_out[c] = ll;
    }
}

// Position-iterate variant on the FP64 tensor path: f^T = Theta . X^T as DMMA.8x8x4 (Theta fragments in registers, X
// fragments straight from global memory / L1: X is 200 KB and every CTA sweeps it once), the logistic terms on the
// 8 (chain, row) pairs of each lane, digits through a warp-private shared-memory transpose so that the planes are
// written as 16-byte row runs.  One warp = 32 chains x 16 rows per step; no CTA-wide synchronisation.
constexpr int kI8VmWarps = 8;
// MT = m-tiles of 8 chains per warp: 4 (32 chains per CTA row, 128 registers, 16 warps per SM) or 2 (16 chains, Theta
// fragments and logistic temporaries halve: 24 warps per SM; every X fragment then feeds 2 instead of 4 DMMAs)
template <int S, int KS, int MT>
__global__ void __launch_bounds__(kI8VmWarps * 32, MT == 4 ? 2 : 3) k_i8_vslice_mma(I8VsArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* exp_tab = reinterpret_cast<double*>(smem_raw);                                   // [256]
    unsigned char* out_all = smem_raw + 256 * 8;                                             // [warps][S][8 MT][16]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int chain0 = blockIdx.x * (MT * 8);
    const int xs = a.xs;
    exp_tab[tid] = exp_table_entry(tid);
    __syncthreads();
    unsigned char* out_w = out_all + (size_t)warp * S * (MT * 8) * 16;

    // Theta fragments: a[m][ks] = theta[chain0 + 8 m + g][4 ks + q]
    double th[MT][KS];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        const int c = chain0 + m * 8 + g;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int d = ks * 4 + q;
            th[m][ks] = (c < a.n_chains && d < a.dim) ? a.theta[(size_t)c * a.dim + d] : 0.0;
        }
    }
    // rows of this CTA (blockIdx.y splits them in 16-row units), 16 rows per warp and step
    const int units_all = a.n_rows_pad / 16;
    const int u_begin = (int)((long long)units_all * blockIdx.y / gridDim.y);
    const int u_end = (int)((long long)units_all * (blockIdx.y + 1) / gridDim.y);
    auto load_b = [&](int row0, double (&b)[KS]) {
        const double* xr = a.x + (size_t)(row0 + g) * xs + q;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) b[ks] = ks * 4 < xs ? __ldg(xr + ks * 4) : 0.0;
    };
    double bnext[KS];
    int u = u_begin + warp;
    if (u < u_end) load_b(u * 16, bnext);
    for (; u < u_end; u += kI8VmWarps) {
        const int row0 = u * 16;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            double b[KS];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) b[ks] = bnext[ks];
            // prefetch the next 8-row tile of this warp
            if (nt == 0) load_b(row0 + 8, bnext);
            else if (u + kI8VmWarps < u_end) load_b((u + kI8VmWarps) * 16, bnext);
            double f[MT][2];
#pragma unroll
            for (int m = 0; m < MT; ++m) f[m][0] = f[m][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int m = 0; m < MT; ++m) dmma884(f[m][0], f[m][1], th[m][ks], b[ks]);
#pragma unroll
            for (int half = 0; half < MT / 2; ++half) {
                double ev[4], qq[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) ev[i] = fast_exp_nonpos(-fabs(f[2 * half + (i >> 1)][i & 1]), exp_tab);
#pragma unroll
                for (int i = 0; i < 4; ++i) qq[i] = fast_rcp_1to2(1.0 + ev[i]);
#pragma unroll
                for (int mm = 0; mm < 2; ++mm) {
                    const int m = 2 * half + mm;
                    unsigned lo0, hi0, lo1, hi1;
                    i8_digits<S>(ev[2 * mm] * qq[2 * mm] * qq[2 * mm], lo0, hi0);                // v = e / (1 + e)^2
                    i8_digits<S>(ev[2 * mm + 1] * qq[2 * mm + 1] * qq[2 * mm + 1], lo1, hi1);
                    unsigned char* dst = out_w + (size_t)(m * 8 + g) * 16 + nt * 8 + 2 * q;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const int byte = S - 1 - s;
                        const unsigned bsel = byte & 3;
                        const unsigned pr = __byte_perm(byte < 4 ? lo0 : hi0, byte < 4 ? lo1 : hi1, bsel | ((4 + bsel) << 4));
                        *reinterpret_cast<unsigned short*>(dst + (size_t)s * (MT * 8) * 16) = (unsigned short)((pr & 0xFFFFu) ^ 0x8080u);
                    }
                }
            }
        }
        __syncwarp();
        const int c = chain0 + lane;
        if (lane < MT * 8 && c < a.n_chains) {
#pragma unroll
            for (int s = 0; s < S; ++s)
                *reinterpret_cast<uint4*>(a.a8 + (size_t)s * a.plane_stride + (size_t)c * a.kp + row0) =
                    *reinterpret_cast<const uint4*>(out_w + ((size_t)s * (MT * 8) + lane) * 16);
        }
        __syncwarp();
    }
}

// Closing variant on the FP64 tensor path: as above for 16 chains per CTA (two m-tiles, so that the gradient accumulators fit
// the register file), plus X^T (t - p) as a second DMMA contraction (R's accumulator fragment -> A fragment by four
// shuffles, as in pass_kernel.cuh), the log-likelihood and c_n.  Every warp sweeps its own rows; the warps' partial
// gradients / log-likelihoods are added in warp order at the end (fixed order: bit-reproducible).
constexpr int kI8VcWarps = 6;      // 2 CTAs x 6 warps per SM at <= 168 registers (gradient accumulators + Theta fragments)
template <int S, int KS>
__global__ void __launch_bounds__(kI8VcWarps * 32, 2) k_i8_vslice_mma_closing(I8VsArgs a) {
    constexpr int MT = 2, DT = (KS + 1) / 2;             // m-tiles of 8 chains; parameter tiles of 8
    constexpr unsigned kFull = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* exp_tab = reinterpret_cast<double*>(smem_raw);                                   // [256]
    double* log_tab = exp_tab + 256;                                                         // [128][2]
    double* red = log_tab + 256;                                                             // [warps][16][DT * 8 + 1]
    unsigned char* out_all = reinterpret_cast<unsigned char*>(red + kI8VcWarps * 16 * (DT * 8 + 1));   // [warps][S][16][16]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int chain0 = blockIdx.x * (MT * 8);
    const int xs = a.xs, tcol = xs - 1;
    for (int i = tid; i < 256; i += kI8VcWarps * 32) exp_tab[i] = exp_table_entry(i);
    if (tid < 128) log_table_entry(tid, log_tab[2 * tid], log_tab[2 * tid + 1]);
    __syncthreads();
    unsigned char* out_w = out_all + (size_t)warp * S * 16 * 16;

    double th[MT][KS];
    double* crow[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        const int c = chain0 + m * 8 + g;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int d = ks * 4 + q;
            th[m][ks] = (c < a.n_chains && d < a.dim) ? a.theta[(size_t)c * a.dim + d] : 0.0;
        }
        crow[m] = nullptr;
        if (c < a.n_chains) {
            const size_t slot = a.cw_cur ? (size_t)(a.cw_cur[c] ^ a.cw_flip) * a.cw_slot : 0;
            crow[m] = a.cbuf + slot + (size_t)c * a.n_rows_pad;
        }
    }
    double gacc[MT][DT][2], ll[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        ll[m] = 0.0;
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) gacc[m][dt][0] = gacc[m][dt][1] = 0.0;
    }
    const int src0 = g * 4 + (q >> 1), src1 = src0 + 2;      // shuffle sources of the C -> A fragment conversion
    const bool odd = q & 1;
    const int units = a.n_rows_pad / 16;
    auto load_b = [&](int row0, double (&b)[KS]) {
        const double* xr = a.x + (size_t)(row0 + g) * xs + q;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) b[ks] = ks * 4 < xs ? __ldg(xr + ks * 4) : 0.0;
    };
    double bnext[KS];
    int u = warp;
    if (u < units) load_b(u * 16, bnext);
    for (; u < units; u += kI8VcWarps) {
        const int row0 = u * 16;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int rt = row0 + nt * 8;                  // first row of this 8-row tile
            double b[KS];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) b[ks] = bnext[ks];
            if (nt == 0) load_b(row0 + 8, bnext);
            else if (u + kI8VcWarps < units) load_b((u + kI8VcWarps) * 16, bnext);
            // labels of this lane's two rows and the gradient contraction's X fragments: X[rt + 4 kk + q][8 dt + g]
            const double t0 = __ldg(a.x + (size_t)(rt + 2 * q) * xs + tcol), t1 = __ldg(a.x + (size_t)(rt + 2 * q + 1) * xs + tcol);
            double xg[2][DT];
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                for (int dt = 0; dt < DT; ++dt) {
                    const int dcol = dt * 8 + g;
                    xg[kk][dt] = dcol < a.dim ? __ldg(a.x + (size_t)(rt + kk * 4 + q) * xs + dcol) : 0.0;
                }
            double f[MT][2];
#pragma unroll
            for (int m = 0; m < MT; ++m) f[m][0] = f[m][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int m = 0; m < MT; ++m) dmma884(f[m][0], f[m][1], th[m][ks], b[ks]);
            double ev[4], qq[4], eq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) ev[i] = fast_exp_nonpos(-fabs(f[i >> 1][i & 1]), exp_tab);
#pragma unroll
            for (int i = 0; i < 4; ++i) qq[i] = fast_rcp_1to2(1.0 + ev[i]);
#pragma unroll
            for (int i = 0; i < 4; ++i) eq[i] = ev[i] * qq[i];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                double vv[2], rr[2], cc[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int i = 2 * m + j;
                    const double fv = f[m][j];
                    const bool pos = fv >= 0.0;
                    const double p = pos ? qq[i] : eq[i];
                    const double om = pos ? eq[i] : qq[i];       // 1 - p
                    vv[j] = eq[i] * qq[i];                       // p (1 - p)
                    const double t = j ? t1 : t0;
                    const bool ovf = fv > 709.782712893384;      // the reference's exp(f) overflows: NaN gradient, -inf log-likelihood
                    rr[j] = ovf ? __longlong_as_double(0x7ff8000000000000LL) : t - p;
                    cc[j] = vv[j] * (om - p);
                    if (rt + 2 * q + j < a.n_rows) {
                        const double l1pe = ovf ? __longlong_as_double(0x7ff0000000000000LL) : fmax(fv, 0.0) + fast_log1p_01(ev[i], log_tab);
                        ll[m] += t * fv - l1pe;
                    }
                }
                if (crow[m]) *reinterpret_cast<double2*>(crow[m] + rt + 2 * q) = make_double2(cc[0], cc[1]);
                unsigned lo0, hi0, lo1, hi1;
                i8_digits<S>(vv[0], lo0, hi0);
                i8_digits<S>(vv[1], lo1, hi1);
                unsigned char* dst = out_w + (size_t)(m * 8 + g) * 16 + nt * 8 + 2 * q;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const int byte = S - 1 - s;
                    const unsigned bsel = byte & 3;
                    const unsigned pr = __byte_perm(byte < 4 ? lo0 : hi0, byte < 4 ? lo1 : hi1, bsel | ((4 + bsel) << 4));
                    *reinterpret_cast<unsigned short*>(dst + (size_t)s * 16 * 16) = (unsigned short)((pr & 0xFFFFu) ^ 0x8080u);
                }
                // X^T (t - p): R (chain g; rows 2q, 2q+1) -> A fragments (chain g; row q) of the tile's two 4-row k-steps
                const double e0 = __shfl_sync(kFull, rr[0], src0), o0 = __shfl_sync(kFull, rr[1], src0);
                const double e1 = __shfl_sync(kFull, rr[0], src1), o1 = __shfl_sync(kFull, rr[1], src1);
                const double ar0 = odd ? o0 : e0, ar1 = odd ? o1 : e1;
#pragma unroll
                for (int dt = 0; dt < DT; ++dt) {
                    dmma884(gacc[m][dt][0], gacc[m][dt][1], ar0, xg[0][dt]);
                    dmma884(gacc[m][dt][0], gacc[m][dt][1], ar1, xg[1][dt]);
                }
            }
        }
        __syncwarp();
        if (lane < MT * 8) {
            const int c = chain0 + lane;
            if (c < a.n_chains) {
#pragma unroll
                for (int s = 0; s < S; ++s)
                    *reinterpret_cast<uint4*>(a.a8 + (size_t)s * a.plane_stride + (size_t)c * a.kp + row0) =
                        *reinterpret_cast<const uint4*>(out_w + ((size_t)s * 16 + lane) * 16);
            }
        }
        __syncwarp();
    }
    // ---- the warps' partial gradients (accumulator layout: chain 8 m + g, parameters 8 dt + 2q, +1) and log-likelihoods
    constexpr int RS = DT * 8 + 1;
    double* my = red + (size_t)warp * 16 * RS;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) {
            my[(m * 8 + g) * RS + dt * 8 + 2 * q] = gacc[m][dt][0];
            my[(m * 8 + g) * RS + dt * 8 + 2 * q + 1] = gacc[m][dt][1];
        }
        double v = ll[m];
        v += __shfl_xor_sync(kFull, v, 1);
        v += __shfl_xor_sync(kFull, v, 2);
        if (q == 0) my[(m * 8 + g) * RS + DT * 8] = v;
    }
    __syncthreads();
    for (int i = tid; i < 16 * RS; i += kI8VcWarps * 32) {
        const int ci = i / RS, d = i - ci * RS, c = chain0 + ci;
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < kI8VcWarps; ++w) sum += red[(size_t)w * 16 * RS + i];
        if (c < a.n_chains) {
            if (d < a.dim) a.grad_out[(size_t)c * a.dim + d] = sum;
            else if (d == DT * 8) a.loglik_out[c] = sum;
        }
    }
}

// ------------------------------------------------------------------------------------------------ the GEMM
template <int S>
__global__ void __launch_bounds__(I8Shape<S>::THREADS, I8Shape<S>::CTAS_PER_SM) k_i8_gemm(const __grid_constant__ CUtensorMap map_a,
                                                               const __grid_constant__ CUtensorMap map_b, I8GemmArgs a) {
    using Sh = I8Shape<S>;
    constexpr int NC = Sh::NC, ST = Sh::STAGES;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)ST * Sh::STAGE_BYTES);
    uint64_t* full = bars;              // [ST] TMA -> MMA
    uint64_t* empty = bars + ST;        // [ST] MMA -> TMA (tcgen05.commit)
    uint64_t* acc_full = bars + 2 * ST; // MMA -> epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ST + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * NC;                 // first packed column
    const int m0 = blockIdx.y * kI8TileM;           // first chain
    const int kb0 = a.kb_split > 0 ? (int)blockIdx.z * a.kb_split : 0;                       // first k-block of this CTA
    const int nkb = a.kb_split > 0 ? (a.k_blocks - kb0 < a.kb_split ? a.k_blocks - kb0 : a.kb_split) : a.k_blocks;

    // one stage of operands: S digit tiles of V (128 chains x 64 bytes of K) and S of KR2(X) (NC columns x 64 bytes)
    auto issue_stage = [&](int kb) {
        const int st = kb % ST;
        unsigned char* sa = base + (size_t)st * Sh::STAGE_BYTES;
        unsigned char* sb = sa + S * Sh::A_SLICE;
        mbar_expect_tx(&full[st], Sh::STAGE_BYTES);
#pragma unroll
        for (int s = 0; s < S; ++s) tma_load_2d(sa + s * Sh::A_SLICE, &map_a, (kb0 + kb) * kI8BlockK, s * a.a_rows + m0, &full[st]);
#pragma unroll
        for (int s = 0; s < S; ++s) tma_load_2d(sb + s * Sh::B_SLICE, &map_b, (kb0 + kb) * kI8BlockK, s * a.b_rows + n0, &full[st]);
    };
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        mbar_fence_init();
        // the first stages need neither TMEM nor the other warps: their latency overlaps the allocation and the CTA barrier
        for (int kb = 0; kb < ST && kb < nkb; ++kb) issue_stage(kb);
    }
    if (warp == 1) tmem_alloc(tmem_slot, Sh::TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = ST; kb < nkb; ++kb) {
                mbar_wait_or_trap(&empty[kb % ST], (uint32_t)(((kb / ST) - 1) & 1));
                issue_stage(kb);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // Digit i of V times digits j = 0 .. S-1-i of KR2(X) go to the classes i .. S-1, which are adjacent TMEM
            // column blocks, and the B digit tiles are adjacent in shared memory: one MMA of N = MB * NC columns covers
            // MB consecutive digits j (fewer, wider MMAs: the A tile is fetched 9 instead of 15 times per K step).
            constexpr int MB = 256 / NC;          // digit tiles per MMA (N <= 256)
            // (Measured and rejected: issuing the mostly-padding last column chunk -- German-shaped: 325 = 3 x 96 + 37 -- as 15
            // un-merged MMAs of N = 48 per K step instead of 9 of N = 96 / 192.  Half the tensor work, but 0.32 -> 0.35 ms
            // per launch: narrow MMAs cost almost as much as wide ones.)
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % ST;
                mbar_wait_or_trap(&full[st], (uint32_t)((kb / ST) & 1));
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(base + (size_t)st * Sh::STAGE_BYTES);
                const uint32_t sb = sa + S * Sh::A_SLICE;
#pragma unroll
                for (int ks = 0; ks < kI8BlockK / 32; ++ks) {
#pragma unroll
                    for (int i = 0; i < S; ++i) {
                        const uint64_t da = umma_desc_k_sw64(sa + i * Sh::A_SLICE + ks * 32);
#pragma unroll
                        for (int j0 = 0; j0 < S - i; j0 += MB) {
                            const int take = S - i - j0 < MB ? S - i - j0 : MB;
                            const uint64_t db = umma_desc_k_sw64(sb + j0 * Sh::B_SLICE + ks * 32);
                            umma_i8_ss(tmem_base + (i + j0) * NC, da, db, umma_idesc_s8(kI8TileM, take * NC),
                                       (kb | ks | i) != 0 ? 1u : 0u);
                        }
                    }
                }
                umma_commit(&empty[st]);          // frees the stage when its MMAs have read it
            }
            umma_commit(acc_full);
        }
    } else {
        // ---- epilogue (12 warps): warp w reads TMEM lanes 32 (w % 4) .. + 31 = chains m0 + 32 (w % 4) + lane, the three
        // warps of a lane quarter split the column groups.  int32 classes -> FP64 exactly (magic-number conversion on
        // the integer pipe + one DADD), classes recombined with exact FMAs and ONE rounding, scaled, staged through the
        // (now idle) operand ring so that the global stores are contiguous 256-byte row segments.
        const int ew = warp - 2, quarter = warp & 3, third = ew >> 2;
        constexpr int WPQ = Sh::EPI_WARPS / 4;                  // warps per TMEM lane quarter
        constexpr int G0 = (NC / 16 + WPQ - 1) / WPQ;           // column groups (of 16) per warp
        const int cg_begin = third * G0, cg_end = cg_begin + G0 < NC / 16 ? cg_begin + G0 : NC / 16;
        const int ncols = (cg_end - cg_begin) * 16;
        constexpr int OS = G0 * 16 + 1;                         // staging row stride (doubles): odd -> conflict-free
        double2* ci_s = reinterpret_cast<double2*>(bars + 16);              // [NC] column scales of this CTA
        for (int i = threadIdx.x - 64; i < NC; i += Sh::THREADS - 64) ci_s[i] = a.colinfo[n0 + i];
        asm volatile("bar.sync 1, %0;" ::"n"(Sh::THREADS - 64) : "memory");
        mbar_wait_or_trap(acc_full, 0);          // all MMAs done: accumulators final, the operand ring is free
        tcgen05_fence_after();
        double* out_s = reinterpret_cast<double*>(base) + (size_t)ew * 32 * OS;
        double* g_dst = a.g_out + (size_t)blockIdx.z * a.split_stride;
        const double alpha_inv = blockIdx.z == 0 ? a.alpha_inv : 0.0;
        const int c_lane = m0 + quarter * 32 + lane;
        const double rs = (a.rowscale && c_lane < a.n_chains) ? a.rowscale[c_lane] : 1.0;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
        for (int cg = cg_begin; cg < cg_end; ++cg) {
            if (n0 + cg * 16 >= a.p2p) continue;             // padding beyond the stored columns of the last chunk
            uint32_t r[S][16];
#pragma unroll
            for (int w = 0; w < S; ++w) tmem_ld16(lane_base + w * NC + cg * 16, r[w]);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                double x[S];
#pragma unroll
                for (int w = 0; w < S; ++w)      // (double)(int32): 2^52 + 2^31 + x is exact, subtract the offset
                    x[w] = __hiloint2double(0x43300000, (int)(r[w][j] ^ 0x80000000u)) - 4503601774854144.0;
                // sum_w 256^(S-1-w) x_w: upper three and lower (S - 3) classes are exact in FP64 (< 2^44), one FMA joins them
                double hi = fma(fma(x[0], 256.0, x[1]), 256.0, x[2]);
                double lo = x[3];
#pragma unroll
                for (int w = 4; w < S; ++w) lo = fma(lo, 256.0, x[w]);
                double t = fma(hi, S == 5 ? 65536.0 : 16777216.0, lo);
                if (a.debug_class >= 0) {
#pragma unroll
                    for (int w = 0; w < S; ++w) if (w == a.debug_class) t = x[w];
                }
                const double2 ci = ci_s[cg * 16 + j];
                out_s[lane * OS + (cg - cg_begin) * 16 + j] = n0 + cg * 16 + j < a.p2 ? fma(t * rs, ci.x, ci.y * alpha_inv) : 0.0;
            }
        }
        __syncwarp();
        const int col0 = n0 + cg_begin * 16;
        auto copy_out = [&](auto n_tag) {
            constexpr int n = decltype(n_tag)::value;
            for (int idx = lane; idx < 32 * n; idx += 32) {
                const int row = idx / n, col = idx - row * n;
                const int c = m0 + quarter * 32 + row;
                if (c < a.n_chains && col0 + col < a.p2p) g_dst[(size_t)c * a.p2p + col0 + col] = out_s[row * OS + col];
            }
        };
        if (ncols == 48) copy_out(std::integral_constant<int, 48>{});
        else if (ncols == 32) copy_out(std::integral_constant<int, 32>{});
        else if (ncols == 16) copy_out(std::integral_constant<int, 16>{});
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, Sh::TMEM_COLS);
    }
}
#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------------ host side
struct I8Planes {               // digit planes of one operand + its tensor map
    signed char* ptr = nullptr;
    int rows = 0;               // rows per digit plane
    CUtensorMap map;
};

inline int i8_kp(int n_rows_pad) { return pad_up(n_rows_pad, kI8BlockK); }
inline int i8_dp(int dim) { return dim <= 8 ? 8 : (dim <= 16 ? 16 : (dim <= 26 ? 26 : 32)); }
template <int S> inline int i8_chunks(int p2) { return (p2 + I8Shape<S>::NC - 1) / I8Shape<S>::NC; }
inline size_t i8_vslice_smem(int xs) { return (size_t)2 * kI8VsRows * xs * 8 + 256 * 8 + 256 * 8; }

#ifdef __CUDACC__
inline size_t i8_vslice_mma_smem(int s, int mt) { return 256 * 8 + (size_t)kI8VmWarps * s * (mt * 8) * 16; }
// 16 chains per warp (MT = 2) unless RMHMC_VSLICE_MT=4
inline int i8_vslice_mt() {
    static const int mt = [] { const char* e = getenv("RMHMC_VSLICE_MT"); return e && atoi(e) == 4 ? 4 : 2; }();
    return mt;
}
// position-iterate builds: DMMA variant (KS = k-steps of 4 parameters)
template <int S, int MT>
inline cudaError_t i8_launch_vslice_mma_mt(const I8VsArgs& a, cudaStream_t stream) {
    const unsigned gx = (unsigned)((a.n_chains + MT * 8 - 1) / (MT * 8));
    unsigned gy = 1;
    const int units = a.n_rows_pad / 16;
    // few chains: split the rows too, until the grid is at least ~4 waves of the resident CTAs (8192 chains x 16 per CTA
    // = 512 CTAs on 444 slots would be 1.15 waves)
    const unsigned resident = 148u * (MT == 4 ? 2 : 3);
    while (gx * gy < 4 * resident && (int)gy * 2 * kI8VmWarps <= units) gy *= 2;
    const dim3 grid(gx, gy);
    const size_t smem = i8_vslice_mma_smem(S, MT);
    const int ks = (a.dim + 3) / 4;
    if (ks <= 2) k_i8_vslice_mma<S, 2, MT><<<grid, kI8VmWarps * 32, smem, stream>>>(a);
    else if (ks <= 4) k_i8_vslice_mma<S, 4, MT><<<grid, kI8VmWarps * 32, smem, stream>>>(a);
    else if (ks <= 7) k_i8_vslice_mma<S, 7, MT><<<grid, kI8VmWarps * 32, smem, stream>>>(a);
    else k_i8_vslice_mma<S, 8, MT><<<grid, kI8VmWarps * 32, smem, stream>>>(a);
    return cudaGetLastError();
}
template <int S>
inline cudaError_t i8_launch_vslice_mma(const I8VsArgs& a, cudaStream_t stream) {
    return i8_vslice_mt() == 4 ? i8_launch_vslice_mma_mt<S, 4>(a, stream) : i8_launch_vslice_mma_mt<S, 2>(a, stream);
}
template <int KS> inline size_t i8_vslice_mma_closing_smem(int s) {
    return (size_t)(256 + 256 + kI8VcWarps * 16 * ((KS + 1) / 2 * 8 + 1)) * 8 + (size_t)kI8VcWarps * s * 16 * 16;
}
// closing builds: DMMA variant, 16 chains per CTA
template <int S>
inline cudaError_t i8_launch_vslice_mma_closing(const I8VsArgs& a, cudaStream_t stream) {
    const unsigned gx = (unsigned)((a.n_chains + 15) / 16);
    const int ks = (a.dim + 3) / 4;
    if (ks <= 2) k_i8_vslice_mma_closing<S, 2><<<gx, kI8VcWarps * 32, i8_vslice_mma_closing_smem<2>(S), stream>>>(a);
    else if (ks <= 4) k_i8_vslice_mma_closing<S, 4><<<gx, kI8VcWarps * 32, i8_vslice_mma_closing_smem<4>(S), stream>>>(a);
    else if (ks <= 7) k_i8_vslice_mma_closing<S, 7><<<gx, kI8VcWarps * 32, i8_vslice_mma_closing_smem<7>(S), stream>>>(a);
    else k_i8_vslice_mma_closing<S, 8><<<gx, kI8VcWarps * 32, i8_vslice_mma_closing_smem<8>(S), stream>>>(a);
    return cudaGetLastError();
}
template <int S, bool CLOSING>
inline cudaError_t i8_launch_vslice(const I8VsArgs& a, cudaStream_t stream) {
    // iterate builds: split the rows in two when that is needed to give every SM at least ~3 CTAs
    const unsigned gx = (unsigned)((a.n_chains + kI8VsThreads - 1) / kI8VsThreads);
    unsigned gy = 1;
    if (!CLOSING) {
        const int blocks = a.n_rows_pad / kI8VsRows;
        while (gx * gy < 148 * 4 && (int)gy * 2 <= blocks && gy < 16) gy *= 2;
    }
    const dim3 grid(gx, gy);
    const size_t smem = i8_vslice_smem(a.xs);
    switch (i8_dp(a.dim)) {
        case 8: k_i8_vslice<S, 8, CLOSING><<<grid, kI8VsThreads, smem, stream>>>(a); break;
        case 16: k_i8_vslice<S, 16, CLOSING><<<grid, kI8VsThreads, smem, stream>>>(a); break;
        case 26: k_i8_vslice<S, 26, CLOSING><<<grid, kI8VsThreads, smem, stream>>>(a); break;
        default: k_i8_vslice<S, 32, CLOSING><<<grid, kI8VsThreads, smem, stream>>>(a);
    }
    return cudaGetLastError();
}
template <int S>
inline cudaError_t i8_launch_gemm(const CUtensorMap& map_a, const CUtensorMap& map_b, const I8GemmArgs& a, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_i8_gemm<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8Shape<S>::SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const unsigned splits = a.kb_split > 0 ? (unsigned)((a.k_blocks + a.kb_split - 1) / a.kb_split) : 1u;
    const dim3 grid((unsigned)i8_chunks<S>(a.p2), (unsigned)((a.n_chains + kI8TileM - 1) / kI8TileM), splits);
    k_i8_gemm<S><<<grid, I8Shape<S>::THREADS, I8Shape<S>::SMEM, stream>>>(map_a, map_b, a);
    return cudaGetLastError();
}
#endif

}  // namespace rmhmc
