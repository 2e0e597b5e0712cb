// C ABI of librmhmc_b200.so (include/rmhmc_b200.h): handle, device layouts, round scheduling.
// Host side only orchestrates; all arithmetic is in the kernels of this directory.
#include "../../include/rmhmc_b200.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "chain_kernels.cuh"
#include "chain_tpc.cuh"
#include "common.cuh"
#include "ess_kernel.cuh"
#include "hmc_kernels.cuh"
#include "hmc_fused.cuh"
#include "i8_metric.cuh"
#include "metric_kernel.cuh"
#include "mf_kernels.cuh"
#include "momfp_kernel.cuh"
#include "pass_kernel.cuh"
#include "mmala_kernels.cuh"
#include "peaks.cuh"
#include "tbuild_kernel.cuh"

using namespace rmhmc;

namespace {
thread_local std::string g_create_error;

struct ProfSlot {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    double ms = 0.0;
    int64_t launches = 0;
};
}  // namespace

// NCCL is resolved at run time (dlopen) so that the library loads on hosts without it; it is only
// needed by the row-sharded mode (rmhmc_comm_init).
namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi& nccl_api() {
    static NcclApi api;
    if (api.lib) return api;
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) return api;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.lib, "ncclAllReduce"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
    return api;
}
}  // namespace

struct rmhmc_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t n_rows = 0;
    int dim = 0, xs = 0, n_rows_pad = 0, p2 = 0, p2p = 0, p3 = 0, p3p = 0, nt = 0, extra_tile = -1;
    int main_tiles = 0, col_ctas = 1;      // metric kernel: tiles spread over G-warps, column CTAs per chain tile
    double alpha = 100.0;
    // data set
    double* x_pad = nullptr;
    double* kr3 = nullptr;          // [Np][P3p] KR3(X), only when it fits kKr3Budget (else formed on the fly)
    double* kr2t = nullptr;         // [P2k][Np] KR2(X)^T: B operand of the leverage GEMM (matrix-free partials)
    double* kr2n = nullptr;         // [Np][P2p] KR2(X): B operand of the plain-GEMM metric build (32 < D only)
    bool suppress_brackets = false; // a caller's Bracket spans several launches
    int p2k = 0;                    // P2 padded to the GEMM's K tile
    bool matrix_free = false;       // partials mode of the engine (rmhmc_set_partials_mode)
    uchar2* pair_tab = nullptr;
    uchar4* tri_tab = nullptr;
    unsigned short* tidx = nullptr;
    unsigned int* tidx32 = nullptr;
    unsigned char *pair_a = nullptr, *pair_b = nullptr;
    // chains
    int64_t n_chains = 0, c_pad = 0;
    bool is_hmc = false;
    bool is_mmala = false;
    int mmala_simplified = 0;
    std::vector<void*> chain_allocs;
    ChainArrays S{};
    EngineParams P{};
    long long* d_remaining = nullptr;
    bool configured = false, rng_set = false;
    int64_t launches = 0;
    bool profiling = false;
    bool fuse_epilogues = false;
    bool gemm_split_k = true;       // plain-GEMM metric build: split K against wave quantisation (RMHMC_GEMM_SPLITK=0 disables)
    bool metric_gemm = true;        // 32 < D: position-iterate metric builds as v-kernel + plain GEMM (RMHMC_METRIC_GEMM=0: fused kernel)
    bool hmc_fused = true;          // HMC: many leapfrog rounds per launch (hmc_fused.cuh) instead of three launches per round
    int hmc_rounds_per_launch = 64;
    int fuse_momentum = 1;          // implicit momentum half-step: 1 all iterates in one k_pass launch, 2 k_mom_fp, 0 unfused
    // row-sharded mode: this handle holds the rows of shard `shard_rank`; every build is all-reduced
    ncclComm_t comm = nullptr;
    int shard_world = 1, shard_rank = 0;
    double* t_tmp = nullptr;        // [C][P3p] contiguous partials of the last build (sharded mode)
    double* split_buf = nullptr;    // partial outputs of row-split metric builds / passes
    size_t split_cap = 0;
    ProfSlot prof[13];
    ncclComm_t stats_comm = nullptr;   // chain-sharded runs: end-of-run statistics only (rmhmc_stats_comm_init)
    int stats_world = 1;
    // INT8-slice metric build on tcgen05 (i8_metric.cuh): digit planes of KR2(X)^T per data set, of V per chain set
    int metric_mode = RMHMC_METRIC_FP64_DMMA;
    int i8_slices = 5;              // digits per operand (RMHMC_I8_SLICES = 5 | 6)
    bool i8_ok = false;             // B planes formed (tensor-map encoder available, planes fit a third of the free memory)
    signed char* b8 = nullptr;      // [S][b_rows][kp]
    double2* colinfo = nullptr;     // [b_rows]
    double* colmax = nullptr;       // [P2p]
    int i8_kp = 0, i8_b_rows = 0;
    CUtensorMap map_b;
    signed char* a8 = nullptr;      // [S][c_pad][kp] (with the chains)
    CUtensorMap map_a;
    // leverage GEMM h = q . KR2(X)^T on the same kernel: digits of KR2(X) by data row (per data set), of q (per chain set)
    bool i8_leverage = true;        // RMHMC_I8_LEVERAGE=0: keep the FP64 DMMA GEMM
    signed char* bl8 = nullptr;     // [S][bl_rows][kpl]
    double2* colinfo_l = nullptr;   // [bl_rows]
    double* rowmax_l = nullptr;     // [Np]
    int i8_kpl = 0, i8_bl_rows = 0;
    CUtensorMap map_bl;
    signed char* aq8 = nullptr;     // [S][c_pad][kpl] (with the chains)
    double* qscale = nullptr;       // [c_pad]
    CUtensorMap map_aq;
    // kernel-variant selection that depends on the chain count (few chains: SM-filling variants); tests pin it
    int launch_regime = RMHMC_REGIME_AUTO;
    int64_t chain_gen = 0;          // bumped whenever the handle's chain set is (re)allocated or freed (rmhmc_chain_generation)
    bool matrix_free_user = false;  // the caller's partials mode; mmala_chains_init overrides matrix_free for its own chain set only
    int64_t tape_base = 0, tape_window = 0;      // iterations covered by the host tape (rng_mode 0)
    mutable std::string err;
};

namespace {

#define CUDA_TRY(h, expr)                                                                     \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                   \
            return RMHMC_E_CUDA;                                                              \
        }                                                                                     \
    } while (0)

int fail(rmhmc_handle* h, int code, const std::string& msg) {
    h->err = msg;
    return code;
}

// true: take the kernel variants meant for batches that do not fill the GPU with the throughput tiles
// (rmhmc_set_launch_regime pins the choice)
bool few_chains(const rmhmc_handle* h, int64_t threshold) {
    if (h->launch_regime == RMHMC_REGIME_SMALL) return true;
    if (h->launch_regime == RMHMC_REGIME_LARGE) return false;
    return h->n_chains < threshold;
}
// 32 < D: v reaches the digit kernel through hbuf, which only the matrix-free mode allocates
bool use_i8(const rmhmc_handle* h) {
    return h->metric_mode == RMHMC_METRIC_INT8_TCGEN05 && h->i8_ok && (h->dim <= kMaxDimWarp || h->matrix_free);
}

// ------------------------------------------------------------------ small layout kernels
__global__ void k_pad_design(const double* __restrict__ xx, const double* __restrict__ t, double* __restrict__ xp,
                             int64_t n_rows, int dim, int xs, int64_t n_rows_pad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_rows_pad * xs) return;
    int64_t r = i / xs;
    int col = (int)(i - r * xs);
    double v = 0.0;
    if (r < n_rows) {
        if (col < dim) v = xx[r * dim + col];
        else if (col == xs - 1) v = t[r];
    }
    xp[i] = v;
}

// KR2(X)^T: row k = packed pair (a, b), column n = design-matrix row: x_na x_nb (zero rows for k >= P2)
__global__ void k_form_kr2t(const double* __restrict__ x, const uchar2* __restrict__ pair_tab, double* __restrict__ kr2t,
                            long long n_rows_pad, int xs, int p2, int p2k) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_rows_pad * p2k) return;
    int k = (int)(i / n_rows_pad);
    long long n = i - (long long)k * n_rows_pad;
    double v = 0.0;
    if (k < p2) {
        uchar2 ab = pair_tab[k];
        v = x[n * xs + ab.x] * x[n * xs + ab.y];
    }
    kr2t[i] = v;
}

__global__ void k_fill(double* p, int64_t n, double v) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// packed G -> dense symmetric
__global__ void k_unpack_g(const double* __restrict__ gp, double* __restrict__ G, int64_t C, int D, int p2p) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= C * D * D) return;
    int64_t c = i / (D * D);
    int r = (int)(i - c * D * D);
    int a = r / D, b = r - a * D;
    int lo = a < b ? a : b, hi = a < b ? b : a;
    G[i] = gp[c * p2p + pair_index(lo, hi, D)];
}
// dense symmetric -> packed
__global__ void k_pack_g(const double* __restrict__ G, double* __restrict__ gp, int64_t C, int D, int p2, int p2p,
                         const unsigned char* __restrict__ pa, const unsigned char* __restrict__ pb) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= C * p2p) return;
    int64_t c = i / p2p;
    int pr = (int)(i - c * p2p);
    gp[i] = pr < p2 ? G[c * D * D + pa[pr] * D + pb[pr]] : 0.0;
}
// packed T -> dense dG[c][d][a][b]
__global__ void k_unpack_t(const double* __restrict__ tp, double* __restrict__ dG, int64_t C, int D, int p3p) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t d3 = (int64_t)D * D * D;
    if (i >= C * d3) return;
    int64_t c = i / d3;
    int r = (int)(i - c * d3);
    int d = r / (D * D), a = (r / D) % D, b = r % D;
    dG[i] = tp[c * p3p + triple_index_any(d, a, b, D)];
}
// full gradient and log joint from the closing build's raw outputs
__global__ void k_seam_finish(const double* __restrict__ theta, const double* __restrict__ grad_raw,
                              const double* __restrict__ loglik, double* grad, double* logjoint, int64_t C, int D,
                              double alpha) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= C) return;
    double lp = 0.0;
    for (int d = 0; d < D; ++d) {
        double th = theta[c * D + d];
        if (grad) grad[c * D + d] = grad_raw[c * D + d] - th / alpha;
        lp += -0.5 * log(2.0 * 3.14159265358979323846 * alpha) - th * th / (2.0 * alpha);
    }
    if (logjoint) logjoint[c] = loglik[c] + lp;
}
// per chain: factor packed G; optional L, inverse, logdet, trace(G^-1 dG_d) from packed T
__host__ inline size_t seam_smem_bytes(int dim, int p3p, bool with_t) {
    return ((size_t)3 * dim * (dim | 1) + 32) * 8 + 8 + (with_t ? (size_t)p3p * 8 : 0);
}
template <int DMAX, int NCH>
__global__ void __launch_bounds__(32) k_seam_factor(EngineParams P, const double* __restrict__ gp,
                                                    const double* __restrict__ tp, double* L, double* Ginv,
                                                    double* logdet, double* trace) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = blockIdx.x, lane = threadIdx.x, D = P.dim, DS = P.ds;
    double* Lsm = reinterpret_cast<double*>(smem_raw);
    double* Msm = Lsm + D * DS;
    double* IG = Msm + D * DS;
    double* outv = IG + D * DS;
    double* Tsm = outv + 32;
    if ((Tsm - Lsm) & 1) ++Tsm;
    double dinv, ig[DMAX];
    {
        double lrow[DMAX];
        load_packed_rows<DMAX>(gp + (size_t)c * P.p2p, lrow, D, lane);
        double ld = chol_regs<DMAX>(lrow, D, lane, dinv);
        store_rows<DMAX>(Lsm, lrow, D, DS, lane);
        if (logdet && lane == 0) logdet[c] = ld;
    }
    if (L)
        for (int idx = lane; idx < D * D; idx += 32) L[(size_t)c * D * D + idx] = Lsm[(idx / D) * DS + (idx % D)];
    if (Ginv || trace) {
        chol_inverse_regs<DMAX>(Lsm, Msm, dinv, ig, D, DS, lane);
#pragma unroll
        for (int b = 0; b < DMAX; ++b)
            if (b < D && lane < D) IG[lane * DS + b] = ig[b];
        __syncwarp();
        if (Ginv)
            for (int idx = lane; idx < D * D; idx += 32) Ginv[(size_t)c * D * D + idx] = IG[(idx / D) * DS + (idx % D)];
    }
    if (trace) {
        PairRegs<NCH> pr;
        load_pairs<NCH>(P, pr, lane);
        const double2* ts = reinterpret_cast<const double2*>(tp + (size_t)c * P.p3p);
        for (int i = lane; i < P.p3p / 2; i += 32) reinterpret_cast<double2*>(Tsm)[i] = ts[i];
        __syncwarp();
        trace_terms<NCH>(P, pr, Tsm, IG, outv, 0, 1, lane);
        if (lane < D) trace[(size_t)c * D + lane] = outv[lane];
    }
}

// D > 32 variant: one CTA per chain
__global__ void __launch_bounds__(kBigThreads) k_seam_factor_big(EngineParams P, const double* __restrict__ gp,
                                                                 const double* __restrict__ tp, double* L, double* Ginv,
                                                                 double* logdet, double* trace) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = blockIdx.x, tid = threadIdx.x, D = P.dim, DS = P.ds;
    double* A = reinterpret_cast<double*>(smem_raw);
    double* B = A + D * DS;
    double* outv = B + D * DS;
    double* q = outv + kMaxDimBig;
    unpack_sym_cta(gp + (size_t)c * P.p2p, A, D, DS);
    double ld = chol_cta(A, D, DS);
    if (logdet && tid == 0) logdet[c] = ld;
    if (L)
        for (int idx = tid; idx < D * D; idx += kBigThreads) L[(size_t)c * D * D + idx] = A[(idx / D) * DS + (idx % D)];
    if (Ginv || trace) {
        chol_inverse_cta(A, B, D, DS);
        if (Ginv)
            for (int idx = tid; idx < D * D; idx += kBigThreads) Ginv[(size_t)c * D * D + idx] = B[(idx / D) * DS + (idx % D)];
    }
    if (trace) {
        for (int p = tid; p < P.p2; p += kBigThreads) {
            int pa = P.pair_a[p], pb = P.pair_b[p];
            q[p] = (pa == pb ? 1.0 : 2.0) * B[pa * DS + pb];
        }
        __syncthreads();
        tensor_contract_big(tp + (size_t)c * P.p3p, q, P.tidx32, outv, D, P.p2);
        if (tid < D) trace[(size_t)c * D + tid] = outv[tid];
    }
}

__global__ void k_remaining(const long long* __restrict__ iter, int64_t C, long long it_stop, long long* out) {
    long long m = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < C; i += (int64_t)gridDim.x * blockDim.x) {
        long long r = it_stop - iter[i];
        if (r > m) m = r;
    }
    if (m > 0) atomicMax(out, m);
}

__global__ void k_copy_i64(const long long* src, int64_t* dst, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
__global__ void k_copy_i32(const int* src, int32_t* dst, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
__global__ void k_copy_theta(const double* theta, const int* cur, size_t slot_stride, double* dst, int64_t C, int D) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= C * D) return;
    int64_t c = i / D;
    dst[i] = theta[(size_t)cur[c] * slot_stride + i];
}

// manifold-MALA / IWLS: mean and lower Cholesky factor of the current proposal distribution of every chain
__global__ void k_copy_proposal(const double* theta, const double* drift, const double* rfac, const int* cur, size_t slot_theta,
                                size_t slot_mat, double* mean, double* chol, int64_t C, int D) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= C * D * D) return;
    const int64_t c = i / (D * D);
    const int r = (int)(i - c * D * D), a = r / D, b = r - a * D;
    const int sl = cur[c];
    if (chol) chol[i] = rfac[(size_t)sl * slot_mat + (size_t)c * D * D + (size_t)b * D + a];      // L[a][b] = R[b][a]
    if (mean && b == 0) mean[c * D + a] = theta[(size_t)sl * slot_theta + c * D + a] + drift[(size_t)sl * slot_theta + c * D + a];
}

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// ------------------------------------------------------------------ profiling brackets
struct Bracket {
    rmhmc_handle* h;
    int kind;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool on;
    Bracket(rmhmc_handle* h_, int kind_) : h(h_), kind(kind_), on(h_->profiling && !h_->suppress_brackets) {
        if (on) {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, h->stream);
        }
    }
    ~Bracket();
};

void drain_profile(rmhmc_handle* h);
Bracket::~Bracket() {
    if (on) {
        cudaEventRecord(e1, h->stream);
        h->prof[kind].ev.emplace_back(e0, e1);
        // bound the number of live events while profiling stays enabled (draining waits for the recorded launches)
        if (h->prof[kind].ev.size() >= 16384) drain_profile(h);
    }
}

void drain_profile(rmhmc_handle* h) {
    for (auto& slot : h->prof) {
        for (auto& pr : slot.ev) {
            float ms = 0.f;
            cudaEventSynchronize(pr.second);
            cudaEventElapsedTime(&ms, pr.first, pr.second);
            slot.ms += ms;
            slot.launches += 1;
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        slot.ev.clear();
    }
}

// ------------------------------------------------------------------ kernel launch helpers
FuseArgs fuse_args(rmhmc_handle* h, int mode, int is_last, int init) {
    FuseArgs f{};
    f.mode = mode; f.is_last = is_last; f.init = init;
    f.step_size = h->P.step_size; f.it_stop = h->P.it_stop;
    const ChainArrays& S = h->S;
    f.mom = S.mom; f.theta = S.theta; f.u0 = S.u0; f.theta_w = S.theta_w; f.dir = S.dir; f.step = S.step;
    f.cur = S.cur; f.iter = S.iter; f.nsteps = S.nsteps; f.renorm_pos = S.renorm_pos;
    f.lfac = S.lfac; f.invg = S.invg; f.logdet = S.logdet;
    f.slot_theta = h->P.slot_theta; f.slot_invg = h->P.slot_invg; f.slot_scalar = h->P.slot_scalar;
    return f;
}

template <int MODE>
int launch_metric(rmhmc_handle* h, const MetricArgs& a_in, const FuseArgs& fz = FuseArgs{}) {
    MetricArgs a = a_in;
    size_t smem = metric_smem_bytes(h->xs, h->p2p, fz.mode != kFuseNone);
    dim3 grid(blocks_for(a.n_chains, kMetricChains), MODE >= 2 ? 1u : (unsigned)h->col_ctas);
    // Few chains and very many rows (BASELINE.json configs[4]: 64 chains, 1.25 M rows per GPU): the chain / column tiles
    // alone leave most SMs idle, so the rows are split over gridDim.z and the partial sums added in split order.
    const int n_blocks_all = h->n_rows_pad / kMetricRows;
    int splits = 1;
    if (fz.mode == kFuseNone && h->launch_regime != RMHMC_REGIME_LARGE && grid.x * grid.y < 74) splits = std::max(1, std::min((int)(296 / (grid.x * grid.y)), n_blocks_all / 16));
    const size_t cg = (size_t)a.n_chains * h->p2p, cd = (size_t)a.n_chains * h->dim, cl = (size_t)a.n_chains;
    if (splits > 1) {
        const size_t need = (size_t)splits * (cg + cd + cl);
        if (need > h->split_cap) {
            if (h->split_buf) CUDA_TRY(h, cudaFree(h->split_buf));
            h->split_buf = nullptr; h->split_cap = 0;
            CUDA_TRY(h, cudaMalloc((void**)&h->split_buf, need * 8));
            h->split_cap = need;
        }
        a.g_out = h->split_buf; a.split_g = cg;
        a.grad_out = h->split_buf + (size_t)splits * cg; a.split_grad = cd;
        a.loglik_out = a.grad_out + (size_t)splits * cd; a.split_ll = cl;
        grid.z = (unsigned)splits;
    }
    if (fz.mode != kFuseNone && grid.y != 1) return fail(h, RMHMC_E_UNSUPPORTED, "fused epilogues need a single column CTA");
    void (*kern)(MetricArgs, FuseArgs) = nullptr;
    int nt = MODE >= 2 ? 1 : h->nt;
    if constexpr (MODE >= 2) {
        kern = k_metric<1, MODE>;            // no G accumulators: one instantiation
    } else {
        switch (nt) {
            case 1: kern = k_metric<1, MODE>; break;
            case 2: kern = k_metric<2, MODE>; break;
            case 3: kern = k_metric<3, MODE>; break;
            case 4: kern = k_metric<4, MODE>; break;
            case 5: kern = k_metric<5, MODE>; break;
            case 6: kern = k_metric<6, MODE>; break;
            default: return fail(h, RMHMC_E_UNSUPPORTED, "metric kernel: dim too large");
        }
    }
    CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        Bracket b(h, MODE == 0 || MODE == 5 ? 0 : (MODE == 3 ? 5 : (MODE == 4 ? 7 : 1)));
        kern<<<grid, kMetricThreads, smem, h->stream>>>(a, fz);
        if (splits > 1) {
            if (MODE <= 1 && a_in.g_out) k_reduce_splits<<<blocks_for(cg, 256), 256, 0, h->stream>>>(a.g_out, cg, splits, a_in.g_out, cg);
            if (MODE >= 1 && a_in.grad_out) k_reduce_splits<<<blocks_for(cd, 256), 256, 0, h->stream>>>(a.grad_out, cd, splits, a_in.grad_out, cd);
            if ((MODE == 1 || MODE == 2 || MODE == 6) && a_in.loglik_out) k_reduce_splits<<<blocks_for(cl, 256), 256, 0, h->stream>>>(a.loglik_out, cl, splits, a_in.loglik_out, cl);
            h->launches += 3;
        }
    }
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}

MetricArgs metric_args(rmhmc_handle* h, int64_t C, const double* theta, double* g_out, double* grad_out,
                       double* loglik_out, double* cbuf) {
    MetricArgs a{};
    a.x = h->x_pad; a.pair_tab = h->pair_tab; a.theta = theta;
    a.g_out = g_out; a.grad_out = grad_out; a.loglik_out = loglik_out; a.cbuf = cbuf; a.extra_tile = h->extra_tile;
    a.tiles_per_cta = kMetricGWarps * h->nt; a.n_main_tiles = h->main_tiles;
    a.n_chains = (int)C; a.n_rows = (int)h->n_rows; a.n_rows_pad = h->n_rows_pad;
    a.dim = h->dim; a.xs = h->xs; a.p2 = h->p2; a.p2p = h->p2p;
    a.alpha_inv = h->shard_rank == 0 ? 1.0 / h->alpha : 0.0;      // the prior term I/alpha is added once across shards
    return a;
}

// closing build of the engine: in matrix-free mode c_n goes to the proposal (flip = 1) / current (flip = 0) slot
MetricArgs closing_args(rmhmc_handle* h, int flip) {
    ChainArrays& S = h->S;
    MetricArgs a = metric_args(h, h->n_chains, S.theta_w, S.g_tmp, S.grad_tmp, S.loglik_tmp, h->matrix_free ? S.cw : S.cbuf);
    if (h->matrix_free) { a.cw_cur = S.cur; a.cw_flip = flip; a.cw_slot = h->P.slot_cw; }
    return a;
}
// passes over the data of the matrix-free mode: MODE 3 quadratic forms (u = uvec), MODE 4 traces (h = hbuf)
MetricArgs pass_args(rmhmc_handle* h, double* out) {
    ChainArrays& S = h->S;
    MetricArgs a = metric_args(h, h->n_chains, S.uvec, nullptr, out, nullptr, S.cw);
    a.aslot = S.aslot; a.cw_slot = h->P.slot_cw; a.hbuf = S.hbuf;
    return a;
}

int launch_tbuild(rmhmc_handle* h, int64_t C, const double* cbuf, double* tpack, const int* cur, int flip,
                  size_t slot_stride) {
    TBuildArgs a{};
    a.x = h->x_pad; a.tri_tab = h->tri_tab; a.cbuf = cbuf; a.tpack = tpack; a.cur = cur; a.flip = flip;
    a.slot_stride = slot_stride; a.n_chains = (int)C; a.n_rows_pad = h->n_rows_pad; a.xs = h->xs;
    a.p3 = h->p3; a.p3p = h->p3p;
    a.kr3 = h->kr3;
    dim3 grid((unsigned)((h->p3p + kTbCols - 1) / kTbCols), blocks_for(C, kTbChains));
    if (h->kr3) {
        size_t smem = tbuild_pre_smem_bytes();
        CUDA_TRY(h, cudaFuncSetAttribute(k_tbuild_pre, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        Bracket b(h, 2);
        k_tbuild_pre<<<grid, kTbThreads, smem, h->stream>>>(a);
    } else {
        const int stages = tbuild_stages(h->xs);
        size_t smem = tbuild_smem_bytes(h->xs, stages);
        auto kern = stages == 3 ? k_tbuild<3> : k_tbuild<2>;
        CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        Bracket b(h, 2);
        kern<<<grid, kTbThreads, smem, h->stream>>>(a);
    }
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}

// leverages of the newest metric, hbuf[c][n] = x_n^T G_c^-1 x_n = sum_pairs q_c[pair] KR2(X)[n][pair]: the same
// TMA-fed DMMA GEMM as the partials build with (A, B, K, columns) = (qpack, KR2(X)^T, P2k, Np)
template <int S>
int i8_leverage_s(rmhmc_handle* h) {
    const int64_t C = h->n_chains;
    {
        Bracket b(h, 12);
        k_i8_qdigits<S><<<blocks_for(C, 8), 256, 0, h->stream>>>(h->S.qpack, h->p2, h->p2k, h->aq8, (size_t)h->c_pad * h->i8_kpl,
                                                                  h->i8_kpl, h->qscale, (int)C);
    }
    I8GemmArgs g{};
    g.g_out = h->S.hbuf; g.colinfo = h->colinfo_l; g.alpha_inv = 0.0; g.rowscale = h->qscale;
    g.n_chains = (int)C; g.p2 = h->n_rows_pad; g.p2p = h->n_rows_pad; g.k_blocks = h->i8_kpl / kI8BlockK;
    g.a_rows = (int)h->c_pad; g.b_rows = h->i8_bl_rows; g.debug_class = -1;
    cudaError_t e;
    {
        Bracket b(h, 6);
        e = i8_launch_gemm<S>(h->map_aq, h->map_bl, g, h->stream);
    }
    if (e != cudaSuccess) return fail(h, RMHMC_E_CUDA, std::string("k_i8_gemm (leverage): ") + cudaGetErrorString(e));
    h->launches += 2;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}
int launch_leverage(rmhmc_handle* h) {
    if (h->aq8 && use_i8(h)) return h->i8_slices == 6 ? i8_leverage_s<6>(h) : i8_leverage_s<5>(h);
    TBuildArgs a{};
    a.kr3 = h->kr2t; a.cbuf = h->S.qpack; a.tpack = h->S.hbuf; a.cur = h->S.cur; a.flip = 0; a.slot_stride = 0;
    a.n_chains = (int)h->n_chains; a.n_rows_pad = h->p2k; a.p3 = h->n_rows_pad; a.p3p = h->n_rows_pad;
    dim3 grid((unsigned)((h->n_rows_pad + kTbCols - 1) / kTbCols), blocks_for(h->n_chains, kTbChains));
    size_t smem = tbuild_pre_smem_bytes();
    CUDA_TRY(h, cudaFuncSetAttribute(k_tbuild_pre, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        Bracket b(h, 6);
        k_tbuild_pre<<<grid, kTbThreads, smem, h->stream>>>(a);
    }
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}

template <typename T>
int dev_alloc(rmhmc_handle* h, T** p, size_t count, std::vector<void*>* track) {
    void* q = nullptr;
    CUDA_TRY(h, cudaMalloc(&q, count * sizeof(T) + 16));
    CUDA_TRY(h, cudaMemsetAsync(q, 0, count * sizeof(T) + 16, h->stream));
    *p = reinterpret_cast<T*>(q);
    if (track) track->push_back(q);
    return RMHMC_OK;
}

void free_chains(rmhmc_handle* h) {
    h->chain_gen += 1;
    for (void* p : h->chain_allocs) cudaFree(p);
    h->chain_allocs.clear();
    h->S = ChainArrays{};
    h->t_tmp = nullptr;
    h->a8 = nullptr; h->aq8 = nullptr; h->qscale = nullptr;
    h->n_chains = 0;
}

void fill_engine_params(rmhmc_handle* h) {
    EngineParams& P = h->P;
    P.n_chains = (int)h->n_chains; P.dim = h->dim; P.ds = h->dim | 1;
    P.p2 = h->p2; P.p2p = h->p2p; P.p3 = h->p3; P.p3p = h->p3p; P.n_rows_pad = h->n_rows_pad;
    P.alpha = h->alpha; P.tidx = h->tidx; P.tidx32 = h->tidx32; P.pair_a = h->pair_a; P.pair_b = h->pair_b;
    size_t C = (size_t)h->n_chains;
    P.slot_theta = C * h->dim; P.slot_scalar = C;
    P.slot_invg = C * h->dim * h->dim; P.slot_t = C * h->p3p;
    P.matrix_free = h->matrix_free ? 1 : 0; P.p2k = h->p2k; P.slot_cw = (size_t)h->c_pad * h->n_rows_pad;
}

int alloc_chains(rmhmc_handle* h, int64_t C, bool hmc) {
    free_chains(h);
    h->n_chains = C;
    h->c_pad = pad_up((int)C, kTbChains);
    h->is_hmc = hmc;
    h->is_mmala = false;
    auto* tr = &h->chain_allocs;
    size_t c = (size_t)C, D = (size_t)h->dim;
    ChainArrays& S = h->S;
    int rc = 0;
    rc |= dev_alloc(h, &S.theta, 2 * c * D, tr);
    rc |= dev_alloc(h, &S.logjoint, 2 * c, tr);
    rc |= dev_alloc(h, &S.grad, 2 * c * D, tr);
    if (!hmc) {
        rc |= dev_alloc(h, &S.lfac, 2 * c * D * D, tr);
        rc |= dev_alloc(h, &S.invg, 2 * c * D * D, tr);
        rc |= dev_alloc(h, &S.logdet, 2 * c, tr);
        rc |= dev_alloc(h, &S.trace, 2 * c * D, tr);
        rc |= dev_alloc(h, &S.u0, c * D, tr);
        if (h->matrix_free) {
            rc |= dev_alloc(h, &S.cw, 2 * (size_t)h->c_pad * h->n_rows_pad, tr);
            rc |= dev_alloc(h, &S.hbuf, (size_t)h->c_pad * h->n_rows_pad, tr);
            rc |= dev_alloc(h, &S.qpack, (size_t)h->c_pad * h->p2k, tr);
            rc |= dev_alloc(h, &S.uvec, c * D, tr);
            rc |= dev_alloc(h, &S.quad_tmp, 2 * c * D, tr);          // quad_tmp | trace_tmp: one exchange when row-sharded
            S.trace_tmp = S.quad_tmp ? S.quad_tmp + c * D : nullptr;
            rc |= dev_alloc(h, &S.aslot, c, tr);
        } else {
            rc |= dev_alloc(h, &S.tpack, 2 * c * h->p3p, tr);
            rc |= dev_alloc(h, &S.cbuf, (size_t)h->c_pad * h->n_rows_pad, tr);
            if (h->comm) rc |= dev_alloc(h, &h->t_tmp, c * h->p3p, tr);
        }
    }
    // build outputs are contiguous (g_tmp | grad_tmp | loglik_tmp) so that the row-sharded mode reduces them in one call
    {
        double* build = nullptr;
        size_t g_len = hmc ? 0 : c * h->p2p;
        rc |= dev_alloc(h, &build, g_len + c * D + c, tr);
        S.g_tmp = hmc ? nullptr : build;
        S.grad_tmp = build + g_len;
        S.loglik_tmp = S.grad_tmp + c * D;
    }
    h->a8 = nullptr;
    if (!hmc && use_i8(h)) {
        rc |= dev_alloc(h, &h->a8, (size_t)h->i8_slices * h->c_pad * h->i8_kp, tr);
        if (!rc && !make_tensor_map_u8_k64(&h->map_a, h->a8, (uint64_t)h->i8_slices * h->c_pad, (uint64_t)h->i8_kp, kI8TileM)) {
            free_chains(h);
            return fail(h, RMHMC_E_CUDA, "cuTensorMapEncodeTiled failed for the V digit planes");
        }
    }
    h->aq8 = nullptr; h->qscale = nullptr;
    if (!hmc && use_i8(h) && h->bl8 && h->matrix_free) {
        rc |= dev_alloc(h, &h->aq8, (size_t)h->i8_slices * h->c_pad * h->i8_kpl, tr);
        rc |= dev_alloc(h, &h->qscale, (size_t)h->c_pad, tr);
        if (!rc && !make_tensor_map_u8_k64(&h->map_aq, h->aq8, (uint64_t)h->i8_slices * h->c_pad, (uint64_t)h->i8_kpl, kI8TileM)) {
            free_chains(h);
            return fail(h, RMHMC_E_CUDA, "cuTensorMapEncodeTiled failed for the q digit planes");
        }
    }
    rc |= dev_alloc(h, &S.mom, c * D, tr);
    rc |= dev_alloc(h, &S.theta_w, c * D, tr);
    rc |= dev_alloc(h, &S.hcur, c, tr);
    rc |= dev_alloc(h, &S.pudot, c, tr);
    rc |= dev_alloc(h, &S.cur, c, tr);
    rc |= dev_alloc(h, &S.step, c, tr);
    rc |= dev_alloc(h, &S.nsteps, c, tr);
    rc |= dev_alloc(h, &S.dir, c, tr);
    rc |= dev_alloc(h, &S.iter, c, tr);
    rc |= dev_alloc(h, &S.accepted, c, tr);
    rc |= dev_alloc(h, &S.leapfrogs, c, tr);
    rc |= dev_alloc(h, &S.renorm_mom, c, tr);
    rc |= dev_alloc(h, &S.renorm_pos, c, tr);
    if (rc) { free_chains(h); return RMHMC_E_CUDA; }
    fill_engine_params(h);
    return RMHMC_OK;
}

// per-chain kernel variants: DMAX = 16 / 32 (unroll bound of the register linear algebra) and
// NCH = ceil(P2 / 32) rounded up to 4 / 11 / 17 (packed pairs per lane in the tensor contractions)
int dmax_variant(const rmhmc_handle* h) { return h->dim <= 16 ? 0 : 1; }
int nch_variant(const rmhmc_handle* h) { return h->p2 <= 128 ? 0 : (h->p2 <= 352 ? 1 : 2); }

bool is_big(const rmhmc_handle* h) { return h->dim > kMaxDimWarp; }

int set_chain_smem_attrs(rmhmc_handle* h) {
    if (is_big(h)) {
        int turn = (int)turn_smem_bytes(h->dim, h->p2, h->p3p, true);
        int seam = (int)(big_mat_smem_bytes(h->dim, 2) + (size_t)h->p2 * 8);
        CUDA_TRY(h, cudaFuncSetAttribute(k_chain_turn<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, turn));
        CUDA_TRY(h, cudaFuncSetAttribute(k_seam_factor_big, cudaFuncAttributeMaxDynamicSharedMemorySize, seam));
        CUDA_TRY(h, cudaFuncSetAttribute(k_chain_factor_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_mat_smem_bytes(h->dim, 2)));
        CUDA_TRY(h, cudaFuncSetAttribute(k_chain_solve_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_mat_smem_bytes(h->dim, 1)));
        return RMHMC_OK;
    }
    int turn = (int)turn_smem_bytes(h->dim, h->p2, h->p3p, false), seam = (int)seam_smem_bytes(h->dim, h->p3p, true);
    CUDA_TRY(h, cudaFuncSetAttribute(k_chain_turn<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, turn));
    CUDA_TRY(h, cudaFuncSetAttribute(k_chain_turn<11, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, turn));
    CUDA_TRY(h, cudaFuncSetAttribute(k_chain_turn<17, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, turn));
    CUDA_TRY(h, cudaFuncSetAttribute(k_seam_factor<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, seam));
    CUDA_TRY(h, cudaFuncSetAttribute(k_seam_factor<32, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, seam));
    CUDA_TRY(h, cudaFuncSetAttribute(k_seam_factor<32, 17>, cudaFuncAttributeMaxDynamicSharedMemorySize, seam));
    CUDA_TRY(h, cudaFuncSetAttribute(k_chain_solve_tpc<kTpcTile>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tpc_smem_bytes(h->dim)));
    CUDA_TRY(h, cudaFuncSetAttribute(k_chain_factor_tpc<kTpcTile>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tpc_smem_bytes(h->dim)));
    return RMHMC_OK;
}

int launch_turn(rmhmc_handle* h, int do_back, int do_front, int init) {
    const unsigned C = (unsigned)h->n_chains;
    const bool big = is_big(h);
    size_t smem = turn_smem_bytes(h->dim, h->p2, h->p3p, big);
    {
        Bracket b(h, 3);
        if (big) k_chain_turn<1, true><<<C, kBigThreads, smem, h->stream>>>(h->P, h->S, do_back, do_front, init);
        else switch (nch_variant(h)) {
            case 0: k_chain_turn<4, false><<<C, kTurnThreads, smem, h->stream>>>(h->P, h->S, do_back, do_front, init); break;
            case 1: k_chain_turn<11, false><<<C, kTurnThreads, smem, h->stream>>>(h->P, h->S, do_back, do_front, init); break;
            default: k_chain_turn<17, false><<<C, kTurnThreads, smem, h->stream>>>(h->P, h->S, do_back, do_front, init);
        }
    }
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}

// matrix-free mode: per-chain halves of a round and one iterate of the implicit momentum half-step
int launch_mf_turn(rmhmc_handle* h, int do_back, int do_front, int init) {
    const unsigned C = (unsigned)h->n_chains;
    {
        Bracket b(h, 3);
        if (is_big(h)) k_mf_turn<128><<<C, 128, 0, h->stream>>>(h->P, h->S, do_back, do_front, init);
        else k_mf_turn<32><<<C, 32, 0, h->stream>>>(h->P, h->S, do_back, do_front, init);
    }
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}
int launch_mf_mom_iter(rmhmc_handle* h, int is_last) {
    const unsigned C = (unsigned)h->n_chains;
    {
        Bracket b(h, 3);
        if (is_big(h)) k_mf_mom_iter<128><<<C, 128, 0, h->stream>>>(h->P, h->S, is_last);
        else k_mf_mom_iter<32><<<C, 32, 0, h->stream>>>(h->P, h->S, is_last);
    }
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}

// thread-per-chain Cholesky stages (chain_tpc.cuh): an alternative formulation, measured and NOT the default.  3x fewer
// warp instructions than the warp-per-chain kernels, but the 32 packed metrics of a warp fill 105 KB of shared memory, so
// only two warps are resident per SM and the kernel is bound by the latency of a single warp per scheduler: solve
// 0.275 ms (warp per chain: 0.272 ms), factor 1.25 ms (0.55 ms) at 65 536 German-shaped chains
// (tests/native/tpc_bench.cu, profiles/r02/tpc_bench.log).  RMHMC_CHAIN_TPC=1 / RMHMC_FACTOR_TPC=1 switch them on.
bool use_tpc(const rmhmc_handle* h) {
    static const int forced = [] { const char* e = getenv("RMHMC_CHAIN_TPC"); return e ? atoi(e) : 0; }();
    return forced != 0 && !is_big(h);
}
bool use_tpc_factor(const rmhmc_handle* h) {
    static const int forced = [] { const char* e = getenv("RMHMC_FACTOR_TPC"); return e ? atoi(e) : 0; }();
    return forced != 0 && !is_big(h);
}

int launch_factor(rmhmc_handle* h, int init) {
    const unsigned C = (unsigned)h->n_chains;
    {
        Bracket b(h, 4);
        if (is_big(h)) k_chain_factor_big<<<C, kBigThreads, big_mat_smem_bytes(h->dim, 2), h->stream>>>(h->P, h->S, init);
        else if (use_tpc_factor(h)) k_chain_factor_tpc<kTpcTile><<<(C + 31) / 32, 32, tpc_smem_bytes(h->dim), h->stream>>>(h->P, h->S, init);
        else switch (chain_order(h->dim)) {
            case 8: k_chain_factor<8><<<C, 32, factor_smem_bytes(8), h->stream>>>(h->P, h->S, init); break;
            case 16: k_chain_factor<16><<<C, 32, factor_smem_bytes(16), h->stream>>>(h->P, h->S, init); break;
            case 25: k_chain_factor<25><<<C, 32, factor_smem_bytes(25), h->stream>>>(h->P, h->S, init); break;
            default: k_chain_factor<32><<<C, 32, factor_smem_bytes(32), h->stream>>>(h->P, h->S, init);
        }
    }
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}

int launch_solve(rmhmc_handle* h, int is_last) {
    const unsigned C = (unsigned)h->n_chains;
    {
        Bracket b(h, 4);
        if (is_big(h)) k_chain_solve_big<<<C, kBigThreads, big_mat_smem_bytes(h->dim, 1), h->stream>>>(h->P, h->S, is_last);
        else if (use_tpc(h)) k_chain_solve_tpc<kTpcTile><<<(C + 31) / 32, 32, tpc_smem_bytes(h->dim), h->stream>>>(h->P, h->S, is_last);
        else switch (chain_order(h->dim)) {
            case 8: k_chain_solve<8><<<C, 32, solve_smem_bytes(8), h->stream>>>(h->P, h->S, is_last); break;
            case 16: k_chain_solve<16><<<C, 32, solve_smem_bytes(16), h->stream>>>(h->P, h->S, is_last); break;
            case 25: k_chain_solve<25><<<C, 32, solve_smem_bytes(25), h->stream>>>(h->P, h->S, is_last); break;
            default: k_chain_solve<32><<<C, 32, solve_smem_bytes(32), h->stream>>>(h->P, h->S, is_last);
        }
    }
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}

int launch_seam_factor(rmhmc_handle* h, const EngineParams& P, int64_t C, const double* gp, const double* tp, double* L,
                       double* Ginv, double* logdet, double* trace) {
    if (is_big(h)) {
        size_t smem = big_mat_smem_bytes(h->dim, 2) + (size_t)h->p2 * 8;
        k_seam_factor_big<<<(unsigned)C, kBigThreads, smem, h->stream>>>(P, gp, tp, L, Ginv, logdet, trace);
    } else {
        size_t smem = seam_smem_bytes(h->dim, h->p3p, tp != nullptr);
        if (h->dim <= 16 && h->p2 <= 128)
            k_seam_factor<16, 4><<<(unsigned)C, 32, smem, h->stream>>>(P, gp, tp, L, Ginv, logdet, trace);
        else if (h->p2 <= 352)
            k_seam_factor<32, 11><<<(unsigned)C, 32, smem, h->stream>>>(P, gp, tp, L, Ginv, logdet, trace);
        else
            k_seam_factor<32, 17><<<(unsigned)C, 32, smem, h->stream>>>(P, gp, tp, L, Ginv, logdet, trace);
    }
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}

int allreduce_sum(rmhmc_handle* h, double* buf, size_t count) {
    if (!h->comm) return RMHMC_OK;
    Bracket b(h, 10);
    ncclResult_t r = nccl_api().AllReduce(buf, buf, count, ncclDouble, ncclSum, h->comm, h->stream);
    if (r != ncclSuccess) return fail(h, RMHMC_E_CUDA, std::string("ncclAllReduce: ") + nccl_api().GetErrorString(r));
    return RMHMC_OK;
}
// outputs of a metric build of the given mode, summed over the row shards
template <int MODE>
int reduce_build(rmhmc_handle* h) {
    if (!h->comm) return RMHMC_OK;
    const size_t c = (size_t)h->n_chains, D = (size_t)h->dim;
    if (MODE == 0) return allreduce_sum(h, h->S.g_tmp, c * h->p2p);
    if (MODE == 1) return allreduce_sum(h, h->S.g_tmp, c * h->p2p + c * D + c);
    return allreduce_sum(h, h->S.grad_tmp, c * D + c);
}
__global__ void k_scatter_t(const double* __restrict__ src, double* __restrict__ tpack, const int* __restrict__ cur, int flip,
                            size_t slot_stride, int64_t C, int p3p) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= C * (int64_t)(p3p / 2)) return;
    int64_t c = i / (p3p / 2);
    int k = (int)(i - c * (p3p / 2));
    reinterpret_cast<double2*>(tpack + (size_t)(cur[c] ^ flip) * slot_stride + (size_t)c * p3p)[k] =
        reinterpret_cast<const double2*>(src + (size_t)c * p3p)[k];
}
// partials build into the proposal (flip = 1) or current (flip = 0) slot; row-sharded: build into a
// contiguous buffer, all-reduce, scatter into the slots
int build_partials(rmhmc_handle* h, int flip) {
    const int64_t C = h->n_chains;
    ChainArrays& S = h->S;
    if (!h->comm) return launch_tbuild(h, C, S.cbuf, S.tpack, S.cur, flip, h->P.slot_t);
    int rc = launch_tbuild(h, C, S.cbuf, h->t_tmp, S.cur, 0, 0);       // slot stride 0: row c of t_tmp
    if (rc) return rc;
    rc = allreduce_sum(h, h->t_tmp, (size_t)C * h->p3p);
    if (rc) return rc;
    k_scatter_t<<<blocks_for(C * (h->p3p / 2), 256), 256, 0, h->stream>>>(h->t_tmp, S.tpack, S.cur, flip, h->P.slot_t, C, h->p3p);
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}

// ---- matrix-free partials (mf_kernels.cuh)
// F iterates of the implicit momentum half-step (rmhmc.py:102-110): quadratic forms by a pass over the data,
// then the per-chain update and u = G^-1 PM for the next pass
constexpr int64_t kPassSmallBelow = (int64_t)148 * 6 * kPassWarpsSmall * 8;      // 14 208 chains
constexpr int64_t kMomFpBelow = 1024;
template <int KIND>
int launch_pass(rmhmc_handle* h) {
    // 64 chains per CTA, or 16 (kPassWarpsSmall) while the 16-chain CTAs are all resident at once (6 per SM): one
    // warp owns 8 chains for the whole kernel either way -- the two tilings are bit-identical -- and with few chains the
    // finer tiles balance the SMs (quad pass, German-shaped: 12 288 chains 0.45 vs 0.57 ms, 16 384 chains 0.76 vs 0.58 ms;
    // profiles/r02/regimes_small_large.txt)
    const bool small = few_chains(h, kPassSmallBelow);
    const int warps = small ? kPassWarpsSmall : kPassWarps;
    const size_t smem = pass_smem_bytes(h->xs, warps, KIND);
    // D = 8 k + 1: the last parameter through FMAs instead of a DMMA k-step and d-tile of its own (pass_kernel.cuh)
    const bool tail = (h->dim & 7) == 1 && h->dim > 8;
    void (*kern)(EngineParams, ChainArrays, const double*, int) =
        small ? (tail ? k_pass<KIND, kPassWarpsSmall, true> : k_pass<KIND, kPassWarpsSmall, false>)
              : (tail ? k_pass<KIND, kPassWarps, true> : k_pass<KIND, kPassWarps, false>);
    CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        Bracket b(h, KIND == kPassTrace || KIND == kPassPair ? 7 : 5);
        kern<<<blocks_for(h->n_chains, warps * 8), warps * 32, smem, h->stream>>>(h->P, h->S, h->x_pad, h->xs);
    }
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}

int mf_momentum_fixed_point(rmhmc_handle* h) {
    ChainArrays& S = h->S;
    const size_t cd = (size_t)h->n_chains * h->dim;
    // all F iterates in one launch (needs no exchange between iterates).  Two formulations: one warp per 8 chains
    // (pass_kernel.cuh; 2.58 vs 3.02 ms at 65 536 German-shaped chains) or 12 cooperating warps per 32 chains
    // (momfp_kernel.cuh; shorter dependent chains, 0.08 vs 0.23 ms when 4096 chains leave most of the GPU idle)
    const bool fusable = !is_big(h) && !h->comm && h->P.n_fixed >= 2 && !h->P.student_t;      // Student-t: per-iterate kernels
    // k_mom_fp only where latency counts (a handful of chains: the single-chain drop-in, the seam tests) or when the
    // SMALL regime is pinned; 8192 chains: k_pass<MOMFP, 2 warps> 0.38 ms, k_mom_fp 0.44 ms
    const bool few = h->launch_regime == RMHMC_REGIME_AUTO ? h->n_chains < kMomFpBelow : few_chains(h, 0);
    if (fusable && (h->fuse_momentum == 1 && !few)) return launch_pass<kPassMomFp>(h);
    if (fusable && (h->fuse_momentum == 2 || (h->fuse_momentum == 1 && few))) {
        const size_t smem = momfp_smem_bytes(h->xs);
        CUDA_TRY(h, cudaFuncSetAttribute(k_mom_fp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            Bracket b(h, 5);
            k_mom_fp<<<blocks_for(h->n_chains, kMetricChains), kMomThreads, smem, h->stream>>>(h->P, S, h->x_pad, h->xs);
        }
        h->launches += 1;
        CUDA_TRY(h, cudaGetLastError());
        return RMHMC_OK;
    }
    for (int fi = 0; fi < h->P.n_fixed; ++fi) {
        int rc = is_big(h) ? launch_metric<3>(h, pass_args(h, S.quad_tmp)) : launch_pass<kPassQuad>(h);
        if (!rc) rc = allreduce_sum(h, S.quad_tmp, cd);
        if (!rc) rc = launch_mf_mom_iter(h, fi + 1 == h->P.n_fixed ? 1 : 0);
        if (rc) return rc;
    }
    return RMHMC_OK;
}
// after the factorisation of the new metric: leverages, tr(G^-1 dG_d) and -- unless init -- the quadratic forms of
// the explicit momentum half-step (rmhmc.py:142-161)
int mf_closing_passes(rmhmc_handle* h, int init) {
    ChainArrays& S = h->S;
    const size_t cd = (size_t)h->n_chains * h->dim;
    int rc = launch_leverage(h);
    if (!rc && !is_big(h)) rc = init ? launch_pass<kPassTrace>(h) : launch_pass<kPassPair>(h);
    if (!rc && is_big(h)) rc = launch_metric<4>(h, pass_args(h, S.trace_tmp));
    if (!rc && is_big(h) && !init) rc = launch_metric<3>(h, pass_args(h, S.quad_tmp));
    if (!rc) rc = init ? allreduce_sum(h, S.trace_tmp, cd) : allreduce_sum(h, S.quad_tmp, 2 * cd);
    return rc;
}

// diagonal pairs of the packed metric += 1/alpha (the plain-GEMM build has no epilogue for it)
__global__ void k_add_prior_diag(double* __restrict__ gp, int64_t C, int D, int p2p, double alpha_inv) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= C * D) return;
    int64_t c = i / D;
    int a = (int)(i - c * D);
    gp[c * p2p + pair_index(a, a, D)] += alpha_inv;
}
// Metric of a position fixed-point iterate, G(theta_w) -> g_tmp.  32 < D with enough chain x column tiles to fill the GPU:
// v -> hbuf (k_metric<5>), then the TMA-fed DMMA GEMM G = V . KR2(X) (k_tbuild_pre) -- the column CTAs of the fused
// kernel would each recompute f and v (14x at D = 100).  Otherwise the fused kernel.
bool use_metric_gemm(const rmhmc_handle* h) {
    return h->metric_gemm && is_big(h) && h->kr2n && h->matrix_free &&
           blocks_for(h->n_chains, kTbChains) * ((h->p2p + kTbCols - 1) / kTbCols) >= 148;
}
int ensure_split_buf(rmhmc_handle* h, size_t need) {
    if (need > h->split_cap) {
        if (h->split_buf) CUDA_TRY(h, cudaFree(h->split_buf));
        h->split_buf = nullptr; h->split_cap = 0;
        CUDA_TRY(h, cudaMalloc((void**)&h->split_buf, need * 8));
        h->split_cap = need;
    }
    return RMHMC_OK;
}
// g_tmp = hbuf . KR2(X) + I/alpha (no brackets of its own)
int launch_metric_gemm(rmhmc_handle* h) {
    ChainArrays& S = h->S;
    const int64_t C = h->n_chains;
    TBuildArgs t{};
    t.kr3 = h->kr2n; t.cbuf = S.hbuf; t.tpack = S.g_tmp; t.cur = S.cur; t.flip = 0; t.slot_stride = 0;
    t.n_chains = (int)C; t.n_rows_pad = h->n_rows_pad; t.p3 = h->p2; t.p3p = h->p2p;
    dim3 grid((unsigned)((h->p2p + kTbCols - 1) / kTbCols), blocks_for(C, kTbChains));
    // wave quantisation: a few hundred tiles on 148 SMs (cfg3: 320 = 2.16 waves) waste up to a third of the last wave;
    // split K so that the tail is at most ~5 % (partial products added in split order)
    const int tiles = (int)(grid.x * grid.y);
    int splits = 1;
    if (h->gemm_split_k) {
        double best = 1e30;
        for (int s = 1; s <= 4 && (h->n_rows_pad / kTbRows) / s >= 64; ++s) {
            const double waves = (double)tiles * s / 148.0, cost = std::ceil(waves) / s;      // time in units of a full tile
            if (cost < best * 0.97) { best = cost; splits = s; }
        }
    }
    const size_t cg = (size_t)C * h->p2p;
    if (splits > 1) {
        int rc = ensure_split_buf(h, (size_t)splits * cg);
        if (rc) return rc;
        t.tpack = h->split_buf; t.split_stride = cg;
        grid.z = (unsigned)splits;
    }
    size_t smem = tbuild_pre_smem_bytes();
    CUDA_TRY(h, cudaFuncSetAttribute(k_tbuild_pre, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tbuild_pre<<<grid, kTbThreads, smem, h->stream>>>(t);
    if (splits > 1) {
        k_reduce_splits<<<blocks_for(cg, 256), 256, 0, h->stream>>>(h->split_buf, cg, splits, S.g_tmp, cg);
        h->launches += 1;
    }
    k_add_prior_diag<<<blocks_for(C * h->dim, 256), 256, 0, h->stream>>>(S.g_tmp, C, h->dim, h->p2p,
                                                                          h->shard_rank == 0 ? 1.0 / h->alpha : 0.0);
    h->launches += 2;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}
// ---- INT8-slice metric build on tcgen05 (i8_metric.cuh)
template <int S>
int i8_form_b_planes(rmhmc_handle* h, cudaStream_t st) {
    k_i8_colmax<<<(unsigned)h->p2, 256, 0, st>>>(h->x_pad, h->pair_tab, h->colmax, h->n_rows_pad, h->xs);
    const long long n = (long long)h->i8_b_rows * h->i8_kp;
    k_i8_form_b<S><<<blocks_for(n, 256), 256, 0, st>>>(h->x_pad, h->pair_tab, h->colmax, h->b8, h->colinfo, h->n_rows_pad,
                                                       h->xs, h->p2, h->i8_b_rows, h->i8_kp);
    h->launches += 2;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}
template <int S>
int i8_form_bl_planes(rmhmc_handle* h, cudaStream_t st) {
    k_i8_rowmax<<<blocks_for(h->n_rows_pad, 8), 256, 0, st>>>(h->x_pad, h->pair_tab, h->rowmax_l, h->n_rows_pad, h->xs, h->p2);
    const long long n = (long long)h->i8_bl_rows * h->i8_kpl;
    k_i8_form_bl<S><<<blocks_for(n, 256), 256, 0, st>>>(h->x_pad, h->pair_tab, h->rowmax_l, h->bl8, h->colinfo_l, h->n_rows_pad,
                                                        h->xs, h->p2, h->i8_bl_rows, h->i8_kpl);
    h->launches += 2;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}
int i8_form_b(rmhmc_handle* h, cudaStream_t st) {
    int rc = h->i8_slices == 6 ? i8_form_b_planes<6>(h, st) : i8_form_b_planes<5>(h, st);
    if (!rc && h->bl8) rc = h->i8_slices == 6 ? i8_form_bl_planes<6>(h, st) : i8_form_bl_planes<5>(h, st);
    return rc;
}

// digit planes of KR2(X)^T + their tensor map; leaves i8_ok false when the shape is outside the kernel's range
int i8_setup(rmhmc_handle* h) {
    if (!tensor_map_encoder()) return RMHMC_OK;
    const int S = h->i8_slices, nc = S == 6 ? I8Shape<6>::NC : I8Shape<5>::NC;
    h->i8_kp = i8_kp(h->n_rows_pad);
    h->i8_b_rows = (h->p2 + nc - 1) / nc * nc;
    if (const char* e = std::getenv("RMHMC_I8_LEVERAGE")) h->i8_leverage = std::atoi(e) != 0;
    {
        // digit planes are S bytes per entry of KR2(X) (and again, by data row, for the leverage GEMM): stay within a
        // third of the free memory, drop the leverage planes first
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const size_t b_bytes = (size_t)S * h->i8_b_rows * h->i8_kp;
        const size_t bl_bytes = (size_t)S * ((h->n_rows_pad + nc - 1) / nc * nc) * pad_up(h->p2, kI8BlockK);
        if (b_bytes > free_b / 3) return RMHMC_OK;
        if (b_bytes + bl_bytes > free_b / 3) h->i8_leverage = false;
    }
    CUDA_TRY(h, cudaMalloc((void**)&h->b8, (size_t)S * h->i8_b_rows * h->i8_kp));
    CUDA_TRY(h, cudaMalloc((void**)&h->colinfo, (size_t)h->i8_b_rows * sizeof(double2)));
    CUDA_TRY(h, cudaMalloc((void**)&h->colmax, (size_t)h->p2p * 8));
    if (h->i8_leverage && h->kr2t) {
        h->i8_kpl = pad_up(h->p2, kI8BlockK);
        h->i8_bl_rows = (h->n_rows_pad + nc - 1) / nc * nc;
        CUDA_TRY(h, cudaMalloc((void**)&h->bl8, (size_t)S * h->i8_bl_rows * h->i8_kpl));
        CUDA_TRY(h, cudaMalloc((void**)&h->colinfo_l, (size_t)h->i8_bl_rows * sizeof(double2)));
        CUDA_TRY(h, cudaMalloc((void**)&h->rowmax_l, (size_t)h->n_rows_pad * 8));
    }
    int rc = i8_form_b(h, h->stream);
    if (rc) return rc;
    if (!make_tensor_map_u8_k64(&h->map_b, h->b8, (uint64_t)S * h->i8_b_rows, (uint64_t)h->i8_kp, (uint32_t)nc))
        return fail(h, RMHMC_E_CUDA, "cuTensorMapEncodeTiled failed for the KR2(X) digit planes");
    if (h->bl8 && !make_tensor_map_u8_k64(&h->map_bl, h->bl8, (uint64_t)S * h->i8_bl_rows, (uint64_t)h->i8_kpl, (uint32_t)nc))
        return fail(h, RMHMC_E_CUDA, "cuTensorMapEncodeTiled failed for the leverage digit planes");
    h->i8_ok = true;
    return RMHMC_OK;
}

// G = (digits of V in a8) . KR2(X) digits + I/alpha -> g_out; K is split when the int32 accumulators could overflow
// (more than kI8MaxRows rows) or when the tiles alone would leave most SMs idle; partial sums are added in split order
template <int S>
int i8_metric_gemm(rmhmc_handle* h, int64_t C, int64_t a_rows, const CUtensorMap& map_a, double* g_out) {
    I8GemmArgs g{};
    g.g_out = g_out; g.colinfo = h->colinfo; g.alpha_inv = h->shard_rank == 0 ? 1.0 / h->alpha : 0.0;
    g.n_chains = (int)C; g.p2 = h->p2; g.p2p = h->p2p; g.k_blocks = h->i8_kp / kI8BlockK;
    g.a_rows = (int)a_rows; g.b_rows = h->i8_b_rows; g.debug_class = -1;
    const int max_kb = kI8MaxRows / kI8BlockK;
    int splits = (g.k_blocks + max_kb - 1) / max_kb;
    const int64_t tiles = (int64_t)i8_chunks<S>(h->p2) * ((C + kI8TileM - 1) / kI8TileM);
    while (tiles * splits < 148 && g.k_blocks / (splits * 2) >= 64) splits *= 2;
    const size_t cg = (size_t)C * h->p2p;
    if (splits > 1) {
        int rc = ensure_split_buf(h, (size_t)splits * cg);
        if (rc) return rc;
        g.kb_split = (g.k_blocks + splits - 1) / splits;
        splits = (g.k_blocks + g.kb_split - 1) / g.kb_split;
        g.g_out = h->split_buf; g.split_stride = cg;
    }
    cudaError_t e = i8_launch_gemm<S>(map_a, h->map_b, g, h->stream);
    if (e != cudaSuccess) return fail(h, RMHMC_E_CUDA, std::string("k_i8_gemm: ") + cudaGetErrorString(e));
    h->launches += 1;
    if (splits > 1) {
        k_reduce_splits<<<blocks_for(cg, 256), 256, 0, h->stream>>>(h->split_buf, cg, splits, g_out, cg);
        h->launches += 1;
        CUDA_TRY(h, cudaGetLastError());
    }
    return RMHMC_OK;
}
// 32 < D: v is already in HBM (k_metric<MODE 5 / 6> -> vbuf[C][Np]); digits, then the GEMM
template <int S>
int i8_gemm_from_v_s(rmhmc_handle* h, int64_t C, const double* vbuf, signed char* a8, int64_t a_rows, const CUtensorMap& map_a,
                     double* g_out) {
    {
        Bracket bv(h, 8);
        const long long n = (long long)C * (h->i8_kp / 16);
        k_i8_vdigits<S><<<blocks_for(n, 256), 256, 0, h->stream>>>(vbuf, h->n_rows_pad, a8, (size_t)a_rows * h->i8_kp, h->i8_kp,
                                                                    (long long)C);
        h->launches += 1;
    }
    CUDA_TRY(h, cudaGetLastError());
    Bracket bg(h, 9);
    return i8_metric_gemm<S>(h, C, a_rows, map_a, g_out);
}
int i8_gemm_from_v(rmhmc_handle* h, int64_t C, const double* vbuf, signed char* a8, int64_t a_rows, const CUtensorMap& map_a,
                   double* g_out) {
    return h->i8_slices == 6 ? i8_gemm_from_v_s<6>(h, C, vbuf, a8, a_rows, map_a, g_out)
                             : i8_gemm_from_v_s<5>(h, C, vbuf, a8, a_rows, map_a, g_out);
}

struct I8Closing {              // outputs of the closing build besides G (null: position-iterate build)
    double* grad_out; double* loglik_out; double* cbuf; const int* cw_cur; int cw_flip; size_t cw_slot;
};
template <int S>
int i8_build_s(rmhmc_handle* h, int64_t C, const double* theta, signed char* a8, int64_t a_rows, const CUtensorMap& map_a,
               double* g_out, const I8Closing* cl) {
    I8VsArgs v{};
    v.x = h->x_pad; v.theta = theta; v.a8 = a8; v.plane_stride = (size_t)a_rows * h->i8_kp; v.kp = h->i8_kp;
    v.n_chains = (int)C; v.n_rows = (int)h->n_rows; v.n_rows_pad = h->n_rows_pad; v.dim = h->dim; v.xs = h->xs;
    cudaError_t e;
    {
        Bracket bv(h, cl ? 11 : 8);
        if (cl) {
            v.grad_out = cl->grad_out; v.loglik_out = cl->loglik_out; v.cbuf = cl->cbuf;
            v.cw_cur = cl->cw_cur; v.cw_flip = cl->cw_flip; v.cw_slot = cl->cw_slot;
            e = i8_launch_vslice_mma_closing<S>(v, h->stream);
        } else {
            e = i8_launch_vslice_mma<S>(v, h->stream);
        }
    }
    if (e != cudaSuccess) return fail(h, RMHMC_E_CUDA, std::string("k_i8_vslice: ") + cudaGetErrorString(e));
    h->launches += 1;
    Bracket bg(h, 9);
    return i8_metric_gemm<S>(h, C, a_rows, map_a, g_out);
}
int i8_build(rmhmc_handle* h, int64_t C, const double* theta, signed char* a8, int64_t a_rows, const CUtensorMap& map_a,
             double* g_out, const I8Closing* cl) {
    return h->i8_slices == 6 ? i8_build_s<6>(h, C, theta, a8, a_rows, map_a, g_out, cl)
                             : i8_build_s<5>(h, C, theta, a8, a_rows, map_a, g_out, cl);
}

int build_metric_iterate(rmhmc_handle* h) {
    ChainArrays& S = h->S;
    const int64_t C = h->n_chains;
    if (use_i8(h) && is_big(h)) {                // 32 < D: f and v by the FP64 kernel (v -> hbuf), then digits + GEMM
        Bracket b(h, 0);
        h->suppress_brackets = true;
        MetricArgs a = metric_args(h, C, S.theta_w, nullptr, nullptr, nullptr, nullptr);
        a.vout = S.hbuf;
        int rc = launch_metric<5>(h, a);
        h->suppress_brackets = false;
        return rc ? rc : i8_gemm_from_v(h, C, S.hbuf, h->a8, h->c_pad, h->map_a, S.g_tmp);
    }
    if (use_i8(h)) {
        Bracket b(h, 0);
        return i8_build(h, C, S.theta_w, h->a8, h->c_pad, h->map_a, S.g_tmp, nullptr);
    }
    if (!use_metric_gemm(h)) return launch_metric<0>(h, metric_args(h, C, S.theta_w, S.g_tmp, nullptr, nullptr, nullptr));
    Bracket b(h, 0);
    h->suppress_brackets = true;
    MetricArgs a = metric_args(h, C, S.theta_w, nullptr, nullptr, nullptr, nullptr);
    a.vout = S.hbuf;
    int rc = launch_metric<5>(h, a);
    if (!rc) rc = launch_metric_gemm(h);
    h->suppress_brackets = false;
    return rc;
}
// closing build of a leapfrog step (G, X^T (t - p), log-likelihood, c_n), same split for 32 < D
int build_metric_closing(rmhmc_handle* h, int flip) {
    if (use_i8(h) && is_big(h)) {
        Bracket b(h, 1);
        h->suppress_brackets = true;
        MetricArgs a = closing_args(h, flip);
        a.g_out = nullptr;
        a.vout = h->S.hbuf;
        int rc = launch_metric<6>(h, a);
        h->suppress_brackets = false;
        return rc ? rc : i8_gemm_from_v(h, h->n_chains, h->S.hbuf, h->a8, h->c_pad, h->map_a, h->S.g_tmp);
    }
    if (use_i8(h)) {
        ChainArrays& S = h->S;
        I8Closing cl{S.grad_tmp, S.loglik_tmp, h->matrix_free ? S.cw : S.cbuf, h->matrix_free ? S.cur : nullptr, flip,
                     h->matrix_free ? h->P.slot_cw : 0};
        Bracket b(h, 1);
        return i8_build(h, h->n_chains, S.theta_w, h->a8, h->c_pad, h->map_a, S.g_tmp, &cl);
    }
    if (!use_metric_gemm(h)) return launch_metric<1>(h, closing_args(h, flip));
    Bracket b(h, 1);
    h->suppress_brackets = true;
    MetricArgs a = closing_args(h, flip);
    a.g_out = nullptr;
    a.vout = h->S.hbuf;
    int rc = launch_metric<6>(h, a);
    if (!rc) rc = launch_metric_gemm(h);
    h->suppress_brackets = false;
    return rc;
}

// The builds of one RMHMC round (rmhmc.py:113-156 for every chain): F-1 position iterates, each a
// metric build + per-chain solve, then the closing metric build, the partials build and the
// per-chain factorisation of the new metric.  The
// per-chain halves around them (k_chain_turn) are launched by the callers.
int rmhmc_round_builds(rmhmc_handle* h) {
    const int64_t C = h->n_chains;
    ChainArrays& S = h->S;
    // h->fuse_epilogues: do the per-chain solve / factorisation in the metric kernels' epilogues.  Measured
    // SLOWER on B200 (metric 1.78 -> 2.95 ms per launch vs 0.62 ms for the stand-alone solve kernel at
    // 65536 chains: 12 warps per SM of straight-line code are instruction-fetch bound while the tensor
    // pipe idles), so it is off by default and kept for small chain counts / experiments.
    const bool fuse = h->fuse_epilogues && !h->comm && h->col_ctas == 1 && !h->matrix_free && !use_i8(h);
    for (int fi = 2; fi <= h->P.n_fixed; ++fi) {
        const int last = fi == h->P.n_fixed ? 1 : 0;
        MetricArgs a = metric_args(h, C, S.theta_w, S.g_tmp, nullptr, nullptr, nullptr);
        int rc = fuse ? launch_metric<0>(h, a, fuse_args(h, kFuseSolve, last, 0)) : build_metric_iterate(h);
        if (rc) return rc;
        if (!fuse) {
            rc = reduce_build<0>(h);            // row-sharded: one exchange per fixed-point iterate
            if (rc) return rc;
            rc = launch_solve(h, last);
            if (rc) return rc;
        }
    }
    int rc = fuse ? launch_metric<1>(h, closing_args(h, 1), fuse_args(h, kFuseFactor, 0, 0)) : build_metric_closing(h, 1);
    if (rc) return rc;
    rc = reduce_build<1>(h);
    if (rc) return rc;
    if (h->matrix_free) {
        rc = launch_factor(h, 0);
        if (!rc) rc = mf_closing_passes(h, 0);
        return rc;
    }
    rc = build_partials(h, 1);
    if (rc || fuse) return rc;
    return launch_factor(h, 0);
}

// n_rounds rounds: front | builds | back+front | builds | ... | back
int rmhmc_rounds(rmhmc_handle* h, int64_t n_rounds) {
    if (n_rounds <= 0) return RMHMC_OK;
    const bool mf = h->matrix_free;
    int rc = mf ? launch_mf_turn(h, 0, 1, 0) : launch_turn(h, 0, 1, 0);
    for (int64_t r = 0; r < n_rounds && !rc; ++r) {
        if (mf) rc = mf_momentum_fixed_point(h);
        if (!rc) rc = rmhmc_round_builds(h);
        if (!rc) rc = mf ? launch_mf_turn(h, 1, r + 1 < n_rounds ? 1 : 0, 0) : launch_turn(h, 1, r + 1 < n_rounds ? 1 : 0, 0);
    }
    return rc;
}

int hmc_round(rmhmc_handle* h) {
    const int64_t C = h->n_chains;
    ChainArrays& S = h->S;
    {
        Bracket b(h, 3);
        k_hmc_front<<<(unsigned)C, 32, 0, h->stream>>>(h->P, S);
    }
    MetricArgs a = metric_args(h, C, S.theta_w, nullptr, S.grad_tmp, S.loglik_tmp, nullptr);
    int rc = launch_metric<2>(h, a);
    if (rc) return rc;
    rc = reduce_build<2>(h);
    if (rc) return rc;
    {
        Bracket b(h, 3);
        k_hmc_back<<<(unsigned)C, 32, 0, h->stream>>>(h->P, S, 0);
    }
    h->launches += 2;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}

// n_rounds leapfrog rounds in ONE launch (hmc_fused.cuh): chain state in registers, X streamed n_rounds times
bool hmc_fusable(const rmhmc_handle* h) { return h->hmc_fused && h->dim <= 32 && !h->comm; }
int hmc_rounds(rmhmc_handle* h, int64_t n_rounds) {
    if (!hmc_fusable(h)) {
        for (int64_t r = 0; r < n_rounds; ++r) {
            int rc = hmc_round(h);
            if (rc) return rc;
        }
        return RMHMC_OK;
    }
    const bool small = few_chains(h, kPassSmallBelow);
    const int warps = small ? kPassWarpsSmall : kHfWarps;
    const size_t smem = hmc_fused_smem_bytes(h->xs, warps);
    const bool tail = (h->dim & 7) == 1 && h->dim > 8;
    void (*kern)(EngineParams, ChainArrays, const double*, int, int, int) =
        small ? (tail ? k_hmc_rounds<kPassWarpsSmall, true> : k_hmc_rounds<kPassWarpsSmall, false>)
              : (tail ? k_hmc_rounds<kHfWarps, true> : k_hmc_rounds<kHfWarps, false>);
    CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // bounded launches: a launch of 64 rounds is ~30 ms at 65 536 German-shaped chains
    for (int64_t done = 0; done < n_rounds;) {
        const int n = (int)std::min<int64_t>(n_rounds - done, h->hmc_rounds_per_launch);
        {
            Bracket b(h, 3);
            kern<<<blocks_for(h->n_chains, warps * 8), warps * 32, smem, h->stream>>>(h->P, h->S, h->x_pad, h->xs, (int)h->n_rows, n);
        }
        h->launches += 1;
        done += n;
    }
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}

// ---- manifold MALA (mmala_kernels.cuh): one round = one MCMC iteration of every chain
int launch_mmala_turn(rmhmc_handle* h, int do_back, int do_front, int init) {
    const unsigned C = (unsigned)h->n_chains;
    const int simp = h->mmala_simplified;
    {
        Bracket b(h, 3);
        switch (chain_order(h->dim)) {
            case 8: k_mmala_turn<8><<<C, 32, 0, h->stream>>>(h->P, h->S, do_back, do_front, init, simp); break;
            case 16: k_mmala_turn<16><<<C, 32, 0, h->stream>>>(h->P, h->S, do_back, do_front, init, simp); break;
            case 25: k_mmala_turn<25><<<C, 32, 0, h->stream>>>(h->P, h->S, do_back, do_front, init, simp); break;
            default: k_mmala_turn<32><<<C, 32, 0, h->stream>>>(h->P, h->S, do_back, do_front, init, simp);
        }
    }
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}
// metric quantities at theta_w into the proposal (flip = 1) / current (flip = 0) slot's inputs
int mmala_builds(rmhmc_handle* h, int flip) {
    int rc = build_metric_closing(h, flip);
    if (!rc) rc = reduce_build<1>(h);
    if (!rc) rc = launch_factor(h, flip ? 0 : 1);
    if (!rc && !h->mmala_simplified) {
        rc = launch_leverage(h);
        if (!rc) rc = launch_pass<kPassTrace>(h);
        if (!rc) rc = allreduce_sum(h, h->S.trace_tmp, (size_t)h->n_chains * h->dim);
    }
    return rc;
}
int mmala_rounds(rmhmc_handle* h, int64_t n_rounds) {
    if (n_rounds <= 0) return RMHMC_OK;
    int rc = launch_mmala_turn(h, 0, 1, 0);
    for (int64_t r = 0; r < n_rounds && !rc; ++r) {
        rc = mmala_builds(h, 1);
        if (!rc) rc = launch_mmala_turn(h, 1, r + 1 < n_rounds ? 1 : 0, 0);
    }
    return rc;
}

int run_until(rmhmc_handle* h, int64_t it_stop, int64_t* rounds_done, bool hmc, bool mmala = false) {
    if (h->n_chains <= 0 || h->is_hmc != hmc || h->is_mmala != mmala)
        return fail(h, RMHMC_E_STATE, "chains not initialised for this sampler");
    if (!h->configured || !h->rng_set) return fail(h, RMHMC_E_STATE, "configure and set a tape / philox seed first");
    if (h->P.rng_mode == 0 && h->P.student_t && !hmc && !mmala && !h->P.tape_z_chi)
        return fail(h, RMHMC_E_STATE, "Student-t momentum under a host tape needs rmhmc_set_tape_chi");
    if (h->P.rng_mode == 0) {
        // a host tape covers iterations [tape_base, tape_base + tape_window): never index outside it
        if (it_stop > h->tape_base + h->tape_window)
            return fail(h, RMHMC_E_INVALID, "it_stop lies beyond the iterations covered by the tape (set_tape window)");
        long long behind = 0;
        CUDA_TRY(h, cudaMemsetAsync(h->d_remaining, 0, sizeof(long long), h->stream));
        k_remaining<<<64, 256, 0, h->stream>>>(h->S.iter, h->n_chains, h->tape_base, h->d_remaining);
        CUDA_TRY(h, cudaMemcpyAsync(&behind, h->d_remaining, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        if (behind > 0) return fail(h, RMHMC_E_STATE, "a chain has completed fewer iterations than the tape's first iteration");
    }
    h->P.it_stop = it_stop;
    int64_t total = 0;
    for (;;) {
        long long rem = 0;
        CUDA_TRY(h, cudaMemsetAsync(h->d_remaining, 0, sizeof(long long), h->stream));
        k_remaining<<<64, 256, 0, h->stream>>>(h->S.iter, h->n_chains, it_stop, h->d_remaining);
        CUDA_TRY(h, cudaMemcpyAsync(&rem, h->d_remaining, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        if (rem <= 0) break;
        // every unfinished iteration needs at least one round; keep the host a bounded distance ahead
        int64_t chunk = rem < 512 ? rem : 512;
        if (hmc) {
            int rc = hmc_rounds(h, chunk);
            if (rc) return rc;
        } else if (mmala) {
            int rc = mmala_rounds(h, chunk);
            if (rc) return rc;
        } else {
            int rc = rmhmc_rounds(h, chunk);
            if (rc) return rc;
        }
        total += chunk;
    }
    if (rounds_done) *rounds_done = total;
    return RMHMC_OK;
}

}  // namespace

// ====================================================================== extern "C"
extern "C" {

const char* rmhmc_version(void) { return "rmhmc_b200 0.3 (sm_100a: fp64 dmma + int8-slice tcgen05 metric build)"; }

const char* rmhmc_last_error(const rmhmc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int rmhmc_create(rmhmc_handle** out, int device, int64_t n_rows, int dim, double alpha, const double* xx_dev,
                 const double* t_dev) {
    if (!out) return RMHMC_E_INVALID;
    *out = nullptr;
    if (n_rows <= 0 || dim <= 0 || !xx_dev || !t_dev || !(alpha > 0)) {
        g_create_error = "rmhmc_create: bad arguments";
        return RMHMC_E_INVALID;
    }
    if (dim > kMaxDimBig) {
        g_create_error = "rmhmc_create: dim > 128 is not supported by this version";
        return RMHMC_E_UNSUPPORTED;
    }
    auto* h = new rmhmc_handle();
    auto bail = [&](int code) {
        g_create_error = h->err;
        rmhmc_destroy(h);
        return code;
    };
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        h->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return bail(RMHMC_E_CUDA);
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        h->err = "rmhmc_create: this library is built for sm_100a (B200) only";
        return bail(RMHMC_E_UNSUPPORTED);
    }
    h->device = device;
    h->n_rows = n_rows; h->dim = dim; h->alpha = alpha;
    h->xs = x_stride(dim);
    h->n_rows_pad = pad_up((int)n_rows, 32);
    h->p2 = num_pairs(dim); h->p2p = pad_up(h->p2, 8);
    h->p3 = num_triples(dim); h->p3p = pad_up(h->p3, 8);
    h->p2k = pad_up(h->p2, kTbRows);
    {
        // packed-column tiles per G-warp; a single left-over tile is split over chain tiles instead
        int tiles = h->p2p / 8;
        const int max_nt = 6;                     // accumulator tiles per G-warp that fit the register file
        if (tiles % 4 == 1 && tiles > 1 && tiles <= kMetricGWarps * max_nt + 1) { h->extra_tile = tiles - 1; tiles -= 1; }
        else h->extra_tile = -1;
        h->main_tiles = tiles;
        h->nt = std::min(max_nt, (tiles + kMetricGWarps - 1) / kMetricGWarps);
        int d_tiles = (dim + 7) / 8;
        h->col_ctas = std::max((tiles + kMetricGWarps * h->nt - 1) / (kMetricGWarps * h->nt), (d_tiles + 3) / 4);
    }

    // index tables
    std::vector<uchar2> pair_tab(h->p2p, make_uchar2(0, 0));
    std::vector<unsigned char> pa(h->p2), pb(h->p2);
    for (int a = 0; a < dim; ++a)
        for (int b = a; b < dim; ++b) {
            int i = pair_index(a, b, dim);
            pair_tab[i] = make_uchar2((unsigned char)a, (unsigned char)b);
            pa[i] = (unsigned char)a; pb[i] = (unsigned char)b;
        }
    std::vector<uchar4> tri_tab(h->p3p, make_uchar4(0, 0, 0, 0));
    for (int i = 0; i < dim; ++i)
        for (int j = i; j < dim; ++j)
            for (int k = j; k < dim; ++k)
                tri_tab[triple_index(i, j, k, dim)] = make_uchar4((unsigned char)i, (unsigned char)j, (unsigned char)k, 0);
    const bool big = dim > kMaxDimWarp;
    std::vector<unsigned short> tidx(big ? 1 : (size_t)dim * h->p2, 0);
    std::vector<unsigned int> tidx32(big ? (size_t)dim * h->p2 : 1, 0);
    for (int d = 0; d < dim; ++d)
        for (int pr = 0; pr < h->p2; ++pr) {
            int ti = triple_index_any(pa[pr], pb[pr], d, dim);
            if (big) tidx32[(size_t)d * h->p2 + pr] = (unsigned int)ti;
            else tidx[(size_t)d * h->p2 + pr] = (unsigned short)ti;
        }

#define CREATE_TRY(expr)                                                          \
    do {                                                                          \
        cudaError_t e__ = (expr);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            h->err = std::string(#expr) + ": " + cudaGetErrorString(e__);         \
            return bail(RMHMC_E_CUDA);                                            \
        }                                                                         \
    } while (0)
    CREATE_TRY(cudaMalloc((void**)&h->x_pad, (size_t)h->n_rows_pad * h->xs * 8));
    CREATE_TRY(cudaMalloc((void**)&h->pair_tab, pair_tab.size() * sizeof(uchar2)));
    CREATE_TRY(cudaMalloc((void**)&h->tri_tab, tri_tab.size() * sizeof(uchar4)));
    CREATE_TRY(cudaMalloc((void**)&h->tidx, tidx.size() * sizeof(unsigned short)));
    CREATE_TRY(cudaMalloc((void**)&h->tidx32, tidx32.size() * sizeof(unsigned int)));
    CREATE_TRY(cudaMalloc((void**)&h->pair_a, pa.size()));
    CREATE_TRY(cudaMalloc((void**)&h->pair_b, pb.size()));
    CREATE_TRY(cudaMalloc((void**)&h->d_remaining, sizeof(long long)));
    CREATE_TRY(cudaMemcpy(h->pair_tab, pair_tab.data(), pair_tab.size() * sizeof(uchar2), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(h->tri_tab, tri_tab.data(), tri_tab.size() * sizeof(uchar4), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(h->tidx, tidx.data(), tidx.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(h->tidx32, tidx32.data(), tidx32.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(h->pair_a, pa.data(), pa.size(), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(h->pair_b, pb.data(), pb.size(), cudaMemcpyHostToDevice));
    int64_t total = (int64_t)h->n_rows_pad * h->xs;
    k_pad_design<<<blocks_for(total, 256), 256>>>(xx_dev, t_dev, h->x_pad, n_rows, dim, h->xs, h->n_rows_pad);
    CREATE_TRY(cudaGetLastError());
    {
        const size_t kKr3Budget = (size_t)1 << 30;      // 1 GiB: German-shaped needs 23 MB, Australian-shaped 4 MB
        size_t bytes = (size_t)h->n_rows_pad * h->p3p * 8;
        if (bytes <= kKr3Budget) {
            CREATE_TRY(cudaMalloc((void**)&h->kr3, bytes));
            long long n = (long long)h->n_rows_pad * h->p3p;
            k_form_kr3<<<blocks_for(n, 256), 256>>>(h->x_pad, h->tri_tab, h->kr3, h->n_rows_pad, h->xs, h->p3p);
            CREATE_TRY(cudaGetLastError());
        }
    }
    {
        // matrix-free partials need KR2(X)^T resident (2.7 MB German-shaped, 4 GB for N = 1e5, D = 100); without it
        // the engine falls back to the materialised tensor build
        const size_t kKr2Budget = (size_t)48 << 30;
        size_t bytes = (size_t)h->p2k * h->n_rows_pad * 8;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (bytes <= kKr2Budget && bytes < free_b / 2) {
            CREATE_TRY(cudaMalloc((void**)&h->kr2t, bytes));
            long long n = (long long)h->n_rows_pad * h->p2k;
            k_form_kr2t<<<blocks_for(n, 256), 256>>>(h->x_pad, h->pair_tab, h->kr2t, h->n_rows_pad, h->xs, h->p2, h->p2k);
            CREATE_TRY(cudaGetLastError());
            h->matrix_free = h->matrix_free_user = true;
        }
        // 32 < D: KR2(X) itself (rows x pairs) for the plain-GEMM metric build
        size_t bytes_n = (size_t)h->n_rows_pad * h->p2p * 8;
        cudaMemGetInfo(&free_b, &total_b);
        if (dim > kMaxDimWarp && h->kr2t && bytes_n <= kKr2Budget && bytes_n < free_b / 3) {
            CREATE_TRY(cudaMalloc((void**)&h->kr2n, bytes_n));
            long long n = (long long)h->n_rows_pad * h->p2p;
            k_form_kr2n<<<blocks_for(n, 256), 256>>>(h->x_pad, h->pair_tab, h->kr2n, h->n_rows_pad, h->xs, h->p2, h->p2p);
            CREATE_TRY(cudaGetLastError());
        }
    }
    if (const char* e = std::getenv("RMHMC_I8_SLICES")) h->i8_slices = std::atoi(e) == 6 ? 6 : 5;
    if (i8_setup(h) != RMHMC_OK) return bail(RMHMC_E_CUDA);
    if (h->i8_ok) h->metric_mode = RMHMC_METRIC_INT8_TCGEN05;       // default where the shape allows it (RMHMC_METRIC_MODE=dmma overrides)
    CREATE_TRY(cudaDeviceSynchronize());
#undef CREATE_TRY
    h->P.n_leapfrog = 6; h->P.step_size = 0.5; h->P.n_fixed = 4;
    if (const char* e = std::getenv("RMHMC_FUSE_MOMENTUM")) h->fuse_momentum = std::atoi(e);
    if (const char* e = std::getenv("RMHMC_HMC_FUSED")) h->hmc_fused = std::atoi(e) != 0;
    if (const char* e = std::getenv("RMHMC_METRIC_GEMM")) h->metric_gemm = std::atoi(e) != 0;
    if (const char* e = std::getenv("RMHMC_GEMM_SPLITK")) h->gemm_split_k = std::atoi(e) != 0;      // A/B switch for profiling
    if (const char* e = std::getenv("RMHMC_METRIC_MODE")) {
        if (std::string(e) == "i8" && h->i8_ok) h->metric_mode = RMHMC_METRIC_INT8_TCGEN05;
        if (std::string(e) == "dmma") h->metric_mode = RMHMC_METRIC_FP64_DMMA;
    }
    if (const char* e = std::getenv("RMHMC_LAUNCH_REGIME")) {
        const std::string r(e);
        h->launch_regime = r == "small" ? RMHMC_REGIME_SMALL : (r == "large" ? RMHMC_REGIME_LARGE : RMHMC_REGIME_AUTO);
    }
    h->P.it_stop = 0; h->P.burn_in = 0; h->P.sample_cap = 0;
    *out = h;
    return RMHMC_OK;
}

void rmhmc_destroy(rmhmc_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    drain_profile(h);
    free_chains(h);
    if (h->comm) nccl_api().CommDestroy(h->comm);
    if (h->stats_comm) nccl_api().CommDestroy(h->stats_comm);
    cudaFree(h->b8); cudaFree(h->colinfo); cudaFree(h->colmax); cudaFree(h->bl8); cudaFree(h->colinfo_l); cudaFree(h->rowmax_l); cudaFree(h->split_buf); cudaFree(h->kr2n); cudaFree(h->x_pad); cudaFree(h->kr3); cudaFree(h->kr2t); cudaFree(h->pair_tab); cudaFree(h->tri_tab); cudaFree(h->tidx); cudaFree(h->tidx32);
    cudaFree(h->pair_a); cudaFree(h->pair_b); cudaFree(h->d_remaining);
    delete h;
}

int rmhmc_update_data(rmhmc_handle* h, const double* xx_dev, const double* t_dev) {
    if (!h || !xx_dev || !t_dev) return h ? fail(h, RMHMC_E_INVALID, "rmhmc_update_data: bad arguments") : RMHMC_E_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int64_t total = (int64_t)h->n_rows_pad * h->xs;
    k_pad_design<<<blocks_for(total, 256), 256, 0, h->stream>>>(xx_dev, t_dev, h->x_pad, h->n_rows, h->dim, h->xs, h->n_rows_pad);
    h->launches += 1;
    if (h->kr3) {
        long long n = (long long)h->n_rows_pad * h->p3p;
        k_form_kr3<<<blocks_for(n, 256), 256, 0, h->stream>>>(h->x_pad, h->tri_tab, h->kr3, h->n_rows_pad, h->xs, h->p3p);
        h->launches += 1;
    }
    if (h->kr2t) {
        long long n = (long long)h->n_rows_pad * h->p2k;
        k_form_kr2t<<<blocks_for(n, 256), 256, 0, h->stream>>>(h->x_pad, h->pair_tab, h->kr2t, h->n_rows_pad, h->xs, h->p2, h->p2k);
        h->launches += 1;
    }
    if (h->kr2n) {
        long long n = (long long)h->n_rows_pad * h->p2p;
        k_form_kr2n<<<blocks_for(n, 256), 256, 0, h->stream>>>(h->x_pad, h->pair_tab, h->kr2n, h->n_rows_pad, h->xs, h->p2, h->p2p);
        h->launches += 1;
    }
    if (h->i8_ok) {
        int rc = i8_form_b(h, h->stream);
        if (rc) return rc;
    }
    CUDA_TRY(h, cudaGetLastError());
    return RMHMC_OK;
}

int rmhmc_set_partials_mode(rmhmc_handle* h, int mode) {
    if (!h || (mode != RMHMC_PARTIALS_TENSOR && mode != RMHMC_PARTIALS_MATRIX_FREE))
        return h ? fail(h, RMHMC_E_INVALID, "rmhmc_set_partials_mode: bad arguments") : RMHMC_E_INVALID;
    if (mode == RMHMC_PARTIALS_MATRIX_FREE && !h->kr2t)
        return fail(h, RMHMC_E_UNSUPPORTED, "rmhmc_set_partials_mode: KR2(X)^T does not fit on this device");
    if (h->n_chains > 0) free_chains(h);
    if (mode != RMHMC_PARTIALS_MATRIX_FREE) h->P.student_t = 0;       // the Student-t variant exists in the matrix-free kernels only
    h->matrix_free = h->matrix_free_user = mode == RMHMC_PARTIALS_MATRIX_FREE;
    return RMHMC_OK;
}
int rmhmc_get_partials_mode(const rmhmc_handle* h) {
    if (!h) return RMHMC_E_INVALID;
    return h->matrix_free ? (int)RMHMC_PARTIALS_MATRIX_FREE : (int)RMHMC_PARTIALS_TENSOR;
}

int rmhmc_set_metric_mode(rmhmc_handle* h, int mode) {
    if (!h || (mode != RMHMC_METRIC_FP64_DMMA && mode != RMHMC_METRIC_INT8_TCGEN05))
        return h ? fail(h, RMHMC_E_INVALID, "rmhmc_set_metric_mode: bad arguments") : RMHMC_E_INVALID;
    if (mode == RMHMC_METRIC_INT8_TCGEN05 && !h->i8_ok)
        return fail(h, RMHMC_E_UNSUPPORTED, "rmhmc_set_metric_mode: the INT8 tcgen05 build needs dim <= 32 and at most 16384 rows");
    if (h->n_chains > 0) free_chains(h);
    h->metric_mode = mode;
    return RMHMC_OK;
}
int rmhmc_get_metric_mode(const rmhmc_handle* h) { return h ? h->metric_mode : RMHMC_E_INVALID; }
int rmhmc_set_launch_regime(rmhmc_handle* h, int regime) {
    if (!h || regime < RMHMC_REGIME_AUTO || regime > RMHMC_REGIME_LARGE) return h ? fail(h, RMHMC_E_INVALID, "rmhmc_set_launch_regime: bad arguments") : RMHMC_E_INVALID;
    h->launch_regime = regime;
    return RMHMC_OK;
}

int rmhmc_comm_unique_id(char* out128) {
    if (!out128) return RMHMC_E_INVALID;
    NcclApi& api = nccl_api();
    if (!api.ok) { g_create_error = "libnccl.so.2 not available"; return RMHMC_E_UNSUPPORTED; }
    ncclUniqueId id;
    if (api.GetUniqueId(&id) != ncclSuccess) return RMHMC_E_CUDA;
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(out128, &id, sizeof(id));
    return RMHMC_OK;
}

int rmhmc_comm_init(rmhmc_handle* h, int world, int rank, const char* id128) {
    if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return h ? fail(h, RMHMC_E_INVALID, "rmhmc_comm_init: bad arguments") : RMHMC_E_INVALID;
    if (h->n_chains > 0) return fail(h, RMHMC_E_STATE, "rmhmc_comm_init: call before chains_init");
    NcclApi& api = nccl_api();
    if (!api.ok) return fail(h, RMHMC_E_UNSUPPORTED, "rmhmc_comm_init: libnccl.so.2 not available");
    CUDA_TRY(h, cudaSetDevice(h->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclResult_t r = api.CommInitRank(&h->comm, world, id, rank);
    if (r != ncclSuccess) { h->comm = nullptr; return fail(h, RMHMC_E_CUDA, std::string("ncclCommInitRank: ") + api.GetErrorString(r)); }
    h->shard_world = world; h->shard_rank = rank;
    return RMHMC_OK;
}

int rmhmc_stats_comm_init(rmhmc_handle* h, int world, int rank, const char* id128) {
    if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return h ? fail(h, RMHMC_E_INVALID, "rmhmc_stats_comm_init: bad arguments") : RMHMC_E_INVALID;
    NcclApi& api = nccl_api();
    if (!api.ok) return fail(h, RMHMC_E_UNSUPPORTED, "rmhmc_stats_comm_init: libnccl.so.2 not available");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->stats_comm) { api.CommDestroy(h->stats_comm); h->stats_comm = nullptr; }
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclResult_t r = api.CommInitRank(&h->stats_comm, world, id, rank);
    if (r != ncclSuccess) { h->stats_comm = nullptr; return fail(h, RMHMC_E_CUDA, std::string("ncclCommInitRank: ") + api.GetErrorString(r)); }
    h->stats_world = world;
    return RMHMC_OK;
}

int rmhmc_stats_gather(rmhmc_handle* h, const double* ess, int64_t n_chains, const double* samples, int64_t n_samples,
                       int64_t chain_stride, int64_t row_stride, double* ess_sum, double* rhat, double* scalars, int n_scalars) {
    if (!h || n_chains <= 0 || n_scalars < 0 || (n_scalars > 0 && !scalars) || (rhat && (!samples || n_samples < 2)))
        return h ? fail(h, RMHMC_E_INVALID, "rmhmc_stats_gather: bad arguments") : RMHMC_E_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int D = h->dim;
    // one exchange: [ess sums D | Rhat moments 3 D | chain count 1 | caller scalars]
    const size_t n = (size_t)4 * D + 1 + n_scalars;
    double* buf = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&buf, n * 8));
    CUDA_TRY(h, cudaMemsetAsync(buf, 0, n * 8, h->stream));
    if (ess && ess_sum) k_ess_colsum<<<D, kEssThreads, 0, h->stream>>>(ess, (long long)n_chains, D, buf);
    if (rhat) k_rhat<<<D, kEssThreads, 0, h->stream>>>(samples, (size_t)chain_stride, (size_t)row_stride, (int)n_chains, (int)n_samples, nullptr, buf + D);
    k_fill<<<1, 32, 0, h->stream>>>(buf + 4 * D, 1, (double)n_chains);
    if (n_scalars) cudaMemcpyAsync(buf + 4 * D + 1, scalars, (size_t)n_scalars * 8, cudaMemcpyDeviceToDevice, h->stream);
    int rc = RMHMC_OK;
    if (h->stats_comm) {
        ncclResult_t r = nccl_api().AllReduce(buf, buf, n, ncclDouble, ncclSum, h->stats_comm, h->stream);
        if (r != ncclSuccess) rc = fail(h, RMHMC_E_CUDA, std::string("ncclAllReduce: ") + nccl_api().GetErrorString(r));
    }
    if (!rc) {
        double c_total = 0.0;
        cudaMemcpyAsync(&c_total, buf + 4 * D, 8, cudaMemcpyDeviceToHost, h->stream);
        cudaStreamSynchronize(h->stream);
        if (ess && ess_sum) cudaMemcpyAsync(ess_sum, buf, (size_t)D * 8, cudaMemcpyDeviceToDevice, h->stream);
        if (rhat) k_rhat_finish<<<blocks_for(D, 128), 128, 0, h->stream>>>(buf + D, D, c_total, (int)n_samples, rhat);
        if (n_scalars) cudaMemcpyAsync(scalars, buf + 4 * D + 1, (size_t)n_scalars * 8, cudaMemcpyDeviceToDevice, h->stream);
        h->launches += 3;
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { h->err = std::string("rmhmc_stats_gather: ") + cudaGetErrorString(e); rc = RMHMC_E_CUDA; }
    }
    cudaFree(buf);
    return rc;
}

int blr_device_peaks(int device, void* cuda_stream, double* fp64_dmma_tflops, double* int8_tcgen05_tops) {
    if (cudaSetDevice(device) != cudaSuccess) return RMHMC_E_CUDA;
    return measure_device_peaks(reinterpret_cast<cudaStream_t>(cuda_stream), fp64_dmma_tflops, int8_tcgen05_tops) == cudaSuccess ? RMHMC_OK : RMHMC_E_CUDA;
}

int rmhmc_set_stream(rmhmc_handle* h, void* cuda_stream) {
    if (!h) return RMHMC_E_INVALID;
    h->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return RMHMC_OK;
}

// ---------------------------------------------------------------------- seams
int rmhmc_metric(rmhmc_handle* h, int64_t C, const double* theta, double* G, double* grad, double* logjoint) {
    if (!h || C <= 0 || !theta) return h ? fail(h, RMHMC_E_INVALID, "rmhmc_metric: bad arguments") : RMHMC_E_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    std::vector<void*> tmp;
    double *gp = nullptr, *graw = nullptr, *ll = nullptr;
    int rc = dev_alloc(h, &gp, (size_t)C * h->p2p, &tmp) | dev_alloc(h, &graw, (size_t)C * h->dim, &tmp) |
             dev_alloc(h, &ll, (size_t)C, &tmp);
    if (!rc) {
        MetricArgs a = metric_args(h, C, theta, gp, graw, ll, nullptr);
        rc = launch_metric<2>(h, a);                       // gradient + log-likelihood
        if (!rc && use_i8(h)) {                            // G through the digit planes, as the engine builds it
            signed char* a8 = nullptr;
            const int64_t a_rows = pad_up((int)C, kI8TileM);
            CUtensorMap map_a;
            rc = dev_alloc(h, &a8, (size_t)h->i8_slices * a_rows * h->i8_kp, &tmp);
            if (!rc && !make_tensor_map_u8_k64(&map_a, a8, (uint64_t)h->i8_slices * a_rows, (uint64_t)h->i8_kp, kI8TileM))
                rc = fail(h, RMHMC_E_CUDA, "cuTensorMapEncodeTiled failed");
            if (!rc && is_big(h)) {
                double* vbuf = nullptr;
                rc = dev_alloc(h, &vbuf, (size_t)C * h->n_rows_pad, &tmp);
                MetricArgs av = metric_args(h, C, theta, nullptr, nullptr, nullptr, nullptr);
                av.vout = vbuf;
                if (!rc) rc = launch_metric<5>(h, av);
                if (!rc) rc = i8_gemm_from_v(h, C, vbuf, a8, a_rows, map_a, gp);
            } else if (!rc) {
                rc = i8_build(h, C, theta, a8, a_rows, map_a, gp, nullptr);
            }
        } else if (!rc) { a.grad_out = nullptr; a.loglik_out = nullptr; rc = launch_metric<0>(h, a); }   // G
    }
    if (!rc) {
        if (G) k_unpack_g<<<blocks_for(C * h->dim * h->dim, 256), 256, 0, h->stream>>>(gp, G, C, h->dim, h->p2p);
        if (grad || logjoint)
            k_seam_finish<<<blocks_for(C, 128), 128, 0, h->stream>>>(theta, graw, ll, grad, logjoint, C, h->dim, h->alpha);
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { h->err = std::string("rmhmc_metric: ") + cudaGetErrorString(e); rc = RMHMC_E_CUDA; }
    }
    for (void* p : tmp) cudaFree(p);
    return rc;
}

int rmhmc_metric_partials(rmhmc_handle* h, int64_t C, const double* theta, double* dG, double* trace) {
    if (!h || C <= 0 || !theta) return h ? fail(h, RMHMC_E_INVALID, "rmhmc_metric_partials: bad arguments") : RMHMC_E_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    std::vector<void*> tmp;
    double *gp = nullptr, *graw = nullptr, *ll = nullptr, *cbuf = nullptr, *tp = nullptr;
    int* cur = nullptr;
    int64_t cpad = pad_up((int)C, kTbChains);
    int rc = dev_alloc(h, &gp, (size_t)C * h->p2p, &tmp) | dev_alloc(h, &graw, (size_t)C * h->dim, &tmp) |
             dev_alloc(h, &ll, (size_t)C, &tmp) | dev_alloc(h, &cbuf, (size_t)cpad * h->n_rows_pad, &tmp) |
             dev_alloc(h, &tp, (size_t)C * h->p3p, &tmp) | dev_alloc(h, &cur, (size_t)C, &tmp);
    if (!rc) {
        MetricArgs a = metric_args(h, C, theta, gp, graw, ll, cbuf);
        rc = launch_metric<1>(h, a);
    }
    if (!rc) rc = launch_tbuild(h, C, cbuf, tp, cur, 0, 0);
    if (!rc) {
        if (dG) k_unpack_t<<<blocks_for(C * h->dim * h->dim * h->dim, 256), 256, 0, h->stream>>>(tp, dG, C, h->dim, h->p3p);
        if (trace) {
            EngineParams P = h->P;
            P.n_chains = (int)C; P.dim = h->dim; P.ds = h->dim | 1; P.p2 = h->p2; P.p2p = h->p2p; P.p3 = h->p3;
            P.p3p = h->p3p; P.tidx = h->tidx; P.tidx32 = h->tidx32; P.pair_a = h->pair_a; P.pair_b = h->pair_b;
            rc = set_chain_smem_attrs(h);
            if (!rc) rc = launch_seam_factor(h, P, C, gp, tp, nullptr, nullptr, nullptr, trace);
        }
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { h->err = std::string("rmhmc_metric_partials: ") + cudaGetErrorString(e); rc = RMHMC_E_CUDA; }
    }
    for (void* p : tmp) cudaFree(p);
    return rc;
}

int rmhmc_chol_logdet(rmhmc_handle* h, int64_t C, const double* G, double* L, double* Ginv, double* logdet) {
    if (!h || C <= 0 || !G) return h ? fail(h, RMHMC_E_INVALID, "rmhmc_chol_logdet: bad arguments") : RMHMC_E_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    std::vector<void*> tmp;
    double* gp = nullptr;
    int rc = dev_alloc(h, &gp, (size_t)C * h->p2p, &tmp);
    if (!rc) {
        k_pack_g<<<blocks_for(C * h->p2p, 256), 256, 0, h->stream>>>(G, gp, C, h->dim, h->p2, h->p2p, h->pair_a, h->pair_b);
        EngineParams P = h->P;
        P.n_chains = (int)C; P.dim = h->dim; P.ds = h->dim | 1; P.p2 = h->p2; P.p2p = h->p2p; P.p3 = h->p3;
        P.p3p = h->p3p; P.tidx = h->tidx; P.tidx32 = h->tidx32; P.pair_a = h->pair_a; P.pair_b = h->pair_b;
        rc = set_chain_smem_attrs(h);
        if (!rc) rc = launch_seam_factor(h, P, C, gp, nullptr, L, Ginv, logdet, nullptr);
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { h->err = std::string("rmhmc_chol_logdet: ") + cudaGetErrorString(e); rc = RMHMC_E_CUDA; }
    }
    for (void* p : tmp) cudaFree(p);
    return rc;
}

// ---------------------------------------------------------------------- engine
static int chains_init_common(rmhmc_handle* h, int64_t C, const double* theta0, bool hmc) {
    if (!h || C <= 0) return h ? fail(h, RMHMC_E_INVALID, "chains_init: bad arguments") : RMHMC_E_INVALID;
    if (C > 0x7fffffff / 64) return fail(h, RMHMC_E_INVALID, "chains_init: too many chains");
    CUDA_TRY(h, cudaSetDevice(h->device));
    h->matrix_free = h->matrix_free_user;         // a previous mMALA chain set may have forced the matrix-free partials
    int rc = alloc_chains(h, C, hmc);
    if (rc) return rc;
    ChainArrays& S = h->S;
    if (theta0)
        CUDA_TRY(h, cudaMemcpyAsync(S.theta_w, theta0, (size_t)C * h->dim * 8, cudaMemcpyDeviceToDevice, h->stream));
    else
        k_fill<<<blocks_for(C * h->dim, 256), 256, 0, h->stream>>>(S.theta_w, C * h->dim, hmc ? 0.0 : 1e-3);
    // trace pointers / samples refer to the previous chain set
    h->P.samples = nullptr; h->P.tr_theta_steps = nullptr; h->P.tr_mom_end = nullptr; h->P.tr_theta_end = nullptr;
    h->P.tr_mom0 = nullptr; h->P.tr_hcur = nullptr; h->P.tr_hprop = nullptr; h->P.tr_flags = nullptr; h->P.tr_iters = 0;
    h->rng_set = false;
    if (hmc) {
        MetricArgs a = metric_args(h, C, S.theta_w, nullptr, S.grad_tmp, S.loglik_tmp, nullptr);
        rc = launch_metric<2>(h, a);
        if (rc) return rc;
        rc = reduce_build<2>(h);
        if (rc) return rc;
        k_hmc_back<<<(unsigned)C, 32, 0, h->stream>>>(h->P, S, 1);
    } else {
        rc = set_chain_smem_attrs(h);
        if (rc) return rc;
        rc = build_metric_closing(h, 0);
        if (rc) return rc;
        rc = reduce_build<1>(h);
        if (rc) return rc;
        if (h->matrix_free) {
            rc = launch_factor(h, 1);
            if (!rc) rc = mf_closing_passes(h, 1);
            if (!rc) rc = launch_mf_turn(h, 1, 0, 1);
        } else {
            rc = build_partials(h, 0);
            if (!rc) rc = launch_factor(h, 1);
            if (!rc) rc = launch_turn(h, 1, 0, 1);
        }
        if (rc) return rc;
    }
    h->launches += 1;
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return RMHMC_OK;
}

int rmhmc_chains_init(rmhmc_handle* h, int64_t C, const double* theta0) { return chains_init_common(h, C, theta0, false); }
int hmc_chains_init(rmhmc_handle* h, int64_t C, const double* theta0) { return chains_init_common(h, C, theta0, true); }

int rmhmc_configure(rmhmc_handle* h, int n_leapfrog, double step_size, int n_fixed) {
    if (!h || n_leapfrog < 1 || n_fixed < 0 || !(step_size > 0)) return h ? fail(h, RMHMC_E_INVALID, "rmhmc_configure: bad arguments") : RMHMC_E_INVALID;
    h->P.n_leapfrog = n_leapfrog; h->P.step_size = step_size; h->P.n_fixed = n_fixed;
    h->configured = true;
    return RMHMC_OK;
}
int hmc_configure(rmhmc_handle* h, int n_leapfrog, double step_size) {
    if (!h || n_leapfrog < 1 || !(step_size > 0)) return h ? fail(h, RMHMC_E_INVALID, "hmc_configure: bad arguments") : RMHMC_E_INVALID;
    h->P.n_leapfrog = n_leapfrog; h->P.step_size = step_size; h->P.n_fixed = 0;
    h->configured = true;
    return RMHMC_OK;
}

int rmhmc_set_tape(rmhmc_handle* h, int64_t it_base, int64_t n_window, const double* z, const double* u_step,
                   const double* z_dir, const double* u_acc) {
    if (!h || n_window <= 0 || !z || !u_step || !z_dir || !u_acc) return h ? fail(h, RMHMC_E_INVALID, "rmhmc_set_tape: bad arguments") : RMHMC_E_INVALID;
    h->P.rng_mode = 0; h->P.tape_base = it_base; h->P.tape_z = z; h->P.tape_u_step = u_step;
    h->P.tape_z_dir = z_dir; h->P.tape_u_acc = u_acc;
    h->tape_base = it_base; h->tape_window = n_window;
    h->rng_set = true;
    return RMHMC_OK;
}
int hmc_set_tape(rmhmc_handle* h, int64_t it_base, int64_t n_window, const double* z, const double* u_step,
                 const double* u_acc) {
    if (!h || n_window <= 0 || !z || !u_step || !u_acc) return h ? fail(h, RMHMC_E_INVALID, "hmc_set_tape: bad arguments") : RMHMC_E_INVALID;
    h->P.rng_mode = 0; h->P.tape_base = it_base; h->P.tape_z = z; h->P.tape_u_step = u_step;
    h->P.tape_z_dir = nullptr; h->P.tape_u_acc = u_acc;
    h->tape_base = it_base; h->tape_window = n_window;
    h->rng_set = true;
    return RMHMC_OK;
}
int rmhmc_set_momentum_family(rmhmc_handle* h, int family) {
    if (!h || (family != RMHMC_MOMENTUM_GAUSSIAN && family != RMHMC_MOMENTUM_STUDENT_T))
        return h ? fail(h, RMHMC_E_INVALID, "rmhmc_set_momentum_family: bad arguments") : RMHMC_E_INVALID;
    if (family == RMHMC_MOMENTUM_STUDENT_T && (!h->matrix_free || is_big(h) || h->comm))
        return fail(h, RMHMC_E_UNSUPPORTED, "rmhmc_set_momentum_family: Student-t needs the matrix-free partials, dim <= 32 and unsharded data");
    h->P.student_t = family == RMHMC_MOMENTUM_STUDENT_T ? 1 : 0;
    return RMHMC_OK;
}
int rmhmc_set_tape_chi(rmhmc_handle* h, const double* z_chi) {
    if (!h || !z_chi) return h ? fail(h, RMHMC_E_INVALID, "rmhmc_set_tape_chi: bad arguments") : RMHMC_E_INVALID;
    h->P.tape_z_chi = z_chi;
    return RMHMC_OK;
}
int rmhmc_set_philox(rmhmc_handle* h, uint64_t seed, int64_t chain_offset) {
    if (!h) return RMHMC_E_INVALID;
    h->P.rng_mode = 1; h->P.seed = seed; h->P.chain_offset = chain_offset;
    h->tape_window = 0;
    h->rng_set = true;
    return RMHMC_OK;
}

int rmhmc_set_samples(rmhmc_handle* h, double* samples, int64_t capacity, int64_t burn_in) {
    if (!h || (samples && capacity <= 0)) return h ? fail(h, RMHMC_E_INVALID, "rmhmc_set_samples: bad arguments") : RMHMC_E_INVALID;
    h->P.samples = samples; h->P.sample_cap = capacity; h->P.burn_in = burn_in;
    return RMHMC_OK;
}

int rmhmc_set_trace(rmhmc_handle* h, int64_t n_iters, double* theta_steps, double* mom_end, double* theta_end,
                    double* mom0, double* h_current, double* h_proposed, int32_t* flags) {
    if (!h) return RMHMC_E_INVALID;
    if (theta_steps && (!mom_end || !theta_end || !mom0 || !h_current || !h_proposed || !flags || n_iters <= 0))
        return fail(h, RMHMC_E_INVALID, "rmhmc_set_trace: all buffers are required");
    h->P.tr_iters = theta_steps ? n_iters : 0;
    h->P.tr_theta_steps = theta_steps; h->P.tr_mom_end = theta_steps ? mom_end : nullptr;
    h->P.tr_theta_end = theta_steps ? theta_end : nullptr; h->P.tr_mom0 = theta_steps ? mom0 : nullptr;
    h->P.tr_hcur = theta_steps ? h_current : nullptr; h->P.tr_hprop = theta_steps ? h_proposed : nullptr;
    h->P.tr_flags = theta_steps ? flags : nullptr;
    return RMHMC_OK;
}

int rmhmc_advance(rmhmc_handle* h, int64_t n_rounds, int64_t it_stop) {
    if (!h || n_rounds < 0) return h ? fail(h, RMHMC_E_INVALID, "rmhmc_advance: bad arguments") : RMHMC_E_INVALID;
    if (h->n_chains <= 0 || h->is_hmc || h->is_mmala) return fail(h, RMHMC_E_STATE, "rmhmc_advance: call rmhmc_chains_init first");
    if (!h->configured || !h->rng_set) return fail(h, RMHMC_E_STATE, "rmhmc_advance: configure and set a tape / philox seed first");
    CUDA_TRY(h, cudaSetDevice(h->device));
    // a host tape bounds the iterations that may run: chains idle once they reach its end
    if (h->P.rng_mode == 0 && it_stop > h->tape_base + h->tape_window) it_stop = h->tape_base + h->tape_window;
    h->P.it_stop = it_stop;
    return rmhmc_rounds(h, n_rounds);
}

// free-running rounds of the other samplers: HMC round = one leapfrog step of every chain (hmc.py:51-62), mMALA round =
// one iteration of every chain
int hmc_advance(rmhmc_handle* h, int64_t n_rounds, int64_t it_stop) {
    if (!h || n_rounds < 0) return h ? fail(h, RMHMC_E_INVALID, "hmc_advance: bad arguments") : RMHMC_E_INVALID;
    if (h->n_chains <= 0 || !h->is_hmc) return fail(h, RMHMC_E_STATE, "hmc_advance: call hmc_chains_init first");
    if (!h->configured || !h->rng_set) return fail(h, RMHMC_E_STATE, "hmc_advance: configure and set a tape / philox seed first");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->P.rng_mode == 0 && it_stop > h->tape_base + h->tape_window) it_stop = h->tape_base + h->tape_window;
    h->P.it_stop = it_stop;
    return hmc_rounds(h, n_rounds);
}
int hmc_set_fused(rmhmc_handle* h, int fused, int rounds_per_launch) {
    if (!h || rounds_per_launch < 0) return h ? fail(h, RMHMC_E_INVALID, "hmc_set_fused: bad arguments") : RMHMC_E_INVALID;
    h->hmc_fused = fused != 0;
    if (rounds_per_launch > 0) h->hmc_rounds_per_launch = rounds_per_launch;
    return RMHMC_OK;
}
int mmala_advance(rmhmc_handle* h, int64_t n_rounds, int64_t it_stop) {
    if (!h || n_rounds < 0) return h ? fail(h, RMHMC_E_INVALID, "mmala_advance: bad arguments") : RMHMC_E_INVALID;
    if (h->n_chains <= 0 || !h->is_mmala) return fail(h, RMHMC_E_STATE, "mmala_advance: call mmala_chains_init first");
    if (!h->rng_set) return fail(h, RMHMC_E_STATE, "mmala_advance: set a tape / philox seed first");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->P.rng_mode == 0 && it_stop > h->tape_base + h->tape_window) it_stop = h->tape_base + h->tape_window;
    h->P.it_stop = it_stop;
    return mmala_rounds(h, n_rounds);
}

int rmhmc_run(rmhmc_handle* h, int64_t it_stop, int64_t* rounds_done) {
    if (!h) return RMHMC_E_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return run_until(h, it_stop, rounds_done, false);
}
int hmc_run(rmhmc_handle* h, int64_t it_stop, int64_t* rounds_done) {
    if (!h) return RMHMC_E_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return run_until(h, it_stop, rounds_done, true);
}

int mmala_chains_init(rmhmc_handle* h, int64_t C, const double* theta0, int simplified, double step_size) {
    if (!h || C <= 0 || !(step_size > 0)) return h ? fail(h, RMHMC_E_INVALID, "mmala_chains_init: bad arguments") : RMHMC_E_INVALID;
    if (h->dim > kMaxDimWarp) return fail(h, RMHMC_E_UNSUPPORTED, "mmala: dim > 32 is not supported");
    if (!simplified && !h->kr2t) return fail(h, RMHMC_E_UNSUPPORTED, "mmala: the full drift needs the matrix-free partials (KR2(X)^T resident)");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const bool saved_mf = h->matrix_free_user;
    h->matrix_free = simplified ? saved_mf : true;       // restored by the next rmhmc / hmc chains_init (matrix_free_user)
    int rc = alloc_chains(h, C, false);
    if (rc) { h->matrix_free = saved_mf; return rc; }
    h->is_mmala = true;
    h->mmala_simplified = simplified < 0 || simplified > 2 ? 1 : simplified;       // 0 full mMALA, 1 simplified, 2 IWLS proposal
    h->P.step_size = step_size; h->P.n_leapfrog = 1; h->P.n_fixed = 0;       // the cached drift / proposal factor depend on eps
    h->configured = true;
    ChainArrays& S = h->S;
    if (theta0)
        CUDA_TRY(h, cudaMemcpyAsync(S.theta_w, theta0, (size_t)C * h->dim * 8, cudaMemcpyDeviceToDevice, h->stream));
    else
        CUDA_TRY(h, cudaMemsetAsync(S.theta_w, 0, (size_t)C * h->dim * 8, h->stream));      // w = zeros(D,1), BLR_mMALA.m:165
    h->P.samples = nullptr; h->P.tr_theta_steps = nullptr; h->P.tr_mom_end = nullptr; h->P.tr_theta_end = nullptr;
    h->P.tr_mom0 = nullptr; h->P.tr_hcur = nullptr; h->P.tr_hprop = nullptr; h->P.tr_flags = nullptr; h->P.tr_iters = 0;
    h->rng_set = false;
    rc = set_chain_smem_attrs(h);
    if (!rc) rc = mmala_builds(h, 0);
    if (!rc) rc = launch_mmala_turn(h, 1, 0, 1);
    if (rc) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return RMHMC_OK;
}
int mmala_set_tape(rmhmc_handle* h, int64_t it_base, int64_t n_window, const double* z, const double* u_acc) {
    if (!h || n_window <= 0 || !z || !u_acc) return h ? fail(h, RMHMC_E_INVALID, "mmala_set_tape: bad arguments") : RMHMC_E_INVALID;
    h->P.rng_mode = 0; h->P.tape_base = it_base; h->P.tape_z = z; h->P.tape_u_step = nullptr;
    h->P.tape_z_dir = nullptr; h->P.tape_u_acc = u_acc;
    h->tape_base = it_base; h->tape_window = n_window;
    h->rng_set = true;
    return RMHMC_OK;
}
int mmala_read_proposal(rmhmc_handle* h, double* mean, double* chol_lower) {
    if (!h) return RMHMC_E_INVALID;
    if (h->n_chains <= 0 || !h->is_mmala) return fail(h, RMHMC_E_STATE, "mmala_read_proposal: call mmala_chains_init first");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int64_t C = h->n_chains;
    k_copy_proposal<<<blocks_for(C * h->dim * h->dim, 256), 256, 0, h->stream>>>(h->S.theta, h->S.grad, h->S.invg, h->S.cur, h->P.slot_theta,
                                                                                h->P.slot_invg, mean, chol_lower, C, h->dim);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return RMHMC_OK;
}
int mmala_run(rmhmc_handle* h, int64_t it_stop, int64_t* rounds_done) {
    if (!h) return RMHMC_E_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return run_until(h, it_stop, rounds_done, false, true);
}

int rmhmc_leapfrog(rmhmc_handle* h, int64_t C, const double* theta, const double* mom, const int32_t* dir,
                   const int32_t* nsteps, double step_size, int n_fixed, double* out_theta, double* out_mom,
                   double* out_h_start, double* out_h_end) {
    if (!h || C <= 0 || !theta || !mom || !dir || !nsteps || !(step_size > 0) || n_fixed < 0)
        return h ? fail(h, RMHMC_E_INVALID, "rmhmc_leapfrog: bad arguments") : RMHMC_E_INVALID;
    int rc = rmhmc_chains_init(h, C, theta);
    if (rc) return rc;
    std::vector<void*> tmp;
    double *th_steps = nullptr, *m_end = nullptr, *t_end = nullptr, *m0 = nullptr, *hc = nullptr, *hp = nullptr;
    int* fl = nullptr;
    const size_t c = (size_t)C, D = (size_t)h->dim;
    int max_steps = 1 << 12;            // bound of the per-step trace rows; steps themselves are unbounded
    rc = dev_alloc(h, &th_steps, c * D, &tmp) | dev_alloc(h, &m_end, c * D, &tmp) | dev_alloc(h, &t_end, c * D, &tmp) |
         dev_alloc(h, &m0, c * D, &tmp) | dev_alloc(h, &hc, c, &tmp) | dev_alloc(h, &hp, c, &tmp) | dev_alloc(h, &fl, c, &tmp);
    (void)max_steps;
    if (!rc) {
        EngineParams saved = h->P;
        h->P.n_leapfrog = 1;            // trace row stride; theta_steps is not reported by this seam
        h->P.step_size = step_size; h->P.n_fixed = n_fixed;
        h->P.rng_mode = 1; h->P.seed = 0; h->P.chain_offset = 0;          // only the (irrelevant) accept draw uses it
        h->P.ext_mom = mom; h->P.ext_nsteps = nsteps; h->P.ext_dir = dir;
        h->P.samples = nullptr;
        h->P.tr_iters = 1; h->P.tr_theta_steps = nullptr; h->P.tr_mom_end = m_end; h->P.tr_theta_end = t_end;
        h->P.tr_mom0 = m0; h->P.tr_hcur = hc; h->P.tr_hprop = hp; h->P.tr_flags = fl;
        h->configured = true; h->rng_set = true;
        rc = run_until(h, 1, nullptr, false);
        h->P.ext_mom = nullptr; h->P.ext_nsteps = nullptr; h->P.ext_dir = nullptr;
        h->P.n_leapfrog = saved.n_leapfrog; h->P.step_size = saved.step_size; h->P.n_fixed = saved.n_fixed;
        h->P.tr_iters = 0; h->P.tr_mom_end = nullptr; h->P.tr_theta_end = nullptr; h->P.tr_mom0 = nullptr;
        h->P.tr_hcur = nullptr; h->P.tr_hprop = nullptr; h->P.tr_flags = nullptr;
        h->rng_set = false;
        if (!rc) {
            if (out_theta) cudaMemcpyAsync(out_theta, t_end, c * D * 8, cudaMemcpyDeviceToDevice, h->stream);
            if (out_mom) cudaMemcpyAsync(out_mom, m_end, c * D * 8, cudaMemcpyDeviceToDevice, h->stream);
            if (out_h_start) cudaMemcpyAsync(out_h_start, hc, c * 8, cudaMemcpyDeviceToDevice, h->stream);
            if (out_h_end) cudaMemcpyAsync(out_h_end, hp, c * 8, cudaMemcpyDeviceToDevice, h->stream);
            cudaError_t e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) { h->err = std::string("rmhmc_leapfrog: ") + cudaGetErrorString(e); rc = RMHMC_E_CUDA; }
        }
    }
    for (void* p : tmp) cudaFree(p);
    return rc;
}

int rmhmc_read_state(rmhmc_handle* h, double* theta, int64_t* iters, int64_t* accepted, int64_t* leapfrogs,
                     int32_t* renorm_mom, int32_t* renorm_pos) {
    if (!h) return RMHMC_E_INVALID;
    if (h->n_chains <= 0) return fail(h, RMHMC_E_STATE, "rmhmc_read_state: no chains");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int64_t C = h->n_chains;
    ChainArrays& S = h->S;
    unsigned gb = blocks_for(C, 256);
    if (theta) k_copy_theta<<<blocks_for(C * h->dim, 256), 256, 0, h->stream>>>(S.theta, S.cur, h->P.slot_theta, theta, C, h->dim);
    if (iters) k_copy_i64<<<gb, 256, 0, h->stream>>>(S.iter, iters, C);
    if (accepted) k_copy_i64<<<gb, 256, 0, h->stream>>>(S.accepted, accepted, C);
    if (leapfrogs) k_copy_i64<<<gb, 256, 0, h->stream>>>(S.leapfrogs, leapfrogs, C);
    if (renorm_mom) k_copy_i32<<<gb, 256, 0, h->stream>>>(S.renorm_mom, renorm_mom, C);
    if (renorm_pos) k_copy_i32<<<gb, 256, 0, h->stream>>>(S.renorm_pos, renorm_pos, C);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return RMHMC_OK;
}

int64_t rmhmc_launch_count(const rmhmc_handle* h) { return h ? h->launches : 0; }
int64_t rmhmc_chain_generation(const rmhmc_handle* h) { return h ? h->chain_gen : -1; }

int rmhmc_profile_enable(rmhmc_handle* h, int enable) {
    if (!h) return RMHMC_E_INVALID;
    drain_profile(h);
    for (auto& s : h->prof) { s.ms = 0.0; s.launches = 0; }
    h->profiling = enable != 0;
    return RMHMC_OK;
}
int rmhmc_profile_read(rmhmc_handle* h, int kind, double* ms, int64_t* launches) {
    if (!h || kind < 0 || kind > 12) return RMHMC_E_INVALID;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    drain_profile(h);
    if (ms) *ms = h->prof[kind].ms;
    if (launches) *launches = h->prof[kind].launches;
    return RMHMC_OK;
}

int blr_ess_batched(int device, void* cuda_stream, const double* samples, int64_t n_chains, int64_t n_samples,
                    int dim, int64_t chain_stride, int64_t row_stride, int64_t max_lag, double* ess) {
    if (!samples || !ess || n_chains <= 0 || n_samples < 2 || dim <= 0 || max_lag < 1 || n_samples > 0x7fffffff / 2)
        return RMHMC_E_INVALID;
    if (dim > 65535) return RMHMC_E_UNSUPPORTED;
    if (cudaSetDevice(device) != cudaSuccess) return RMHMC_E_CUDA;
    int n_fft = 1;
    while (n_fft < n_samples) n_fft *= 2;           // tools.py:16-19
    n_fft += 1;                                     // tools.py:23
    if (max_lag > n_fft - 1) return RMHMC_E_INVALID;         // tools.py:26: the reference slices the nFFT-point circular ACF
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    dim3 grid((unsigned)n_chains, (unsigned)dim);
    if (n_samples <= 24000) {
        size_t smem = (size_t)n_samples * 8;
        if (cudaFuncSetAttribute(k_ess, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return RMHMC_E_CUDA;
        k_ess<<<grid, kEssThreads, smem, st>>>(samples, (size_t)chain_stride, (size_t)row_stride, (int)n_samples, (int)max_lag, n_fft,
                                               ess, dim, nullptr, nullptr, nullptr);
        return cudaGetLastError() == cudaSuccess ? RMHMC_OK : RMHMC_E_CUDA;
    }
    // longer series: the centred copy goes to a global scratch row per (chain, parameter)
    double* scratch = nullptr;
    if (cudaMalloc((void**)&scratch, (size_t)n_chains * dim * n_samples * 8) != cudaSuccess) return RMHMC_E_CUDA;
    k_ess<<<grid, kEssThreads, 0, st>>>(samples, (size_t)chain_stride, (size_t)row_stride, (int)n_samples, (int)max_lag, n_fft, ess,
                                        dim, nullptr, nullptr, scratch);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(scratch);
    return e == cudaSuccess ? RMHMC_OK : RMHMC_E_CUDA;
}

int blr_autocorr(int device, void* cuda_stream, const double* series, int64_t n_series, int64_t n_samples,
                 int64_t n_lag, double* acf) {
    if (!series || !acf || n_series <= 0 || n_samples < 2 || n_lag < 0) return RMHMC_E_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return RMHMC_E_CUDA;
    int n_fft = 1;
    while (n_fft < n_samples) n_fft *= 2;           // tools.py:16-19
    n_fft += 1;                                     // tools.py:23
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    dim3 grid((unsigned)(n_lag + 1), (unsigned)n_series);
    k_acf_raw<<<grid, kEssThreads, 0, st>>>(series, (int)n_samples, n_fft, (int)n_lag, acf);
    int64_t total = n_series * (n_lag + 1);
    k_acf_normalise<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(acf, (int)n_lag, (int)n_series);
    k_acf_lag0<<<(unsigned)((n_series + 255) / 256), 256, 0, st>>>(acf, (int)n_lag, (int)n_series);
    return cudaGetLastError() == cudaSuccess ? RMHMC_OK : RMHMC_E_CUDA;
}

int blr_rhat(int device, void* cuda_stream, const double* samples, int64_t n_chains, int64_t n_samples, int dim,
             int64_t chain_stride, int64_t row_stride, double* rhat) {
    if (!samples || !rhat || n_chains < 2 || n_samples < 2 || dim <= 0) return RMHMC_E_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return RMHMC_E_CUDA;
    k_rhat<<<(unsigned)dim, kEssThreads, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
        samples, (size_t)chain_stride, (size_t)row_stride, (int)n_chains, (int)n_samples, rhat);
    return cudaGetLastError() == cudaSuccess ? RMHMC_OK : RMHMC_E_CUDA;
}

int blr_ess_ragged(int device, void* cuda_stream, const double* samples, int64_t n_chains, int64_t max_samples,
                   int dim, int64_t chain_stride, int64_t row_stride, const int64_t* starts, const int64_t* counts,
                   double* ess) {
    if (!samples || !ess || !starts || !counts || n_chains <= 0 || max_samples < 2 || dim <= 0) return RMHMC_E_INVALID;
    if (max_samples > 24000 || dim > 65535) return RMHMC_E_UNSUPPORTED;
    if (cudaSetDevice(device) != cudaSuccess) return RMHMC_E_CUDA;
    size_t smem = (size_t)max_samples * 8;
    if (cudaFuncSetAttribute(k_ess, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return RMHMC_E_CUDA;
    dim3 grid((unsigned)n_chains, (unsigned)dim);
    k_ess<<<grid, kEssThreads, smem, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
        samples, (size_t)chain_stride, (size_t)row_stride, (int)max_samples, 0, 0, ess, dim,
        reinterpret_cast<const long long*>(starts), reinterpret_cast<const long long*>(counts));
    return cudaGetLastError() == cudaSuccess ? RMHMC_OK : RMHMC_E_CUDA;
}

}  // extern "C"
