// Metric-partials build: T[chain, (i,j,k)] = sum_n c_n x_ni x_nj x_nk, c_n = v_n (1 - 2 p_n).
//
// The reference forms the D matrices dG/dw_d = X^T diag(c x_d) X one by one (rmhmc.py:64-75,
// :142-153).  Stacked they are a fully symmetric 3-tensor, so only the P3 = D(D+1)(D+2)/6 packed
// entries i<=j<=k are computed (4-5x fewer flops), as ONE chain-batched contraction on the FP64
// tensor cores (DMMA.8x8x4):
//     T[chains x P3] = Cw[chains x rows] . KR3(X)[rows x P3],  KR3(X)[n,(i,j,k)] = x_ni x_nj x_nk
// KR3 is never written to global memory: while a warp runs the DMMAs of row block rb it also forms its
// 2-row slice of the [32 rows x 128 columns] KR3 tile of block rb+1 in shared memory, from which all 16
// warps then load their B fragments.  Measured on B200 (profiles/microbench/dmma_dmul_mix.cu): a DMUL
// whose result feeds a DMMA costs ~5.4 FP64-pipe cycles instead of 2, and with per-lane operand
// formation the four chain-group warps repeated the same products -- 8 dependent DMULs per 16 DMMAs
// capped the previous version at 81 % of the tensor pipe; a dedicated producer warp per sub-partition
// was starved by the DMMA warps' pipe arbitration (66 %).  Cw tiles (written by the closing metric
// build) and X row blocks stream through 1-D bulk TMA into an mbarrier ring.  This kernel carries ~60%
// of the algorithmic flops.
#pragma once
#include "common.cuh"

namespace rmhmc {

constexpr int kTbChains = 128;    // chains per CTA (4 chain groups of 32)
constexpr int kTbCols = 128;      // packed-triple columns per CTA
constexpr int kTbRows = 32;       // rows per staged block (K tile)
constexpr int kTbStages = 3;         // ring depth (2 when the X rows are too wide for shared memory)
constexpr int kTbAS = kTbRows + 4;   // smem stride of the Cw tile (4*odd: conflict-free A fragments)
constexpr int kTbKS = kTbCols + 4;   // smem stride of the KR3 tile (= 4 mod 16: conflict-free B fragments)
constexpr int kTbWarps = 16;         // 4 (chains) x 4 (columns) warps of 32 x 32
constexpr int kTbThreads = kTbWarps * 32;

struct TBuildArgs {
    const double* kr3;        // [Np][P3p] precomputed KR3(X) (small problems) or null
    const double* x;          // [Np][XS]
    const uchar4* tri_tab;    // [P3p] packed column -> (i, j, k)
    const double* cbuf;       // [Cpad][Np]
    double* tpack;            // [2][C][P3p]
    const int* cur;           // [C] current slot per chain
    int flip;                 // write to slot cur ^ flip
    size_t slot_stride;       // doubles between the two T slots
    int n_chains, n_rows_pad, xs, p3, p3p;
    // k_tbuild_pre only: gridDim.z > 1 splits K (the rows) over z; split z writes its PARTIAL product at
    // tpack + z * split_stride (no slot addressing), to be added in split order by k_reduce_splits
    size_t split_stride;
};

__host__ inline size_t tbuild_smem_bytes(int xs, int stages) {
    return ((size_t)stages * ((size_t)kTbChains * kTbAS + (size_t)kTbRows * xs) + 2 * (size_t)kTbRows * kTbKS) * 8 + 16 * 8;
}
__host__ inline int tbuild_stages(int xs) { return tbuild_smem_bytes(xs, kTbStages) <= 227 * 1024 ? kTbStages : 2; }

#ifdef __CUDACC__
template <int ST>
__global__ void __launch_bounds__(kTbThreads, 1) k_tbuild(TBuildArgs a) {
    constexpr int MC = kTbChains, KB = kTbRows, AS = kTbAS, KS = kTbKS, NW = kTbWarps;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int xs = a.xs;
    double* a_ring = reinterpret_cast<double*>(smem_raw);                 // [ST][MC][AS]   Cw tiles
    double* x_ring = a_ring + (size_t)ST * MC * AS;                       // [ST][KB][xs]   X row blocks
    double* kr = x_ring + (size_t)ST * KB * xs;                           // [2][KB][KS]    KR3 tiles
    uint64_t* full = reinterpret_cast<uint64_t*>(kr + 2 * (size_t)KB * KS);   // [ST] TMA landed (Cw + X of a block)
    uint64_t* empty = full + ST;                                          // [ST] block's Cw consumed by all warps
    uint64_t* kr_full = empty + ST;                                       // [2]  KR3 tile written by all warps
    uint64_t* kr_empty = kr_full + 2;                                     // [2]  KR3 tile consumed by all warps

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int chain0 = blockIdx.y * MC;
    if (chain0 >= a.n_chains) return;
    const int n_blocks = a.n_rows_pad / KB;
    const uint32_t x_bytes = (uint32_t)(KB * xs * 8);
    const uint32_t stage_bytes = x_bytes + (uint32_t)(MC * KB * 8);

    if (tid == 0) {
        for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
        for (int s = 0; s < 2; ++s) { mbar_init(&kr_full[s], NW); mbar_init(&kr_empty[s], NW); }
        mbar_fence_init();
    }
    __syncthreads();

    // every warp loads 8 chain rows of the Cw tile; thread 0 also arms the barrier and loads the X block
    auto issue = [&](int rb) {
        const int stage = rb % ST;
        if (tid == 0) {
            mbar_expect_tx(&full[stage], stage_bytes);
            tma_bulk_g2s(x_ring + (size_t)stage * KB * xs, a.x + (size_t)rb * KB * xs, x_bytes, &full[stage]);
        }
        if (lane < MC / NW) {
            const int m = warp * (MC / NW) + lane;
            tma_bulk_g2s(a_ring + ((size_t)stage * MC + m) * AS,
                         a.cbuf + (size_t)(chain0 + m) * a.n_rows_pad + (size_t)rb * KB, (uint32_t)(KB * 8), &full[stage]);
        }
    };
    for (int s = 0; s < 2 && s < n_blocks; ++s) issue(s);

    // KR3 production: warp w forms rows 2w, 2w+1 of the tile; lane l the columns l, l+32, l+64, l+96
    int ci[4], cj[4], ck[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int col = blockIdx.x * kTbCols + lane + 32 * i;
        uchar4 t = col < a.p3p ? a.tri_tab[col] : make_uchar4(0, 0, 0, 0);
        ci[i] = t.x; cj[i] = t.y; ck[i] = t.z;
    }
    auto produce = [&](int rb) {            // KR3 tile of block rb from its X rows (which must have landed)
        const double* xb = x_ring + (size_t)(rb % ST) * KB * xs;
        double* dst = kr + (size_t)(rb & 1) * KB * KS;
#pragma unroll
        for (int rr = 0; rr < KB / NW; ++rr) {
            const int r = warp * (KB / NW) + rr;
            const double* xr = xb + (size_t)r * xs;
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[(size_t)r * KS + lane + 32 * i] = xr[ci[i]] * xr[cj[i]] * xr[ck[i]];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&kr_full[rb & 1]);
    };
    mbar_wait(&full[0], 0);
    produce(0);

    const int wm = warp & 3, wn = warp >> 2;
    double acc[4][4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[m][nt][0] = acc[m][nt][1] = 0.0;

    for (int rb = 0; rb < n_blocks; ++rb) {
        const int stage = rb % ST, buf = rb & 1;
        mbar_wait(&full[stage], (uint32_t)((rb / ST) & 1));
        mbar_wait(&kr_full[buf], (uint32_t)((rb >> 1) & 1));
        const double* as = a_ring + (size_t)stage * MC * AS + (size_t)(wm * 32 + g) * AS + q;
        const double* kb = kr + (size_t)buf * KB * KS + (size_t)q * KS + wn * 32 + g;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll 2
            for (int ks = half * (KB / 8); ks < (half + 1) * (KB / 8); ++ks) {
                double af[4], bf[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) af[m] = as[(size_t)m * 8 * AS + ks * 4];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) bf[nt] = kb[(size_t)ks * 4 * KS + nt * 8];
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) dmma884(acc[m][nt][0], acc[m][nt][1], af[m], bf[nt]);
            }
            if (half == 0 && rb + 1 < n_blocks) {
                // form this warp's slice of the next block's KR3 tile between the two DMMA halves
                mbar_wait(&full[(rb + 1) % ST], (uint32_t)(((rb + 1) / ST) & 1));
                if (rb >= 1) mbar_wait(&kr_empty[(rb + 1) & 1], (uint32_t)(((rb - 1) >> 1) & 1));
                produce(rb + 1);
            }
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&kr_empty[buf]);
            mbar_arrive(&empty[stage]);
        }
        // block rb+2 goes into the stage of block rb+2-ST once everybody has released that block
        if (rb + 2 < n_blocks) {
            const int prev = rb + 2 - ST;          // ST = 3: block rb-1 (released long ago); ST = 2: block rb
            if (prev >= 0) mbar_wait(&empty[prev % ST], (uint32_t)((prev / ST) & 1));
            issue(rb + 2);
        }
    }

    const int col0 = blockIdx.x * kTbCols + wn * 32;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        int c = chain0 + wm * 32 + m * 8 + g;
        if (c >= a.n_chains) continue;
        double* out = a.tpack + (size_t)(a.cur[c] ^ a.flip) * a.slot_stride + (size_t)c * a.p3p;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            int col = col0 + nt * 8 + 2 * q;
            if (col < a.p3p) {
                double o0 = col < a.p3 ? acc[m][nt][0] : 0.0;
                double o1 = col + 1 < a.p3 ? acc[m][nt][1] : 0.0;
                *reinterpret_cast<double2*>(out + col) = make_double2(o0, o1);
            }
        }
    }
}

// ---- small problems: KR3(X) (N x P3 doubles; 23 MB for the German-shaped data) is formed once when the
// data set is bound and stays L2-resident, so the partials build is a plain TMA-fed DMMA GEMM
// T = Cw . KR3 with no FP64 multiplies competing for the tensor pipe.
// KR2(X)[n, (a,b)] = x_na x_nb, [Np][P2p] row-major: B operand of the plain-GEMM metric build (32 < D)
__global__ void k_form_kr2n(const double* __restrict__ x, const uchar2* __restrict__ pair_tab, double* __restrict__ kr2,
                            long long n_rows_pad, int xs, int p2, int p2p) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_rows_pad * p2p) return;
    long long r = i / p2p;
    int col = (int)(i - r * p2p);
    double v = 0.0;
    if (col < p2) {
        uchar2 ab = pair_tab[col];
        v = x[r * xs + ab.x] * x[r * xs + ab.y];
    }
    kr2[i] = v;
}

__global__ void k_form_kr3(const double* __restrict__ x, const uchar4* __restrict__ tri_tab, double* __restrict__ kr3,
                           long long n_rows_pad, int xs, int p3p) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_rows_pad * p3p) return;
    long long r = i / p3p;
    int col = (int)(i - r * p3p);
    uchar4 t = tri_tab[col];
    const double* xr = x + r * xs;
    kr3[i] = xr[t.x] * xr[t.y] * xr[t.z];
}

__host__ inline size_t tbuild_pre_smem_bytes() {
    return (size_t)kTbStages * ((size_t)kTbChains * kTbAS + (size_t)kTbRows * kTbKS) * 8 + 16 * 8;
}

__global__ void __launch_bounds__(kTbThreads, 1) k_tbuild_pre(TBuildArgs a) {
    constexpr int MC = kTbChains, KB = kTbRows, AS = kTbAS, KS = kTbKS, NW = kTbWarps, ST = kTbStages;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* a_ring = reinterpret_cast<double*>(smem_raw);                 // [ST][MC][AS]  Cw tiles
    double* k_ring = a_ring + (size_t)ST * MC * AS;                       // [ST][KB][KS]  KR3 tiles
    uint64_t* full = reinterpret_cast<uint64_t*>(k_ring + (size_t)ST * KB * KS);
    uint64_t* empty = full + ST;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int chain0 = blockIdx.y * MC;
    if (chain0 >= a.n_chains) return;
    const int n_blocks_all = a.n_rows_pad / KB;
    const int rb_begin = (int)((long long)n_blocks_all * blockIdx.z / gridDim.z);
    const int n_blocks = (int)((long long)n_blocks_all * (blockIdx.z + 1) / gridDim.z) - rb_begin;       // this split's K blocks
    const int colbase = blockIdx.x * kTbCols;
    const int ncols = min(kTbCols, a.p3p - colbase);                     // multiple of 8
    const uint32_t stage_bytes = (uint32_t)(MC * KB * 8 + KB * ncols * 8);

    if (tid == 0) {
        for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
        mbar_fence_init();
    }
    __syncthreads();

    // every warp loads 8 chain rows of the Cw tile and 2 rows of the KR3 tile; thread 0 arms the barrier
    auto issue = [&](int rb) {
        const int stage = rb % ST;
        if (tid == 0) mbar_expect_tx(&full[stage], stage_bytes);
        if (lane < MC / NW) {
            const int m = warp * (MC / NW) + lane;
            tma_bulk_g2s(a_ring + ((size_t)stage * MC + m) * AS,
                         a.cbuf + (size_t)(chain0 + m) * a.n_rows_pad + (size_t)(rb_begin + rb) * KB, (uint32_t)(KB * 8), &full[stage]);
        } else if (lane < MC / NW + KB / NW) {
            const int r = warp * (KB / NW) + (lane - MC / NW);
            tma_bulk_g2s(k_ring + ((size_t)stage * KB + r) * KS,
                         a.kr3 + (size_t)((rb_begin + rb) * KB + r) * a.p3p + colbase, (uint32_t)(ncols * 8), &full[stage]);
        }
    };
    for (int s = 0; s < ST - 1 && s < n_blocks; ++s) issue(s);

    const int wm = warp & 3, wn = warp >> 2;
    double acc[4][4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[m][nt][0] = acc[m][nt][1] = 0.0;

    for (int rb = 0; rb < n_blocks; ++rb) {
        const int stage = rb % ST;
        mbar_wait(&full[stage], (uint32_t)((rb / ST) & 1));
        const double* as = a_ring + (size_t)stage * MC * AS + (size_t)(wm * 32 + g) * AS + q;
        const double* kb = k_ring + (size_t)stage * KB * KS + (size_t)q * KS + wn * 32 + g;
#pragma unroll 4
        for (int ks = 0; ks < KB / 4; ++ks) {
            double af[4], bf[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) af[m] = as[(size_t)m * 8 * AS + ks * 4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) bf[nt] = kb[(size_t)ks * 4 * KS + nt * 8];
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma884(acc[m][nt][0], acc[m][nt][1], af[m], bf[nt]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        const int nb = rb + ST - 1;
        if (nb < n_blocks) {
            if (rb >= 1) mbar_wait(&empty[nb % ST], (uint32_t)(((rb - 1) / ST) & 1));
            issue(nb);
        }
    }

    const int col0 = colbase + wn * 32;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        int c = chain0 + wm * 32 + m * 8 + g;
        if (c >= a.n_chains) continue;
        double* out = gridDim.z > 1 ? a.tpack + blockIdx.z * a.split_stride + (size_t)c * a.p3p
                                    : a.tpack + (size_t)(a.cur[c] ^ a.flip) * a.slot_stride + (size_t)c * a.p3p;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            int col = col0 + nt * 8 + 2 * q;
            if (col < a.p3p) {
                double o0 = col < a.p3 ? acc[m][nt][0] : 0.0;
                double o1 = col + 1 < a.p3 ? acc[m][nt][1] : 0.0;
                *reinterpret_cast<double2*>(out + col) = make_double2(o0, o1);
            }
        }
    }
}
#endif  // __CUDACC__

}  // namespace rmhmc
