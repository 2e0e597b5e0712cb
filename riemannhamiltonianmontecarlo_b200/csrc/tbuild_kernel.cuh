// Metric-partials build: T[chain, (i,j,k)] = sum_n c_n x_ni x_nj x_nk, c_n = v_n (1 - 2 p_n).
//
// The reference forms the D matrices dG/dw_d = X^T diag(c x_d) X one by one (rmhmc.py:64-75,
// :142-153).  Stacked they are a fully symmetric 3-tensor, so only the P3 = D(D+1)(D+2)/6 packed
// entries i<=j<=k are computed (4-5x fewer flops), as ONE chain-batched contraction on the FP64
// tensor cores (DMMA.8x8x4):
//     T[chains x P3] = Cw[chains x rows] . KR3(X)[rows x P3],  KR3(X)[n,(i,j,k)] = x_ni x_nj x_nk
// KR3 is never materialised: each lane forms its B-fragment element from three staged X values.
// Cw tiles (written by the closing metric build) and X row blocks both stream through 1-D bulk TMA
// into a 3-stage mbarrier ring filled by a dedicated producer warp.  This kernel carries ~60% of the
// algorithmic flops.
#pragma once
#include "common.cuh"

namespace rmhmc {

constexpr int kTbChains = 128;    // chains per CTA
constexpr int kTbCols = 128;      // packed-triple columns per CTA
constexpr int kTbRows = 32;       // rows per staged block (K tile)
constexpr int kTbStages = 3;
constexpr int kTbAS = kTbRows + 4;  // smem stride of the Cw tile (4*odd)

struct TBuildArgs {
    const double* x;          // [Np][XS]
    const uchar4* tri_tab;    // [P3p] packed column -> (i, j, k)
    const double* cbuf;       // [Cpad][Np]
    double* tpack;            // [2][C][P3p]
    const int* cur;           // [C] current slot per chain
    int flip;                 // write to slot cur ^ flip
    size_t slot_stride;       // doubles between the two T slots
    int n_chains, n_rows_pad, xs, p3, p3p;
};

__host__ inline size_t tbuild_smem_bytes(int xs) {
    return (size_t)kTbStages * ((size_t)kTbChains * kTbAS + (size_t)kTbRows * xs) * 8 + 2 * kTbStages * 8;
}

#ifdef __CUDACC__
// 16 warps, 4 x 4 over the 128 x 128 tile (32 chains x 32 columns each).  The 3-stage ring is filled
// with bulk TMA only -- 128 row segments of Cw (256 B each, landing in the padded smem rows; every warp
// issues 8 of them) and one contiguous X row block per stage -- all completing on the stage's `full`
// mbarrier; a stage is refilled once all 16 warps have released it through its `empty` mbarrier.  There
// is no CTA-wide barrier in the main loop, so warps drift apart and their fragment loads overlap the
// other warps' DMMAs.
constexpr int kTbWarps = 16;
constexpr int kTbThreads = kTbWarps * 32;

__global__ void __launch_bounds__(kTbThreads, 1) k_tbuild(TBuildArgs a) {
    constexpr int MC = kTbChains, KB = kTbRows, AS = kTbAS, ST = kTbStages;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int xs = a.xs;
    double* a_ring = reinterpret_cast<double*>(smem_raw);                 // [ST][MC][AS]
    double* x_ring = a_ring + (size_t)ST * MC * AS;                       // [ST][KB][xs]
    uint64_t* full = reinterpret_cast<uint64_t*>(x_ring + (size_t)ST * KB * xs);
    uint64_t* empty = full + ST;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int chain0 = blockIdx.y * MC;
    if (chain0 >= a.n_chains) return;
    const int n_blocks = a.n_rows_pad / KB;
    const uint32_t x_bytes = (uint32_t)(KB * xs * 8);
    const uint32_t stage_bytes = x_bytes + (uint32_t)(MC * KB * 8);

    if (tid == 0) {
        for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kTbWarps); }
        mbar_fence_init();
    }
    __syncthreads();

    // every warp loads 8 chain rows of the Cw tile; warp 0 also arms the barrier and loads the X block
    auto issue = [&](int rb) {
        const int stage = rb % ST;
        if (tid == 0) {
            mbar_expect_tx(&full[stage], stage_bytes);
            tma_bulk_g2s(x_ring + (size_t)stage * KB * xs, a.x + (size_t)rb * KB * xs, x_bytes, &full[stage]);
        }
        if (lane < MC / kTbWarps) {
            const int m = warp * (MC / kTbWarps) + lane;
            tma_bulk_g2s(a_ring + ((size_t)stage * MC + m) * AS,
                         a.cbuf + (size_t)(chain0 + m) * a.n_rows_pad + (size_t)rb * KB, (uint32_t)(KB * 8), &full[stage]);
        }
    };
    for (int s = 0; s < ST - 1 && s < n_blocks; ++s) issue(s);

    const int wm = warp & 3, wn = warp >> 2;
    const int col0 = blockIdx.x * kTbCols + wn * 32;
    int ti[4], tj[4], tk[4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        int col = col0 + nt * 8 + g;
        uchar4 t = col < a.p3p ? a.tri_tab[col] : make_uchar4(0, 0, 0, 0);
        ti[nt] = t.x; tj[nt] = t.y; tk[nt] = t.z;
    }
    double acc[4][4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc[m][nt][0] = acc[m][nt][1] = 0.0;

    for (int rb = 0; rb < n_blocks; ++rb) {
        const int stage = rb % ST;
        mbar_wait(&full[stage], (uint32_t)((rb / ST) & 1));
        const double* as = a_ring + (size_t)stage * MC * AS + (size_t)(wm * 32 + g) * AS + q;
        const double* xb = x_ring + (size_t)stage * KB * xs;
#pragma unroll 2
        for (int ks = 0; ks < KB / 4; ++ks) {
            const double* xr = xb + (size_t)(ks * 4 + q) * xs;
            double af[4], bf[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) af[m] = as[(size_t)m * 8 * AS + ks * 4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) bf[nt] = xr[ti[nt]] * xr[tj[nt]] * xr[tk[nt]];
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma884(acc[m][nt][0], acc[m][nt][1], af[m], bf[nt]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        // refill the stage of block rb-1 (released by everybody by now, or very soon) with block rb+ST-1
        const int nb = rb + ST - 1;
        if (nb < n_blocks) {
            if (rb >= 1) mbar_wait(&empty[nb % ST], (uint32_t)(((rb - 1) / ST) & 1));
            issue(nb);
        }
    }

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
#endif  // __CUDACC__

}  // namespace rmhmc
