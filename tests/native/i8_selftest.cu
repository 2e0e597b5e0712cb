// Stand-alone bring-up / regression test of the INT8-slice metric build (csrc/i8_metric.cuh) on a B200:
//   1. raw digit GEMM: every TMEM accumulator class and the int64 recombination against exact CPU integer arithmetic
//      (ragged chain tile, partial last column chunk, more K blocks than pipeline stages);
//   2. end to end: k_i8_colmax / k_i8_form_b / k_i8_vslice / k_i8_gemm against an FP64 CPU evaluation of
//      G = X^T diag(v) X + I/alpha, X^T (t - p), the log-likelihood and c_n on German-shaped synthetic data;
//   3. timing of the two kernels at 65 536 chains (CUDA events).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tests/native/i8_selftest tests/native/i8_selftest.cu
// Exit code 0 = all checks passed.  Test infrastructure: nothing here is shipped.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../riemannhamiltonianmontecarlo_b200/csrc/i8_metric.cuh"

using namespace rmhmc;

#define CK(expr)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            std::printf("CUDA error %s at %s:%d: %s\n", #expr, __FILE__, __LINE__, cudaGetErrorString(e__)); \
            std::exit(2);                                                                             \
        }                                                                                             \
    } while (0)

template <int S>
static int test_raw_gemm() {
    constexpr int NC = I8Shape<S>::NC;
    const int kp = 6 * kI8BlockK, n_chains = 200, a_rows = 256, p2 = NC + NC / 2 + 6, p2p = pad_up(p2, 8), b_rows = 2 * NC;
    std::mt19937 rng(123);
    std::uniform_int_distribution<int> dig(-128, 127);
    std::vector<signed char> ha((size_t)S * a_rows * kp), hb((size_t)S * b_rows * kp);
    for (auto& v : ha) v = (signed char)dig(rng);
    for (auto& v : hb) v = (signed char)dig(rng);
    std::vector<double2> hci(b_rows);
    for (int c = 0; c < b_rows; ++c) hci[c] = make_double2(c < p2 ? std::ldexp(1.0, -20) : 0.0, 0.0);
    signed char *da, *db;
    double2* dci;
    double* dg;
    CK(cudaMalloc(&da, ha.size())); CK(cudaMalloc(&db, hb.size()));
    CK(cudaMalloc(&dci, hci.size() * sizeof(double2)));
    CK(cudaMalloc(&dg, (size_t)n_chains * p2p * 8));
    CK(cudaMemcpy(da, ha.data(), ha.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dci, hci.data(), hci.size() * sizeof(double2), cudaMemcpyHostToDevice));
    CUtensorMap ma, mb;
    if (!make_tensor_map_u8_k64(&ma, da, (uint64_t)S * a_rows, kp, kI8TileM) || !make_tensor_map_u8_k64(&mb, db, (uint64_t)S * b_rows, kp, NC)) {
        std::printf("tensor map encode failed\n");
        return 1;
    }
    // exact class sums on the CPU
    std::vector<long long> cls((size_t)S * n_chains * p2, 0);
    for (int w = 0; w < S; ++w)
        for (int i = 0; i <= w; ++i)
            for (int c = 0; c < n_chains; ++c)
                for (int col = 0; col < p2; ++col) {
                    long long s = 0;
                    const signed char* ar = &ha[((size_t)i * a_rows + c) * kp];
                    const signed char* br = &hb[((size_t)(w - i) * b_rows + col) * kp];
                    for (int k = 0; k < kp; ++k) s += (int)ar[k] * (int)br[k];
                    cls[((size_t)w * n_chains + c) * p2 + col] += s;
                }
    int bad_total = 0;
    std::vector<double> hg((size_t)n_chains * p2p);
    for (int dbg = -1; dbg < S; ++dbg) {
        CK(cudaMemset(dg, 0xff, hg.size() * 8));
        I8GemmArgs a{};
        a.g_out = dg; a.colinfo = dci; a.alpha_inv = 0.0; a.n_chains = n_chains; a.p2 = p2; a.p2p = p2p;
        a.k_blocks = kp / kI8BlockK; a.a_rows = a_rows; a.b_rows = b_rows; a.debug_class = dbg;
        CK(i8_launch_gemm<S>(ma, mb, a, 0));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hg.data(), dg, hg.size() * 8, cudaMemcpyDeviceToHost));
        int bad = 0;
        double worst = 0.0;
        for (int c = 0; c < n_chains; ++c)
            for (int col = 0; col < p2p; ++col) {
                double want = 0.0;
                if (col < p2) {
                    long long t;
                    if (dbg >= 0) t = cls[((size_t)dbg * n_chains + c) * p2 + col];
                    else {
                        t = cls[((size_t)0 * n_chains + c) * p2 + col];
                        for (int w = 1; w < S; ++w) t = t * 256 + cls[((size_t)w * n_chains + c) * p2 + col];
                    }
                    want = (double)t * std::ldexp(1.0, -20);
                    if (S > 5 && dbg < 0) {          // the kernel joins two exact halves with one FMA
                        long long th = cls[((size_t)0 * n_chains + c) * p2 + col], tl = 0;
                        for (int w = 1; w < 3; ++w) th = th * 256 + cls[((size_t)w * n_chains + c) * p2 + col];
                        for (int w = 3; w < S; ++w) tl = tl * 256 + cls[((size_t)w * n_chains + c) * p2 + col];
                        want = std::fma((double)th, 16777216.0, (double)tl) * std::ldexp(1.0, -20);
                    }
                }
                const double got = hg[(size_t)c * p2p + col];
                if (!(got == want)) {
                    if (bad < 6) std::printf("  raw S=%d class %d: chain %d col %d got %.17g want %.17g\n", S, dbg, c, col, got, want);
                    ++bad;
                    worst = std::fmax(worst, std::fabs(got - want));
                }
            }
        std::printf("raw gemm S=%d class %2d: %d mismatches of %d (worst |diff| %.3g)\n", S, dbg, bad, n_chains * p2p, worst);
        bad_total += bad;
    }
    cudaFree(da); cudaFree(db); cudaFree(dci); cudaFree(dg);
    return bad_total;
}

struct Problem {
    int n_rows, dim, xs, n_rows_pad, p2, p2p;
    std::vector<double> x_pad;         // [Np][xs], label in column xs-1
    std::vector<uchar2> pair_tab;
};
static Problem make_problem(int n_rows, int dim, unsigned seed) {
    Problem p;
    p.n_rows = n_rows; p.dim = dim; p.xs = x_stride(dim); p.n_rows_pad = pad_up(n_rows, 32);
    p.p2 = num_pairs(dim); p.p2p = pad_up(p.p2, 8);
    p.x_pad.assign((size_t)p.n_rows_pad * p.xs, 0.0);
    std::mt19937_64 rng(seed);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::uniform_real_distribution<double> ud(0.0, 1.0);
    std::vector<double> beta(dim);
    for (auto& b : beta) b = 0.5 * nd(rng);
    for (int n = 0; n < n_rows; ++n) {
        double* r = &p.x_pad[(size_t)n * p.xs];
        r[0] = 1.0;
        double prev = 0.0, f = beta[0];
        for (int d = 1; d < dim; ++d) { prev = 0.3 * prev + std::sqrt(0.91) * nd(rng); r[d] = prev; f += prev * beta[d]; }
        r[p.xs - 1] = ud(rng) < 1.0 / (1.0 + std::exp(-f)) ? 1.0 : 0.0;
    }
    p.pair_tab.assign(p.p2p, make_uchar2(0, 0));
    for (int a = 0; a < dim; ++a)
        for (int b = a; b < dim; ++b) p.pair_tab[pair_index(a, b, dim)] = make_uchar2((unsigned char)a, (unsigned char)b);
    return p;
}

template <int S>
static int test_end_to_end(int n_rows, int dim, int n_chains, double theta_sd, bool time_it) {
    constexpr int NC = I8Shape<S>::NC;
    Problem P = make_problem(n_rows, dim, 99 + dim);
    const int kp = i8_kp(P.n_rows_pad), a_rows = pad_up(n_chains, kI8TileM), chunks = i8_chunks<S>(P.p2), b_rows = chunks * NC;
    const double alpha = 100.0;
    std::mt19937_64 rng(7);
    std::normal_distribution<double> nd(0.0, theta_sd);
    std::vector<double> theta((size_t)n_chains * dim);
    for (auto& t : theta) t = nd(rng);
    for (int d = 0; d < dim; ++d) theta[d] = 0.0;          // chain 0: f = 0 everywhere, v = 1/4 exactly (largest digit value)
    double *dx, *dth, *dcolmax, *dg, *dgrad, *dll, *dcb;
    uchar2* dpt;
    signed char *da, *db;
    double2* dci;
    CK(cudaMalloc(&dx, P.x_pad.size() * 8)); CK(cudaMalloc(&dth, theta.size() * 8)); CK(cudaMalloc(&dcolmax, P.p2p * 8));
    CK(cudaMalloc(&dg, (size_t)n_chains * P.p2p * 8)); CK(cudaMalloc(&dgrad, (size_t)n_chains * dim * 8)); CK(cudaMalloc(&dll, (size_t)n_chains * 8));
    CK(cudaMalloc(&dcb, (size_t)a_rows * P.n_rows_pad * 8));
    CK(cudaMalloc(&dpt, P.pair_tab.size() * sizeof(uchar2)));
    CK(cudaMalloc(&da, (size_t)S * a_rows * kp)); CK(cudaMalloc(&db, (size_t)S * b_rows * kp)); CK(cudaMalloc(&dci, b_rows * sizeof(double2)));
    CK(cudaMemset(da, 0, (size_t)S * a_rows * kp));
    CK(cudaMemcpy(dx, P.x_pad.data(), P.x_pad.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dth, theta.data(), theta.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dpt, P.pair_tab.data(), P.pair_tab.size() * sizeof(uchar2), cudaMemcpyHostToDevice));
    k_i8_colmax<<<P.p2, 256>>>(dx, dpt, dcolmax, P.n_rows_pad, P.xs);
    {
        long long n = (long long)b_rows * kp;
        k_i8_form_b<S><<<(unsigned)((n + 255) / 256), 256>>>(dx, dpt, dcolmax, db, dci, P.n_rows_pad, P.xs, P.p2, b_rows, kp);
    }
    CK(cudaDeviceSynchronize());
    CUtensorMap ma, mb;
    if (!make_tensor_map_u8_k64(&ma, da, (uint64_t)S * a_rows, kp, kI8TileM) || !make_tensor_map_u8_k64(&mb, db, (uint64_t)S * b_rows, kp, NC)) {
        std::printf("tensor map encode failed\n");
        return 1;
    }
    I8VsArgs v{};
    v.x = dx; v.theta = dth; v.a8 = da; v.plane_stride = (size_t)a_rows * kp; v.kp = kp;
    v.n_chains = n_chains; v.n_rows = P.n_rows; v.n_rows_pad = P.n_rows_pad; v.dim = dim; v.xs = P.xs;
    v.grad_out = dgrad; v.loglik_out = dll; v.cbuf = dcb; v.cw_cur = nullptr; v.cw_flip = 0; v.cw_slot = 0;
    I8GemmArgs g{};
    g.g_out = dg; g.colinfo = dci; g.alpha_inv = 1.0 / alpha; g.n_chains = n_chains; g.p2 = P.p2; g.p2p = P.p2p;
    g.k_blocks = kp / kI8BlockK; g.a_rows = a_rows; g.b_rows = b_rows; g.debug_class = -1;
    auto set_attr = [&]() {
        const int smem = (int)i8_vslice_smem(P.xs);
        (void)smem;       // < 48 KB: no attribute needed
    };
    set_attr();

    int bad = 0;
    const int n_check = n_chains < 24 ? n_chains : 24;
    for (int variant = 0; variant < 4; ++variant) {       // 0 scalar iterate kernel, 1 scalar closing kernel, 2 DMMA iterate kernel, 3 DMMA closing kernel
        const int closing = variant == 3 ? 1 : variant;
        if (variant == 3) { CK(cudaMemset(dgrad, 0xff, (size_t)n_chains * dim * 8)); CK(cudaMemset(dll, 0xff, (size_t)n_chains * 8)); CK(cudaMemset(dcb, 0xff, (size_t)a_rows * P.n_rows_pad * 8)); }
        CK(cudaMemset(dg, 0xff, (size_t)n_chains * P.p2p * 8));
        CK(cudaMemset(da, 0x55, (size_t)S * a_rows * kp));
        if (variant == 1) CK((i8_launch_vslice<S, true>(v, 0)));
        else if (variant == 0) CK((i8_launch_vslice<S, false>(v, 0)));
        else if (variant == 2) CK((i8_launch_vslice_mma<S>(v, 0)));
        else CK((i8_launch_vslice_mma_closing<S>(v, 0)));
        CK(i8_launch_gemm<S>(ma, mb, g, 0));
        CK(cudaDeviceSynchronize());
        std::vector<double> hg((size_t)n_chains * P.p2p), hgrad((size_t)n_chains * dim), hll(n_chains), hcb;
        CK(cudaMemcpy(hg.data(), dg, hg.size() * 8, cudaMemcpyDeviceToHost));
        if (closing == 1) {
            hcb.resize((size_t)a_rows * P.n_rows_pad);
            CK(cudaMemcpy(hgrad.data(), dgrad, hgrad.size() * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hll.data(), dll, hll.size() * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hcb.data(), dcb, hcb.size() * 8, cudaMemcpyDeviceToHost));
        }
        double eg = 0, egr = 0, ell = 0, ecb = 0;
        for (int ci = 0; ci < n_check; ++ci) {
            const int c = ci < n_check / 2 ? ci : n_chains - 1 - (ci - n_check / 2);       // first and last chains (ragged tile)
            std::vector<double> G(P.p2, 0.0), gr(dim, 0.0);
            double ll = 0.0, gmax = 0.0, cbmax = 0.0, cberr = 0.0;
            for (int n = 0; n < P.n_rows; ++n) {
                const double* r = &P.x_pad[(size_t)n * P.xs];
                double f = 0.0;
                for (int d = 0; d < dim; ++d) f += r[d] * theta[(size_t)c * dim + d];
                const double p = 1.0 / (1.0 + std::exp(-f)), vv = p * (1.0 - p), t = r[P.xs - 1];
                for (int a = 0; a < dim; ++a)
                    for (int b = a; b < dim; ++b) G[pair_index(a, b, dim)] += vv * r[a] * r[b];
                for (int d = 0; d < dim; ++d) gr[d] += (t - p) * r[d];
                ll += t * f - (std::fmax(f, 0.0) + std::log1p(std::exp(-std::fabs(f))));
                if (closing == 1) {
                    const double cw = vv * (1.0 - 2.0 * p);
                    cbmax = std::fmax(cbmax, std::fabs(cw));
                    cberr = std::fmax(cberr, std::fabs(cw - hcb[(size_t)c * P.n_rows_pad + n]));
                }
            }
            for (int a = 0; a < dim; ++a) G[pair_index(a, a, dim)] += 1.0 / alpha;
            double err = 0.0;
            for (int k = 0; k < P.p2; ++k) { gmax = std::fmax(gmax, std::fabs(G[k])); err = std::fmax(err, std::fabs(G[k] - hg[(size_t)c * P.p2p + k])); }
            eg = std::fmax(eg, err / gmax);
            if (closing == 1) {
                double m = 0, e = 0;
                for (int d = 0; d < dim; ++d) { m = std::fmax(m, std::fabs(gr[d])); e = std::fmax(e, std::fabs(gr[d] - hgrad[(size_t)c * dim + d])); }
                egr = std::fmax(egr, e / m);
                ell = std::fmax(ell, std::fabs(ll - hll[c]) / std::fabs(ll));
                ecb = std::fmax(ecb, cberr / std::fmax(cbmax, 1e-300));
            }
        }
        const double tol_g = S == 5 ? 2e-11 : 1e-12;
        std::printf("end-to-end S=%d N=%d D=%d C=%d theta_sd=%.2f %s: max rel err G %.3g (tol %.1g)", S, n_rows, dim, n_chains, theta_sd,
                    variant == 1 ? "closing" : (variant == 0 ? "iterate" : (variant == 2 ? "iterate(dmma)" : "closing(dmma)")), eg, tol_g);
        if (closing == 1) std::printf(", grad %.3g, loglik %.3g, c_n %.3g", egr, ell, ecb);
        std::printf("\n");
        if (!(eg < tol_g)) ++bad;
        if (closing == 1 && !(egr < 1e-12 && ell < 1e-12 && ecb < 1e-12)) ++bad;
    }
    if (time_it) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        auto timeit = [&](const char* name, auto&& fn, double ops) {
            for (int i = 0; i < 3; ++i) fn();
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            const int reps = 20;
            for (int i = 0; i < reps; ++i) fn();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            ms /= reps;
            std::printf("time S=%d C=%d %-16s %.4f ms  (%.1f T%s/s)\n", S, n_chains, name, ms, ops / (ms * 1e-3) / 1e12, name[0] == 'g' ? "OP" : "FLOP");
        };
        const double el = (double)n_chains * P.n_rows_pad;
        timeit("vslice", [&]() { CK((i8_launch_vslice<S, false>(v, 0))); }, el * 2.0 * (dim + 20));
        timeit("vslice_mma", [&]() { CK((i8_launch_vslice_mma<S>(v, 0))); }, el * 2.0 * (dim + 20));
        timeit("vslice_closing", [&]() { CK((i8_launch_vslice<S, true>(v, 0))); }, el * 2.0 * (2 * dim + 32));
        timeit("vslice_mma_closing", [&]() { CK((i8_launch_vslice_mma_closing<S>(v, 0))); }, el * 2.0 * (2 * dim + 32));
        timeit("gemm", [&]() { CK(i8_launch_gemm<S>(ma, mb, g, 0)); }, 2.0 * (S * (S + 1) / 2) * (double)a_rows * kp * b_rows);
        timeit("vslice_mma+gemm", [&]() { CK((i8_launch_vslice_mma<S>(v, 0))); CK(i8_launch_gemm<S>(ma, mb, g, 0)); }, 2.0 * (double)n_chains * P.n_rows * (P.p2 + dim));
    }
    cudaFree(dx); cudaFree(dth); cudaFree(dcolmax); cudaFree(dg); cudaFree(dgrad); cudaFree(dll); cudaFree(dcb); cudaFree(dpt);
    cudaFree(da); cudaFree(db); cudaFree(dci);
    return bad;
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? std::atoi(argv[1]) : 0;      // 0 checks, 1 checks + timing, 2 only the 65 536-chain case (profiling)
    const bool timing = mode == 1;
    int bad = 0;
    if (mode == 2) {
        bad += test_end_to_end<5>(1000, 25, 65536, 0.3, true);
        std::printf(bad ? "I8 SELFTEST FAILED (%d)\n" : "I8 SELFTEST OK\n", bad);
        return bad ? 1 : 0;
    }
    bad += test_raw_gemm<5>();
    bad += test_raw_gemm<6>();
    bad += test_end_to_end<5>(1000, 25, 300, 0.3, false);
    bad += test_end_to_end<5>(690, 15, 130, 2.0, false);
    bad += test_end_to_end<5>(532, 8, 37, 0.5, false);
    bad += test_end_to_end<6>(1000, 25, 300, 0.3, false);
    bad += test_end_to_end<5>(270, 32, 129, 0.2, false);
    if (timing) {
        bad += test_end_to_end<5>(1000, 25, 65536, 0.3, true);
        bad += test_end_to_end<6>(1000, 25, 65536, 0.3, true);
        bad += test_end_to_end<5>(690, 15, 4096, 0.3, true);
        bad += test_end_to_end<5>(1000, 25, 8192, 0.3, true);
        bad += test_end_to_end<5>(1000, 25, 16384, 0.3, true);
        bad += test_end_to_end<5>(1000, 25, 32768, 0.3, true);
    }
    std::printf(bad ? "I8 SELFTEST FAILED (%d)\n" : "I8 SELFTEST OK\n", bad);
    return bad ? 1 : 0;
}
