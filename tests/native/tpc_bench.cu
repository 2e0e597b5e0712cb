// Stand-alone check + timing of the thread-per-chain Cholesky kernels (csrc/chain_tpc.cuh) against the warp-per-chain
// kernels (csrc/chain_kernels.cuh) on random SPD metrics.  Not part of the library; built and run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tpc_bench tpc_bench.cu && ./tpc_bench [chains] [dim]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../riemannhamiltonianmontecarlo_b200/csrc/chain_tpc.cuh"

using namespace rmhmc;

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e__)); exit(2); } } while (0)

__global__ void k_fill(double* g, double* mom, double* theta, double* u0, int C, int D, int p2p) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    unsigned long long s = 0x9E3779B97F4A7C15ull * (c + 1);
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) * (1.0 / 9007199254740992.0) - 0.5; };
    for (int a = 0; a < D; ++a)
        for (int b = a; b < D; ++b) g[(size_t)c * p2p + pair_index(a, b, D)] = a == b ? 0.6 * D + 4.0 * rnd() : rnd();
    for (int d = 0; d < D; ++d) { mom[(size_t)c * D + d] = 4 * rnd(); theta[(size_t)c * D + d] = rnd(); u0[(size_t)c * D + d] = rnd(); }
}

template <class T> T* dalloc(size_t n) { T* p; CK(cudaMalloc(&p, n * sizeof(T))); CK(cudaMemset(p, 0, n * sizeof(T))); return p; }

static double maxdiff(const double* a, const double* b, size_t n, const char* what) {
    std::vector<double> ha(n), hb(n);
    CK(cudaMemcpy(ha.data(), a, n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), b, n * 8, cudaMemcpyDeviceToHost));
    double worst = 0, scale = 0;
    for (size_t i = 0; i < n; ++i) { worst = fmax(worst, fabs(ha[i] - hb[i])); scale = fmax(scale, fabs(ha[i])); if (ha[i] != ha[i] || hb[i] != hb[i]) worst = 1e300; }
    printf("  %-10s max |diff| %.3e (max |value| %.3e)\n", what, worst, scale);
    return worst / (scale + 1e-300);
}

template <class F> float time_ms(F&& launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

template <int N> int run(int C, int D) {
    const int p2 = num_pairs(D), p2p = pad_up(p2, 8), p2k = pad_up(p2, 32), DD = D * D;
    EngineParams P{};
    P.n_chains = C; P.dim = D; P.ds = D | 1; P.p2 = p2; P.p2p = p2p; P.it_stop = 1 << 30; P.step_size = 0.5; P.alpha = 100;
    P.matrix_free = 1; P.p2k = p2k; P.slot_theta = (size_t)C * D; P.slot_scalar = C; P.slot_invg = (size_t)C * DD;
    ChainArrays S{}, T{};
    S.g_tmp = dalloc<double>((size_t)C * p2p); S.mom = dalloc<double>((size_t)C * D); S.theta = dalloc<double>(2 * (size_t)C * D);
    S.u0 = dalloc<double>((size_t)C * D); S.theta_w = dalloc<double>((size_t)C * D); S.iter = dalloc<long long>(C);
    S.nsteps = dalloc<int>(C); S.cur = dalloc<int>(C); S.step = dalloc<int>(C); S.dir = dalloc<int>(C); S.renorm_pos = dalloc<int>(C);
    S.lfac = dalloc<double>(2 * (size_t)C * DD); S.invg = dalloc<double>(2 * (size_t)C * DD); S.logdet = dalloc<double>(2 * (size_t)C);
    S.qpack = dalloc<double>((size_t)(C + 128) * p2k); S.uvec = dalloc<double>((size_t)C * D); S.aslot = dalloc<int>(C);
    T = S;
    T.theta_w = dalloc<double>((size_t)C * D); T.lfac = dalloc<double>(2 * (size_t)C * DD); T.invg = dalloc<double>(2 * (size_t)C * DD);
    T.logdet = dalloc<double>(2 * (size_t)C); T.qpack = dalloc<double>((size_t)(C + 128) * p2k); T.uvec = dalloc<double>((size_t)C * D);
    T.renorm_pos = dalloc<int>(C); T.aslot = dalloc<int>(C);
    k_fill<<<(C + 127) / 128, 128>>>(S.g_tmp, S.mom, S.theta, S.u0, C, D, p2p);
    std::vector<int> ones(C, 1);
    CK(cudaMemcpy(S.nsteps, ones.data(), C * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(S.dir, ones.data(), C * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_chain_solve_tpc<kTpcTile>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tpc_smem_bytes(D)));
    CK(cudaFuncSetAttribute(k_chain_factor_tpc<kTpcTile>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tpc_smem_bytes(D)));
    CK(cudaDeviceSynchronize());
    int bad = 0;
    const int reps = 20;
    printf("C = %d, D = %d (warp-per-chain order %d), tpc smem %zu bytes per 32 chains\n", C, D, N, tpc_smem_bytes(D));
    float t_old = time_ms([&] { k_chain_solve<N><<<C, 32, solve_smem_bytes(N)>>>(P, S, 1); }, reps);
    float t_new = time_ms([&] { k_chain_solve_tpc<kTpcTile><<<(C + 31) / 32, 32, tpc_smem_bytes(D)>>>(P, T, 1); }, reps);
    printf("solve : warp-per-chain %.4f ms, thread-per-chain %.4f ms (%.2fx)\n", t_old, t_new, t_old / t_new);
    bad += maxdiff(S.theta_w, T.theta_w, (size_t)C * D, "theta_w") > 1e-12;
    float f_old = time_ms([&] { k_chain_factor<N><<<C, 32, factor_smem_bytes(N)>>>(P, S, 0); }, reps);
    float f_new = time_ms([&] { k_chain_factor_tpc<kTpcTile><<<(C + 31) / 32, 32, tpc_smem_bytes(D)>>>(P, T, 0); }, reps);
    printf("factor: warp-per-chain %.4f ms, thread-per-chain %.4f ms (%.2fx)\n", f_old, f_new, f_old / f_new);
    bad += maxdiff(S.lfac + P.slot_invg, T.lfac + P.slot_invg, (size_t)C * DD, "L") > 1e-12;
    bad += maxdiff(S.invg + P.slot_invg, T.invg + P.slot_invg, (size_t)C * DD, "G^-1") > 1e-12;
    bad += maxdiff(S.logdet + P.slot_scalar, T.logdet + P.slot_scalar, C, "logdet") > 1e-12;
    bad += maxdiff(S.qpack, T.qpack, (size_t)C * p2k, "qpack") > 1e-12;
    bad += maxdiff(S.uvec, T.uvec, (size_t)C * D, "uvec") > 1e-12;
    printf(bad ? "MISMATCH\n" : "match\n");
    return bad;
}

int main(int argc, char** argv) {
    const int C = argc > 1 ? atoi(argv[1]) : 65536, D = argc > 2 ? atoi(argv[2]) : 25;
    switch (chain_order(D)) {
        case 8: return run<8>(C, D);
        case 16: return run<16>(C, D);
        case 25: return run<25>(C, D);
        default: return run<32>(C, D);
    }
}
