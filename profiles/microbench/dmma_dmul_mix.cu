// Microbenchmark: cost of interleaving DMUL / DFMA with DMMA.8x8x4 on the FP64 pipe (B200, sm_100a).
// Each warp runs ITERS x { 16 DMMA (independent accumulators) + NMUL DMUL feeding the B operands }.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_dmul_mix dmma_dmul_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;
template <int NMUL, bool DEP>
__global__ void k(double* out, double a, double b) {
    double c0[16], c1[16], m[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c0[i] = threadIdx.x; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = b + i * 1e-3 + threadIdx.x * 1e-6;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NMUL; ++i) m[i] = m[i] * 1.0000001;          // DMUL on the FP64 pipe
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            double bb = DEP ? m[i & 7] : b;
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(bb));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += m[i];
    if (s == 12345.678) out[0] = s;
}
template <typename K>
void run(const char* name, K kern, int warps, int nmul, double* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) kern<<<148, warps * 32>>>(d, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); kern<<<148, warps * 32>>>(d, 1.0000001, 1e-9); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double cyc = best * 1e-3 * 1.965e9;                      // per SM
    double per_iter = cyc / ITERS / (warps / 4.0);           // cycles per (16 DMMA + nmul DMUL) per warp-slot on one SMSP
    printf("%-26s warps/SM=%2d  %7.3f ms  cycles per 16-DMMA group per SMSP-warp = %7.1f (ideal %d + %d)\n", name, warps, best, per_iter, 256, 2 * nmul);
}
int main() {
    double* d; cudaMalloc(&d, 64);
    for (int w : {4, 8, 16}) {
        run("16 DMMA", k<0, false>, w, 0, d);
        run("16 DMMA + 4 DMUL (indep)", k<4, false>, w, 4, d);
        run("16 DMMA + 8 DMUL (indep)", k<8, false>, w, 8, d);
        run("16 DMMA + 4 DMUL (dep B)", k<4, true>, w, 4, d);
        run("16 DMMA + 8 DMUL (dep B)", k<8, true>, w, 8, d);
    }
    return 0;
}
