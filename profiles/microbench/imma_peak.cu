// Microbenchmark: legacy INT8 tensor-core issue peak on B200 (sm_100a): mma.sync.m16n8k32.s8 (SASS IMMA.16832.S8.S8)
// and, for comparison, bf16 mma.sync.m16n8k16.  Input for DESIGN.md section 7 (INT8-slice emulation of the FP64 metric
// build): is the legacy warp-level path fast enough, or does it have to be tcgen05.mma.kind::i8?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_peak imma_peak.cu
// Run:   ./imma_peak     (prints TOP/s resp. TFLOP/s per variant; CUDA-event timed, 3 warm-ups)
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

template <int ILP>
__global__ void k_imma(int* out, unsigned a, unsigned b) {
    int c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                         : "r"(a), "r"(b), "r"(a), "r"(b), "r"(b), "r"(a));
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123456789) out[0] = s;
}

template <int ILP>
__global__ void k_hmma_bf16(float* out, unsigned a, unsigned b) {
    float c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a), "r"(b), "r"(a), "r"(b), "r"(b), "r"(a));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678f) out[0] = s;
}

template <typename K>
double time_ms(K launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int* oi; float* of;
    CK(cudaMalloc(&oi, 64)); CK(cudaMalloc(&of, 64));
    printf("%s, %d SMs\n", prop.name, sms);
    for (int warps : {4, 8, 16}) {
        const int blocks = sms * 2, threads = warps * 32;
        const double total_warps = (double)blocks * warps;
        {
            double ms = time_ms([&] { k_imma<8><<<blocks, threads>>>(oi, 0x01020304u, 0x05060708u); });
            double ops = total_warps * ITERS * 8 * (2.0 * 16 * 8 * 32);
            printf("mma.sync m16n8k32 s8   ILP 8, %2d warps/CTA, 2 CTAs/SM: %8.1f TOP/s\n", warps, ops / (ms * 1e-3) / 1e12);
        }
        {
            double ms = time_ms([&] { k_hmma_bf16<8><<<blocks, threads>>>(of, 0x3f803f80u, 0x3f803f80u); });
            double ops = total_warps * ITERS * 8 * (2.0 * 16 * 8 * 16);
            printf("mma.sync m16n8k16 bf16 ILP 8, %2d warps/CTA, 2 CTAs/SM: %8.1f TFLOP/s\n", warps, ops / (ms * 1e-3) / 1e12);
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
