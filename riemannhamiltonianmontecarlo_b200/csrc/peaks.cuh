// Live roofline denominators (blr_device_peaks): the issue peaks of the two tensor paths this library uses, measured
// on the device the benchmark runs on.  MEASURED_PEAKS.json (driver-written) holds only HBM and bf16 figures.
//   FP64  DMMA.8x8x4 (mma.sync m8n8k4 f64), 8 independent accumulator pairs per warp, 16 warps per SM
//   INT8  tcgen05.mma.kind::i8 128 x 256 x 32 from a resident shared-memory tile into TMEM, one CTA per SM
// Both kernels do nothing but issue MMAs, so the figures are upper bounds for any kernel on that path.
#pragma once
#include "umma_common.cuh"

namespace rmhmc {

#ifdef __CUDACC__
constexpr int kPeakIters = 2048;

__global__ void __launch_bounds__(512) k_peak_dmma(double* out, double a, double b) {
    constexpr int ILP = 8;
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x; c1[i] = i; }
    for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;          // keeps the loop alive
}

// one CTA per SM: 128 threads; thread 0 issues kPeakIters MMAs of 128 x 256 x 32 on a 4 KB + 8 KB operand tile
__global__ void __launch_bounds__(128, 1) k_peak_i8(int* out) {
    extern __shared__ unsigned char peak_smem[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(peak_smem) + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + 128 * 64 + 256 * 64);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < (128 * 64 + 256 * 64) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x01010101u;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(slot, 256);
    // shared-memory writes above must be visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        const uint64_t da = umma_desc_k_sw64(smem_u32(base)), db = umma_desc_k_sw64(smem_u32(base + 128 * 64));
        const uint32_t idesc = umma_idesc_s8(128, 256);
        for (int it = 0; it < kPeakIters; ++it) umma_i8_ss(tmem, da, db, idesc, it ? 1u : 0u);
        umma_commit(bar);
        mbar_wait_or_trap(bar, 0);
        tcgen05_fence_after();
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t r[16];
        tmem_ld16(tmem, r);
        tmem_ld_wait();
        if (r[0] == 0x7fffffffu) out[0] = (int)r[1];
        tcgen05_fence_before();
    }
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

// returns cudaSuccess and fills TFLOP/s (FP64 DMMA) and TOP/s (INT8 tcgen05); best of 3 timed launches each
inline cudaError_t measure_device_peaks(cudaStream_t st, double* dmma_tflops, double* i8_tops) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    double* dout = nullptr;
    e = cudaMalloc((void**)&dout, 64);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t smem_i8 = 128 * 64 + 256 * 64 + 1024 + 64;
    e = cudaFuncSetAttribute(k_peak_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_i8);
    double best_d = 0.0, best_i = 0.0;
    for (int rep = 0; rep < 4 && e == cudaSuccess; ++rep) {
        float ms = 0.f;
        cudaEventRecord(e0, st);
        k_peak_dmma<<<sms * 2, 512, 0, st>>>(dout, 1.0000001, 1e-9);
        cudaEventRecord(e1, st);
        e = cudaEventSynchronize(e1);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) break;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * 8 * 4 * 8 /*ILP*/ * kPeakIters * (512 / 32) * (double)(sms * 2);
        if (rep > 0 && ms > 0) best_d = std::fmax(best_d, flops / (ms * 1e-3) / 1e12);
        cudaEventRecord(e0, st);
        k_peak_i8<<<sms, 128, smem_i8, st>>>(reinterpret_cast<int*>(dout));
        cudaEventRecord(e1, st);
        e = cudaEventSynchronize(e1);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) break;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = 2.0 * 128 * 256 * 32 * kPeakIters * (double)sms;
        if (rep > 0 && ms > 0) best_i = std::fmax(best_i, ops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dout);
    if (dmma_tflops) *dmma_tflops = best_d;
    if (i8_tops) *i8_tops = best_i;
    return e;
}
#endif

}  // namespace rmhmc
