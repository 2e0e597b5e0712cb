// Blackwell (sm_100a) tensor-core plumbing used by the INT8-slice metric build (i8_metric.cuh):
// tcgen05.mma with TMEM accumulators, tcgen05.ld, tensor-map TMA loads, and the host-side tensor-map encoder
// (cuTensorMapEncodeTiled resolved through cudaGetDriverEntryPoint: the library does not link libcuda).
//
// SASS to look for (profiles/r02/sass_evidence.md): UTCIMMA (tcgen05.mma.kind::i8), LDTM (tcgen05.ld),
// UTMALDG (cp.async.bulk.tensor), UTCBAR (tcgen05.commit).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"

namespace rmhmc {

// ---------------------------------------------------------------- host: tensor maps
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tensorMapEncodeTiled tensor_map_encoder() {
    static PFN_tensorMapEncodeTiled fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_tensorMapEncodeTiled>(p);
    return fn;
}

// 2-D byte matrix [rows][row_bytes] (row-major, K contiguous), box = box_rows x 64 bytes, SWIZZLE_64B: lands in shared
// memory as the K-major canonical UMMA layout (8-row x 64-byte swizzle atoms, 512 bytes between atoms)
inline bool make_tensor_map_u8_k64(CUtensorMap* map, const void* base, uint64_t rows, uint64_t row_bytes, uint32_t box_rows) {
    PFN_tensorMapEncodeTiled enc = tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t dims[2] = {row_bytes, rows};
    cuuint64_t strides[1] = {row_bytes};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------- device: TMA tensor load
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

// mbarrier wait that turns a dead pipeline (bad descriptor, lost TMA) into a trap instead of a hung GPU:
// ~2 s of polling at 2 GHz
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

// ---------------------------------------------------------------- device: TMEM
// one full warp executes alloc / dealloc (.sync.aligned); the allocated base address is written to shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t n_cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(n_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_addr, uint32_t n_cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "r"(n_cols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 16 consecutive 32-bit columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- device: UMMA descriptors and issue
// K-major operand tile in shared memory, SWIZZLE_64B: rows of 64 bytes, 8-row atoms 512 bytes apart (SBO), the
// 16-byte chunks of a row XOR-ed with (row >> 1) & 3 by TMA on the way in and by the tensor core on the way out.
// Field layout (cute/arch/mma_sm100_desc.hpp, SmemDescriptor): start >> 4 in [0,14), LBO >> 4 in [16,30),
// SBO >> 4 in [32,46), version = 1 in [46,48), layout type in [61,64) (4 = SWIZZLE_64B).
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                  // leading byte offset: unused for swizzled K-major layouts
    d |= (uint64_t)(512u >> 4) << 32;        // stride byte offset between 8-row atoms
    d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
    d |= (uint64_t)4 << 61;                  // SWIZZLE_64B
    return d;
}
// Instruction descriptor of tcgen05.mma.kind::i8 (InstrDescriptor in the same header): S32 accumulate (c_format = 2,
// bits [4,6)), signed 8-bit A and B (a_format = b_format = 1, bits [7,10) / [10,13)), both K-major, N >> 3 in [17,23),
// M >> 4 in [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_s8(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] . B[smem]^T, issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_i8_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all tcgen05.mma issued so far by this thread arrive on `bar` when they have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#endif  // __CUDACC__

}  // namespace rmhmc
