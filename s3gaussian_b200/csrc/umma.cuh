// umma.cuh - minimal hand-written tcgen05 (5th-gen tensor core) building blocks for sm_100a:
// TMEM allocation, shared-memory matrix descriptors for the un-swizzled K-major canonical
// layout, the kind::tf32 MMA issue, commit -> mbarrier, and TMEM -> register loads.
//
// Canonical K-major, no-swizzle operand tile X[R rows][K cols] of 32-bit elements:
//   "core matrix" = 8 rows x 16 bytes (4 elements), stored as 128 contiguous bytes;
//   byte offset(r,k) = (r/8)*SBO + (k/4)*LBO + (r%8)*16 + (k%4)*4
// with LBO = 128 (next core matrix along K) and SBO = (K/4)*128 (next 8-row group).
// One MMA consumes K = 8 elements (two core matrices along K); the descriptor start address
// advances by 256 bytes per k-step.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace s3g {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// element (r,k) -> float index inside a canonical tile with K columns
__device__ __forceinline__ int canon_idx(int r, int k, int K) {
    return (r >> 3) * (K >> 2) * 32 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3);
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);            // start address      bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;      // leading byte offset bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;      // stride byte offset  bits [32,46)
    d |= (uint64_t)1 << 46;                                 // descriptor version 1 (Blackwell)
    return d;                                               // base offset 0, layout type 0 = no swizzle
}

// kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4)        // c_format  = F32
         | (2u << 7)        // a_format  = TF32
         | (2u << 10)       // b_format  = TF32
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// same, with the operand majors chosen per operand: bit 15 / bit 16 of the instruction descriptor select
// MN-major ("transposed") A / B.  A canonical K-major tile X[R][K] read as an MN-major operand is X^T:
//   MN index = k (4 contiguous elements, groups of 4 every 128 bytes  -> SBO' = 128),
//   K  index = r (8 rows 16 bytes apart = one core matrix; next 8-row group (K/4)*128 bytes on -> LBO').
// One MMA consumes K' = 8 rows; the descriptor start address advances by (K/4)*128 bytes per k-step.
// (Used by the weight-gradient products delta^T * act of the backward decoder; NOT yet run on hardware -
// tools/dev_umma.py --mn is the first thing to run in round 2.)
__host__ __device__ constexpr uint32_t make_idesc_tf32_major(int M, int N, bool a_mn, bool b_mn) {
    return make_idesc_tf32(M, N) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16);
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T ; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// all previously issued MMAs of this thread -> arrive on the mbarrier when complete
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// this warp's 32 TMEM lanes x 32 consecutive columns -> 32 registers per thread (thread = lane = row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace umma
}  // namespace s3g
