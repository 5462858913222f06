// umma_probe.cu - descriptor probe for the tcgen05 building blocks: the HOST supplies the raw shared-memory images
// of both operands, every descriptor field and the per-k-step address advances; the kernel issues the MMAs and
// returns the 128 x N accumulator.  tools/dev_umma.py --probe uses it to find out, on hardware, which
// (major-ness, swizzle, LBO, SBO) combinations address which shared-memory words - e.g. with B = identity the
// result row m lists exactly the words the tensor core read as A(m, k).
#include <cstdio>

#include "../../include/s3g_b200.h"
#include "umma.cuh"

using namespace s3g::umma;

namespace {
struct ProbeArgs {
    const float* a_img;     // a_words floats, copied verbatim to shared memory (128-byte aligned base)
    const float* b_img;     // b_words floats
    int a_words, b_words;
    float* D;               // [128][N]
    int N, ksteps;
    uint32_t idesc;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;      // bytes
    uint32_t a_layout, b_layout;              // descriptor layout_type field (bits 61..63)
    uint32_t a_step, b_step;                  // bytes added to the start address per k-step
};

__device__ __forceinline__ uint64_t desc_of(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return make_smem_desc(addr, lbo, sbo) | ((uint64_t)(layout & 7u) << 61);
}

__global__ void __launch_bounds__(128) umma_probe_kernel(ProbeArgs p) {
    extern __shared__ __align__(1024) float smem[];
    float* sA = smem;
    float* sB = smem + ((p.a_words + 255) & ~255);      // 1 KB aligned
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < p.a_words; e += 128) sA[e] = p.a_img[e];
    for (int e = tid; e < p.b_words; e += 128) sB[e] = p.b_img[e];
    if (warp == 0) tmem_alloc(&tmem_base, 64);
    if (tid == 0) mbar_init(&bar, 1);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        for (int k = 0; k < p.ksteps; ++k) {
            const uint64_t ad = desc_of(smem_u32(sA) + k * p.a_step, p.a_lbo, p.a_sbo, p.a_layout);
            const uint64_t bd = desc_of(smem_u32(sB) + k * p.b_step, p.b_lbo, p.b_sbo, p.b_layout);
            mma_tf32(tmem, ad, bd, p.idesc, k > 0);
        }
        commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    for (int c0 = 0; c0 < p.N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (c0 + i < p.N) p.D[(size_t)tid * p.N + c0 + i] = v[i];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}
}  // namespace

// Development tool (tools/dev_umma.py --probe); descriptor fields are passed through unchecked on purpose.
extern "C" int s3g_umma_probe(const float* a_img, int a_words, const float* b_img, int b_words, float* D, int N,
                              int ksteps, unsigned idesc, unsigned a_lbo, unsigned a_sbo, unsigned a_layout,
                              unsigned a_step, unsigned b_lbo, unsigned b_sbo, unsigned b_layout, unsigned b_step,
                              void* stream) {
    if (N % 8 || N > 64 || N < 8 || a_words <= 0 || b_words <= 0) return S3G_ERR_ARG;
    const size_t smem = (size_t)(((a_words + 255) & ~255) + b_words) * sizeof(float) + 1024;
    if (smem > 220 * 1024) return S3G_ERR_ARG;
    if (cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return S3G_ERR_CUDA;
    ProbeArgs p{a_img, b_img, a_words, b_words, D, N, ksteps, idesc, a_lbo, a_sbo, b_lbo, b_sbo, a_layout, b_layout, a_step, b_step};
    umma_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
    return cudaGetLastError() == cudaSuccess ? S3G_OK : S3G_ERR_CUDA;
}
