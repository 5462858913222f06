// umma_test.cu - self-test of the tcgen05 building blocks (umma.cuh): D[128 x N] = A[128 x K] * B[N x K]^T,
// once as a single TF32 pass and once as the 3xTF32 split used by the decoder.
#include <cstdio>

#include "../../include/s3g_b200.h"
#include "umma.cuh"

using namespace s3g::umma;

namespace {
__device__ __forceinline__ void split(float x, float& hi, float& lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    uint32_t l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
    lo = __uint_as_float(l);
}

// 128 threads. A: [128][K] row-major, B: [N][K] row-major, D: [128][N].  K % 8 == 0, N % 16 == 0, N <= 64.
__global__ void __launch_bounds__(128) umma_selftest_kernel(const float* A, const float* B, float* D, int K, int N, int three_pass) {
    extern __shared__ __align__(128) float smem[];
    float* sAh = smem;                   // [128][K] canonical
    float* sAl = sAh + 128 * K;
    float* sBh = sAl + 128 * K;          // [N][K] canonical
    float* sBl = sBh + N * K;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int e = tid; e < 128 * K; e += 128) {
        const int r = e / K, k = e - r * K;
        float hi, lo;
        split(A[e], hi, lo);
        sAh[canon_idx(r, k, K)] = three_pass ? hi : A[e];
        sAl[canon_idx(r, k, K)] = lo;
    }
    for (int e = tid; e < N * K; e += 128) {
        const int r = e / K, k = e - r * K;
        float hi, lo;
        split(B[e], hi, lo);
        sBh[canon_idx(r, k, K)] = three_pass ? hi : B[e];
        sBl[canon_idx(r, k, K)] = lo;
    }
    if (warp == 0) tmem_alloc(&tmem_base, 64);
    if (tid == 0) mbar_init(&bar, 1);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base;

    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        const uint32_t sbo = (uint32_t)(K / 4) * 128u;
        bool acc = false;
        for (int k0 = 0; k0 < K; k0 += 8) {
            const uint32_t off = (uint32_t)(k0 / 4) * 128u;
            const uint64_t ah = make_smem_desc(smem_u32(sAh) + off, 128, sbo), al = make_smem_desc(smem_u32(sAl) + off, 128, sbo);
            const uint64_t bh = make_smem_desc(smem_u32(sBh) + off, 128, sbo), bl = make_smem_desc(smem_u32(sBl) + off, 128, sbo);
            if (three_pass) {
                mma_tf32(tmem, al, bh, idesc, acc);
                mma_tf32(tmem, ah, bl, idesc, true);
                mma_tf32(tmem, ah, bh, idesc, true);
            } else {
                mma_tf32(tmem, ah, bh, idesc, acc);
            }
            acc = true;
        }
        commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    // thread = row (TMEM lane); warp w reads lanes 32w..32w+31
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (c0 + i < N) D[(size_t)tid * N + c0 + i] = v[i];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

}  // namespace

extern "C" int s3g_umma_selftest(const float* A, const float* B, float* D, int K, int N, int three_pass, void* stream) {
    if (K % 8 || N % 16 || N > 64 || K > 128) return S3G_ERR_ARG;
    const size_t smem = (size_t)(2 * 128 * K + 2 * N * K) * sizeof(float);
    if (cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return S3G_ERR_CUDA;
    umma_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(A, B, D, K, N, three_pass);
    return cudaGetLastError() == cudaSuccess ? S3G_OK : S3G_ERR_CUDA;
}
