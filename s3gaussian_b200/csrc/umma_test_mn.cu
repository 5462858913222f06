// umma_test_mn.cu - MN-major operand self-test of the tcgen05 building blocks (round-2 building block, see
// umma.cuh and docs/round2_tc_backward.md).  Its own translation unit so that umma_test.cu stays byte-identical
// to the binary the GPU tests validated.  NOT yet run on hardware.
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

// MN-major check: D[128 x N] = A^T * B with A: [128 g][128 m], B: [128 g][N], both row-major in global memory and
// stored in shared memory exactly like the decoder stores activations (canonical K-major tiles whose ROWS are the
// 128 Gaussians).  The reduction runs over the rows, so both operands are read through MN-major descriptors.
__global__ void __launch_bounds__(128) umma_selftest_mn_kernel(const float* A, const float* B, float* D, int N, int three_pass) {
    extern __shared__ __align__(128) float smem[];
    constexpr int MA = 128;
    float* sAh = smem;                   // [128 g][128 m] canonical
    float* sAl = sAh + 128 * MA;
    float* sBh = sAl + 128 * MA;         // [128 g][N] canonical
    float* sBl = sBh + 128 * N;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 128 * MA; e += 128) {
        const int r = e / MA, k = e - r * MA;
        float hi, lo;
        split(A[e], hi, lo);
        sAh[canon_idx(r, k, MA)] = three_pass ? hi : A[e];
        sAl[canon_idx(r, k, MA)] = lo;
    }
    for (int e = tid; e < 128 * N; e += 128) {
        const int r = e / N, k = e - r * N;
        float hi, lo;
        split(B[e], hi, lo);
        sBh[canon_idx(r, k, N)] = three_pass ? hi : B[e];
        sBl[canon_idx(r, k, N)] = lo;
    }
    if (warp == 0) tmem_alloc(&tmem_base, 64);
    if (tid == 0) mbar_init(&bar, 1);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32_major(128, N, true, true);
        const uint32_t stepA = (uint32_t)(MA / 4) * 128u, stepB = (uint32_t)(N / 4) * 128u;   // one 8-row group
        bool acc = false;
        for (int g0 = 0; g0 < 128; g0 += 8) {
            const uint32_t oa = (uint32_t)(g0 / 8) * stepA, ob = (uint32_t)(g0 / 8) * stepB;
            // MN-major: SBO' = 128 (next group of 4 along MN), LBO' = the 8-row group stride (unused at K' = 8)
            const uint64_t ah = make_smem_desc(smem_u32(sAh) + oa, stepA, 128), al = make_smem_desc(smem_u32(sAl) + oa, stepA, 128);
            const uint64_t bh = make_smem_desc(smem_u32(sBh) + ob, stepB, 128), bl = make_smem_desc(smem_u32(sBl) + ob, stepB, 128);
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

// D[128 m][N] = A^T B, A: [128][128], B: [128][N] (N % 16 == 0, N <= 64).  Round-2 building block; see umma.cuh.
extern "C" int s3g_umma_selftest_mn(const float* A, const float* B, float* D, int N, int three_pass, void* stream) {
    if (N % 16 || N > 64 || N < 16) return S3G_ERR_ARG;
    const size_t smem = (size_t)(2 * 128 * 128 + 2 * 128 * N) * sizeof(float);
    if (cudaFuncSetAttribute(umma_selftest_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return S3G_ERR_CUDA;
    umma_selftest_mn_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(A, B, D, N, three_pass);
    return cudaGetLastError() == cudaSuccess ? S3G_OK : S3G_ERR_CUDA;
}
