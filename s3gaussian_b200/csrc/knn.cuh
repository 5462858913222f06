// knn.cuh - mean squared distance to the 3 nearest neighbours of every point: the initialiser of the
// Gaussian scales (`distCUDA2` of submodules/simple-knn, called at scene/gaussian_model.py:153).
//
// Same result as the reference (exact 3-NN, self excluded, coincident points count with distance 0,
// (d0 + d1 + d2) / 3 with the squared distance written as dx*dx + dy*dy + dz*dz), different search
// structure: points are sorted along a 30-bit Morton curve with our onesweep radix sort, gathered into a
// sorted float4 array (coalesced scans instead of points[indices[i]] gathers), boxed in runs of 256 with a
// second level of 32-box super boxes; a query first bounds its answer with its six curve neighbours and
// then visits only (super) boxes whose AABB is nearer than the current third-best distance.  The bounding
// box of the cloud is reduced on the device - no host round trip anywhere.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace s3g {

constexpr int KNN_BOX = 256;        // points per box (= block size of the box kernel)
constexpr int KNN_SUPER = 32;       // boxes per super box

struct KnnBox { float lo[3], hi[3]; };

// floats mapped to unsigned ints whose order matches the float order (for atomicMin / atomicMax)
__device__ __forceinline__ uint32_t knn_ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float knn_unord(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// bounds[0..2] = min (ordered encoding), bounds[3..5] = max; initialised to 0xffffffff / 0 by the host memsets
__global__ void __launch_bounds__(256) knn_bounds_kernel(int P, const float* __restrict__ pts, uint32_t* bounds) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * 256 + threadIdx.x; i < P; i += gridDim.x * 256) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(pts + 3 * (size_t)i + c);
            lo[c] = fminf(lo[c], v);
            hi[c] = fmaxf(hi[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            atomicMin(bounds + c, knn_ord(lo[c]));
            atomicMax(bounds + 3 + c, knn_ord(hi[c]));
        }
    }
}

__device__ __forceinline__ uint32_t knn_spread10(uint32_t x) {     // 10 bits -> every third bit
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

__global__ void __launch_bounds__(256) knn_morton_kernel(int P, const float* __restrict__ pts,
                                                        const uint32_t* __restrict__ bounds, uint32_t* keys,
                                                        uint32_t* vals) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    uint32_t code = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float lo = knn_unord(bounds[c]), hi = knn_unord(bounds[3 + c]);
        const float ext = hi - lo;
        float t = ext > 0.f ? (__ldg(pts + 3 * (size_t)i + c) - lo) / ext : 0.f;
        t = fminf(fmaxf(t, 0.f), 1.f);
        code |= knn_spread10((uint32_t)(t * 1023.f)) << c;
    }
    keys[i] = code;
    vals[i] = (uint32_t)i;
}

// sorted[k] = {xyz of point order[k], its original index}; boxes[b] = AABB of sorted[b*256 .. b*256+255]
__global__ void __launch_bounds__(KNN_BOX) knn_box_kernel(int P, const float* __restrict__ pts,
                                                         const uint32_t* __restrict__ order, float4* sorted,
                                                         KnnBox* boxes) {
    __shared__ float s_lo[3][KNN_BOX / 32], s_hi[3][KNN_BOX / 32];
    const int k = blockIdx.x * KNN_BOX + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (k < P) {
        const uint32_t src = order[k];
        const float x = __ldg(pts + 3 * (size_t)src), y = __ldg(pts + 3 * (size_t)src + 1), z = __ldg(pts + 3 * (size_t)src + 2);
        sorted[k] = make_float4(x, y, z, __uint_as_float(src));
        lo[0] = hi[0] = x; lo[1] = hi[1] = y; lo[2] = hi[2] = z;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if ((threadIdx.x & 31) == 0) { s_lo[c][threadIdx.x >> 5] = lo[c]; s_hi[c][threadIdx.x >> 5] = hi[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        float a = s_lo[c][0], b = s_hi[c][0];
#pragma unroll
        for (int w = 1; w < KNN_BOX / 32; ++w) { a = fminf(a, s_lo[c][w]); b = fmaxf(b, s_hi[c][w]); }
        boxes[blockIdx.x].lo[c] = a;
        boxes[blockIdx.x].hi[c] = b;
    }
}

// super[s] = union of boxes[s*32 .. s*32+31]
__global__ void __launch_bounds__(32) knn_super_kernel(int nboxes, const KnnBox* __restrict__ boxes, KnnBox* super) {
    const int b = blockIdx.x * KNN_SUPER + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (b < nboxes) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { lo[c] = boxes[b].lo[c]; hi[c] = boxes[b].hi[c]; }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if (threadIdx.x == 0) { super[blockIdx.x].lo[c] = lo[c]; super[blockIdx.x].hi[c] = hi[c]; }
    }
}

// squared distance from p to the box (0 inside)
__device__ __forceinline__ float knn_box_dist(const KnnBox& b, float x, float y, float z) {
    const float dx = fmaxf(fmaxf(b.lo[0] - x, x - b.hi[0]), 0.f);
    const float dy = fmaxf(fmaxf(b.lo[1] - y, y - b.hi[1]), 0.f);
    const float dz = fmaxf(fmaxf(b.lo[2] - z, z - b.hi[2]), 0.f);
    return dx * dx + dy * dy + dz * dz;
}

// keep the three smallest squared distances, ascending
__device__ __forceinline__ void knn_push(float qx, float qy, float qz, const float4 c, float (&best)[3]) {
    const float dx = c.x - qx, dy = c.y - qy, dz = c.z - qz;
    float d = dx * dx + dy * dy + dz * dz;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (best[j] > d) { const float t = best[j]; best[j] = d; d = t; }
    }
}

__global__ void __launch_bounds__(256) knn_search_kernel(int P, const float4* __restrict__ sorted,
                                                        const KnnBox* __restrict__ boxes, int nboxes,
                                                        const KnnBox* __restrict__ super, int nsuper,
                                                        float* __restrict__ out) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= P) return;
    const float4 q = sorted[k];
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    // an upper bound of the third-nearest distance from the six neighbours along the curve
    for (int i = max(0, k - 3); i <= min(P - 1, k + 3); ++i)
        if (i != k) knn_push(q.x, q.y, q.z, sorted[i], best);
    const float reject = best[2];
    best[0] = best[1] = best[2] = FLT_MAX;
    for (int s = 0; s < nsuper; ++s) {
        const float ds = knn_box_dist(super[s], q.x, q.y, q.z);
        if (ds > reject || ds > best[2]) continue;
        const int b1 = min(nboxes, (s + 1) * KNN_SUPER);
        for (int b = s * KNN_SUPER; b < b1; ++b) {
            const float db = knn_box_dist(boxes[b], q.x, q.y, q.z);
            if (db > reject || db > best[2]) continue;
            const int i1 = min(P, (b + 1) * KNN_BOX);
            for (int i = b * KNN_BOX; i < i1; ++i)
                if (i != k) knn_push(q.x, q.y, q.z, sorted[i], best);
        }
    }
    out[__float_as_uint(q.w)] = (best[0] + best[1] + best[2]) / 3.0f;
}

}  // namespace s3g
