// peer.cuh - gradient all-reduce over peer-mapped memory (NVLink 5 / NVSwitch), two kernels:
//   reduce-scatter : rank r sums slice r of every rank's buffer with direct peer loads and keeps the result
//   all-gather     : every rank copies the reduced slices from their owners
// The buffers are one symmetric allocation per rank (same size, peer-mapped into every process); the
// cross-rank ordering (contributions written -> reduce -> gather -> buffer reusable) is provided by the
// caller's device-side barriers between the launches, so these kernels never wait on anything: they are
// pure load/store streams.  Every element is summed in rank order by exactly one rank, so all ranks end up
// with bit-identical results.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace s3g {

constexpr int PEER_MAX_RANKS = 16;
constexpr int PEER_THREADS = 512;

struct PeerArgs {
    float* buf[PEER_MAX_RANKS];   // peer-mapped base pointers, buf[rank] is local
    int world, rank;
    long long n4;                 // float4 elements in the whole buffer
    long long chunk4;             // float4 elements per rank slice (last slice may be shorter)
};

__device__ __forceinline__ float4 ld_peer(const float4* p) {
    float4 v;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int WORLD>
__global__ void __launch_bounds__(PEER_THREADS) peer_reduce_scatter_kernel(const __grid_constant__ PeerArgs a) {
    const int world = WORLD > 0 ? WORLD : a.world;
    const long long lo = (long long)a.rank * a.chunk4;
    const long long hi = (lo + a.chunk4 < a.n4) ? lo + a.chunk4 : a.n4;
    for (long long i = lo + (long long)blockIdx.x * PEER_THREADS + threadIdx.x; i < hi; i += (long long)gridDim.x * PEER_THREADS) {
        float4 v[WORLD > 0 ? WORLD : PEER_MAX_RANKS];
#pragma unroll
        for (int p = 0; p < (WORLD > 0 ? WORLD : PEER_MAX_RANKS); ++p)
            if (p < world) v[p] = ld_peer(reinterpret_cast<const float4*>(a.buf[p]) + i);
        float4 s = v[0];
#pragma unroll
        for (int p = 1; p < (WORLD > 0 ? WORLD : PEER_MAX_RANKS); ++p)
            if (p < world) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
        reinterpret_cast<float4*>(a.buf[a.rank])[i] = s;
    }
}

// grid.y = owner offset 1..world-1: this rank pulls the reduced slice of rank (rank + y) % world
__global__ void __launch_bounds__(PEER_THREADS) peer_all_gather_kernel(const __grid_constant__ PeerArgs a) {
    const int owner = (a.rank + 1 + blockIdx.y) % a.world;
    const long long lo = (long long)owner * a.chunk4;
    const long long hi = (lo + a.chunk4 < a.n4) ? lo + a.chunk4 : a.n4;
    const float4* src = reinterpret_cast<const float4*>(a.buf[owner]);
    float4* dst = reinterpret_cast<float4*>(a.buf[a.rank]);
    const long long stride = (long long)gridDim.x * PEER_THREADS;
    long long i = lo + (long long)blockIdx.x * PEER_THREADS + threadIdx.x;
    for (; i + 3 * stride < hi; i += 4 * stride) {          // four independent remote loads in flight per thread
        const float4 v0 = ld_peer(src + i), v1 = ld_peer(src + i + stride), v2 = ld_peer(src + i + 2 * stride),
                     v3 = ld_peer(src + i + 3 * stride);
        dst[i] = v0; dst[i + stride] = v1; dst[i + 2 * stride] = v2; dst[i + 3 * stride] = v3;
    }
    for (; i < hi; i += stride) dst[i] = ld_peer(src + i);
}

// ---- NVLS variant (in-switch reduction), one kernel ---------------------------------------------------------
// mc = the multicast address of the symmetric buffer (every rank's copy behind one pointer).  Rank r owns slice
// r: multimem.ld_reduce returns the SUM over all ranks of the addressed 16 bytes (the NVSwitch reduces),
// multimem.st writes the result into every rank's copy.  Half the NVLink traffic of the peer-load version.
// NOT yet run on hardware (written at the end of round 1 without GPU budget; tools/dev_peer.py --nvls checks
// it against NCCL and times it) - bench.py does not use it.
__global__ void __launch_bounds__(PEER_THREADS) peer_nvls_all_reduce_kernel(float* mc, int rank, long long n4,
                                                                           long long chunk4) {
    const long long lo = (long long)rank * chunk4;
    const long long hi = (lo + chunk4 < n4) ? lo + chunk4 : n4;
    for (long long i = lo + (long long)blockIdx.x * PEER_THREADS + threadIdx.x; i < hi; i += (long long)gridDim.x * PEER_THREADS) {
        float4 v;
        float* p = mc + 4 * i;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                     ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
}

// ---- second half of the exchange that preprocess_backward_kernel<.., DP> starts (PeerSink, preprocess.cuh) --------
// Every rank's visible gradient rows already sit in the staging area of the slice's owner, one sub-slice per source
// rank: stage = [world][chunk] floats, local.  The owner adds the `world` sub-slices in rank order (every element is
// summed by exactly one rank in a fixed order -> all ranks end up with identical bits), stores the result into
// EVERY rank's gradient bucket - one multimem.st through the multicast mapping where the allocation has one (the
// NVSwitch replicates it), direct peer stores otherwise - and leaves the staging sub-slices zeroed for the next step.
struct ReduceGatherArgs {
    float* stage;                       // local staging [world][chunk]
    float* bucket[PEER_MAX_RANKS];      // peer-mapped gradient buckets (used when mc == nullptr)
    float* mc;                          // multicast address of the bucket, or nullptr
    int world, rank;
    long long chunk4;                   // float4 per slice
    long long n4;                       // float4 in the whole bucket (the last slice may be shorter)
};

template <int WORLD>
__global__ void __launch_bounds__(PEER_THREADS) peer_reduce_gather_kernel(const __grid_constant__ ReduceGatherArgs a) {
    const int world = WORLD > 0 ? WORLD : a.world;
    const long long lo = (long long)a.rank * a.chunk4;
    const long long len = (lo + a.chunk4 <= a.n4) ? a.chunk4 : (a.n4 > lo ? a.n4 - lo : 0);
    float4* st = reinterpret_cast<float4*>(a.stage);
    for (long long i = (long long)blockIdx.x * PEER_THREADS + threadIdx.x; i < len; i += (long long)gridDim.x * PEER_THREADS) {
        float4 v[WORLD > 0 ? WORLD : PEER_MAX_RANKS];
#pragma unroll
        for (int p = 0; p < (WORLD > 0 ? WORLD : PEER_MAX_RANKS); ++p)
            if (p < world) v[p] = st[(long long)p * a.chunk4 + i];
        float4 s = v[0];
#pragma unroll
        for (int p = 1; p < (WORLD > 0 ? WORLD : PEER_MAX_RANKS); ++p)
            if (p < world) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < (WORLD > 0 ? WORLD : PEER_MAX_RANKS); ++p)
            if (p < world) st[(long long)p * a.chunk4 + i] = z;
        if (a.mc) {
            float* dst = a.mc + 4 * (lo + i);
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                         ::"l"(dst), "f"(s.x), "f"(s.y), "f"(s.z), "f"(s.w) : "memory");
        } else {
#pragma unroll
            for (int p = 0; p < (WORLD > 0 ? WORLD : PEER_MAX_RANKS); ++p)
                if (p < world) reinterpret_cast<float4*>(a.bucket[p])[lo + i] = s;
        }
    }
}

}  // namespace s3g
