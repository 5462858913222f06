// binning.cuh - tile-instance emission and per-tile ranges (sm_100a).
//
// The reference duplicates every visible Gaussian into one (tile<<32|depth)
// key per touched tile in Gaussian-index order and radix-sorts all R 64-bit
// keys on 32+log2(tiles) bits (DGR/cuda_rasterizer/rasterizer_impl.cu:70-111,
// :301-309).  An LSD radix sort processes the depth digits first; every
// instance of a Gaussian carries the same depth bits, so those four digit
// passes are hoisted in front of the duplication and run over the P Gaussians
// instead of the R instances.  Emission then walks the depth-sorted Gaussians
// and only the tile-id digits remain to be sorted over R (two passes at
// 1920x1280).  All passes are stable, ties in depth keep Gaussian-index order,
// so the resulting point_list is bit-identical to the reference's.
#pragma once
#include "common.cuh"

namespace s3g {

// One warp per 32 consecutive depth-sorted Gaussians; the warp's output segment
// is contiguous, so lanes stride over it and binary-search the owner
// (perfectly balanced, coalesced stores; replaces the serial per-Gaussian loop
// of duplicateWithKeys, rasterizer_impl.cu:98-109).
__global__ void __launch_bounds__(256)
emit_instances_kernel(uint32_t P, const uint32_t* __restrict__ order,
                      const uint32_t* __restrict__ offsets,
                      const uint32_t* __restrict__ tiles_touched, const ushort4* __restrict__ rect,
                      int grid_x, uint32_t* __restrict__ inst_tile, uint32_t* __restrict__ inst_idx) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    // lanes past the end carry off = UINT_MAX so the search never selects them
    uint32_t g = 0, n = 0, off = 0xFFFFFFFFu, end = 0;
    uint32_t rx = 0, ry = 0, rw = 1;
    if (k < P) {
        g = order[k];
        n = tiles_touched[g];
        off = offsets[k];
        end = off + n;
        if (n) {
            ushort4 r = rect[g];
            rx = r.x;
            ry = r.y;
            rw = (uint32_t)(r.z - r.x);
        }
    }
    const uint32_t warp_base = __shfl_sync(0xffffffffu, off, 0);   // lane 0 is always valid
    const uint32_t warp_end = __reduce_max_sync(0xffffffffu, end);
    const uint32_t total = warp_end > warp_base ? warp_end - warp_base : 0u;   // all-invalid warp -> 0
    for (uint32_t j0 = 0; j0 < total; j0 += 32) {
        const uint32_t j = j0 + lane;
        const uint32_t target = warp_base + j;
        // largest lane L with off_L <= target
        int L = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            uint32_t o = __shfl_sync(0xffffffffu, off, L + step);   // L+step <= 31 by construction
            if (o <= target) L += step;
        }
        const uint32_t oL = __shfl_sync(0xffffffffu, off, L);
        const uint32_t gL = __shfl_sync(0xffffffffu, g, L);
        const uint32_t xL = __shfl_sync(0xffffffffu, rx, L);
        const uint32_t yL = __shfl_sync(0xffffffffu, ry, L);
        const uint32_t wL = __shfl_sync(0xffffffffu, rw, L);
        if (j < total) {
            const uint32_t i = target - oL;
            const uint32_t ty = yL + i / wL;
            const uint32_t tx = xL + i % wL;
            inst_tile[target] = ty * (uint32_t)grid_x + tx;
            inst_idx[target] = gL;
        }
    }
}

// rasterizer_impl.cu:116-138 (ranges pre-zeroed by the caller, :311)
__global__ void __launch_bounds__(256)
tile_ranges_kernel(uint32_t L, const uint32_t* __restrict__ tiles, uint2* __restrict__ ranges) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    const uint32_t cur = tiles[idx];
    if (idx == 0) {
        ranges[cur].x = 0;
    } else {
        const uint32_t prev = tiles[idx - 1];
        if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
    }
    if (idx == L - 1) ranges[cur].y = L;
}

}  // namespace s3g
