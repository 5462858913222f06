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
//
// The kernel also counts the instances of every tile (shared-memory privatised when the tile
// table fits, SMEM_HIST): with that histogram the per-tile [start,end) ranges are an exclusive
// scan (tile_offsets_kernel) instead of a boundary search over the sorted keys
// (identifyTileRanges, rasterizer_impl.cu:116-138), and the digit histograms of the tile-id sort
// passes follow from it without another read of the keys.
//
// `cap` bounds the writes: the binning arena is sized before the instance count is known on the
// host (see s3g_rasterize_forward); an overflowing call is detected there and repeated.
template <bool SMEM_HIST>
__global__ void __launch_bounds__(512)
emit_instances_kernel(uint32_t P, const uint32_t* __restrict__ order,
                      const uint32_t* __restrict__ offsets,
                      const uint32_t* __restrict__ tiles_touched, const ushort4* __restrict__ rect,
                      int grid_x, uint32_t n_tiles, uint32_t cap, uint32_t* __restrict__ inst_tile,
                      uint32_t* __restrict__ inst_idx, uint32_t* __restrict__ tile_hist) {
    extern __shared__ uint32_t s_hist[];
    if (SMEM_HIST) {
        for (uint32_t i = threadIdx.x; i < n_tiles; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    for (uint32_t base = blockIdx.x * blockDim.x; base < P; base += gridDim.x * blockDim.x) {
        const uint32_t k = base + threadIdx.x;
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
                const uint32_t t = ty * (uint32_t)grid_x + tx;
                if (target < cap) {
                    inst_tile[target] = t;
                    inst_idx[target] = gL;
                }
                if (SMEM_HIST) atomicAdd(&s_hist[t], 1u);
                else atomicAdd(&tile_hist[t], 1u);
            }
        }
    }
    if (SMEM_HIST) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_tiles; i += blockDim.x) {
            const uint32_t c = s_hist[i];
            if (c) atomicAdd(&tile_hist[i], c);
        }
    }
}

// Exclusive scan of the tile histogram -> per-tile ranges (rasterizer_impl.cu:116-138: [start,end) of
// every touched tile, zeros elsewhere, :311) and the digit histograms of the npass 8-bit passes
// of the tile-id sort.  One block; every thread owns a run of consecutive tiles.
// If the instance count outgrew the capacity the lists are incomplete: every range is written empty, so the
// composite of this attempt draws the background only (the host repeats the tail with a larger arena).
__global__ void __launch_bounds__(1024)
tile_offsets_kernel(uint32_t n_tiles, const uint32_t* __restrict__ tile_hist, uint2* __restrict__ ranges,
                    int npass, int end_bit, uint32_t* __restrict__ digit_hist /*[npass][RADIX]*/,
                    const uint32_t* __restrict__ n_dev, uint32_t cap) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_dig[SORT_MAX_PASSES][RADIX];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < SORT_MAX_PASSES * RADIX; i += blockDim.x) (&s_dig[0][0])[i] = 0;
    const bool overflow = n_dev != nullptr && *n_dev > cap;
    const uint32_t per = (n_tiles + blockDim.x - 1) / blockDim.x;
    const uint32_t t0 = min(n_tiles, (uint32_t)tid * per), t1 = min(n_tiles, t0 + per);
    uint32_t sum = 0;
    for (uint32_t t = t0; t < t1; ++t) sum += tile_hist[t];
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();      // also orders the s_dig zeroing before the atomics below
    uint32_t run = inc - sum;
    for (int w = 0; w < warp; ++w) run += s_warp[w];
    for (uint32_t t = t0; t < t1; ++t) {
        const uint32_t c = tile_hist[t];
        ranges[t] = (c && !overflow) ? make_uint2(run, run + c) : make_uint2(0u, 0u);
        if (c) {
            for (int p = 0; p < npass; ++p) {
                const int shift = p * RADIX_BITS;
                const int bits = min(RADIX_BITS, end_bit - shift);
                atomicAdd(&s_dig[p][(t >> shift) & ((1u << bits) - 1u)], c);
            }
        }
        run += c;
    }
    __syncthreads();
    for (int i = tid; i < npass * RADIX; i += blockDim.x) digit_hist[i] = (&s_dig[0][0])[i];
}

}  // namespace s3g
