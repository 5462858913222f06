// common.cuh - shared definitions for the sm_100a splatting kernels.
//
// Layout of the three opaque state arenas (the forward->backward contract,
// SURVEY.md 8a10; the reference carves its own at rasterizer_impl.cu:155-194).
// Everything per-Gaussian that the tile kernels gather is a 16-byte record so a
// gather is one aligned LDG.128 / one 32-byte sector.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace s3g {

constexpr int TILE_X = 16;   // DGR/cuda_rasterizer/config.h:15-17
constexpr int TILE_Y = 16;
constexpr int TILE_PIX = TILE_X * TILE_Y;

constexpr uint32_t DEPTH_KEY_INVISIBLE = 0xFFFFFFFFu;

// ---- arena carving -------------------------------------------------------
struct Carver {
    char* abase;   // 128-byte aligned base (NULL when only measuring)
    size_t off;    // always relative to abase
    bool measuring;
    __host__ explicit Carver(char* b) : off(0), measuring(b == nullptr) {
        abase = reinterpret_cast<char*>(((uintptr_t)b + 127) & ~(uintptr_t)127);
    }
    template <typename T>
    __host__ T* take(size_t count) {
        off = (off + 127) & ~(size_t)127;
        T* p = measuring ? nullptr : reinterpret_cast<T*>(abase + off);
        off += count * sizeof(T);
        return p;
    }
};

// radix sort geometry (radix_sort.cuh)
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;   // 4096 items per block
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_MAX_PASSES = 4;

__host__ __device__ inline int64_t div_up64(int64_t a, int64_t b) { return (a + b - 1) / b; }

struct SortTemp {
    uint32_t* hist;      // [SORT_MAX_PASSES][RADIX] global digit histograms
    uint32_t* status;    // [SORT_MAX_PASSES][nblk][RADIX] decoupled look-back words
    uint32_t* tickets;   // [SORT_MAX_PASSES] dynamic block ids (+ spare)
    size_t zero_bytes;   // bytes from hist to the end that must be zeroed before a sort
    __host__ static size_t bytes(int64_t n) {
        char* z = nullptr;
        Carver c(z);
        carve(c, n);
        return c.off + 128;
    }
    __host__ static SortTemp carve(Carver& c, int64_t n) {
        SortTemp t;
        int64_t nblk = div_up64(n > 0 ? n : 1, SORT_TILE);
        t.hist = c.take<uint32_t>((size_t)SORT_MAX_PASSES * RADIX);
        size_t start = c.off - (size_t)SORT_MAX_PASSES * RADIX * sizeof(uint32_t);
        t.tickets = c.take<uint32_t>(32);
        t.status = c.take<uint32_t>((size_t)SORT_MAX_PASSES * nblk * RADIX);
        t.zero_bytes = c.off - start;
        return t;
    }
};

// Per-Gaussian state (P-indexed).
struct GeomState {
    float4* xyAB;            // {pix.x, pix.y, conic.x, conic.y}          (forward.cu:252-254)
    float4* Cod;             // {conic.z, opacity, tau_cull, own index (bits)}
    float4* rgb;             // {r,g,b, view depth}: SH colour or copy of colors_precomp (forward.cu:243-246)
    float4* aux;             // {r,g,b,-} of the optional second colour set composited on the same lists
    uint32_t* depth_key;     // float bits of depth, or DEPTH_KEY_INVISIBLE
    uint32_t* tiles_touched; // forward.cu:255
    ushort4* rect;           // {min.x, min.y, max.x, max.y} of getRect (auxiliary.h:46-56)
    uint8_t* clamped;        // bit c set: channel c clamped at 0 (forward.cu:67-69)
    int* internal_radii;     // used when the caller passes radii == NULL
    float* grad_rec;         // [P][GRAD_REC] raster-gradient accumulators (backward only)
    // scratch (dead after forward)
    uint32_t* order_a;       // sort ping-pong: Gaussian ids
    uint32_t* key_b;
    uint32_t* order_b;
    uint32_t* offsets;       // exclusive scan of tiles_touched in depth order
    uint64_t* scan_status;   // chained-scan look-back words
    uint32_t* scan_misc;     // [0]=ticket, [1]=grid-barrier arrivals, [2..3]=total (u64): the instance count, read on the device by the binning
    SortTemp sort;
    size_t zero_bytes;       // scan_status .. end of sort temp: zeroed by ONE memset before preprocess
    __host__ static GeomState carve(char* base, int64_t P, size_t* total = nullptr) {
        Carver c(base);
        GeomState g;
        g.xyAB = c.take<float4>(P);
        g.Cod = c.take<float4>(P);
        g.rgb = c.take<float4>(P);
        g.aux = c.take<float4>(P);
        g.depth_key = c.take<uint32_t>(P);
        g.tiles_touched = c.take<uint32_t>(P);
        g.rect = c.take<ushort4>(P);
        g.clamped = c.take<uint8_t>(P);
        g.internal_radii = c.take<int>(P);
        g.grad_rec = c.take<float>((size_t)P * 16);
        g.order_a = c.take<uint32_t>(P);
        g.key_b = c.take<uint32_t>(P);
        g.order_b = c.take<uint32_t>(P);
        g.offsets = c.take<uint32_t>(P);
        g.scan_status = c.take<uint64_t>((size_t)div_up64(P > 0 ? P : 1, 2048) + 1);
        g.scan_misc = c.take<uint32_t>(32);
        g.sort = SortTemp::carve(c, P);
        g.zero_bytes = c.measuring ? 0 : (size_t)((c.abase + c.off) - reinterpret_cast<char*>(g.scan_status));
        if (total) *total = c.off + 128;
        return g;
    }
};

// Per-instance state, indexed by instance; carved for a CAPACITY >= num_rendered (the arena is sized before the
// count is known on the host).  point_list comes first so that its address does not depend on the capacity:
// it is the only field the backward pass reads.
struct BinningState {
    uint32_t* point_list;        // sorted Gaussian ids (rasterizer_impl.cu:184)
    uint32_t* tile_a;            // scratch ping-pong
    uint32_t* idx_a;
    uint32_t* tile_b;
    uint32_t* idx_b;
    SortTemp sort;
    __host__ static BinningState carve(char* base, int64_t R, size_t* total = nullptr) {
        Carver c(base);
        BinningState b;
        size_t n = (size_t)(R > 0 ? R : 1);
        b.point_list = c.take<uint32_t>(n);
        b.tile_a = c.take<uint32_t>(n);
        b.idx_a = c.take<uint32_t>(n);
        b.tile_b = c.take<uint32_t>(n);
        b.idx_b = c.take<uint32_t>(n);
        b.sort = SortTemp::carve(c, R);
        if (total) *total = c.off + 128;
        return b;
    }
};

// Per-pixel / per-tile state.
struct ImageState {
    float* final_T;        // accum_alpha (rasterizer_impl.cu:175)
    uint32_t* n_contrib;   // rasterizer_impl.cu:176
    uint2* ranges;         // per-tile [start,end) (rasterizer_impl.cu:177)
    uint32_t* tile_hist;   // instances per tile (scratch of the forward)
    __host__ static ImageState carve(char* base, int W, int H, size_t* total = nullptr) {
        Carver c(base);
        ImageState s;
        size_t N = (size_t)W * H;
        size_t tiles = (size_t)((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
        s.final_T = c.take<float>(N);
        s.n_contrib = c.take<uint32_t>(N);
        s.ranges = c.take<uint2>(tiles);
        s.tile_hist = c.take<uint32_t>(tiles);
        if (total) *total = c.off + 128;
        return s;
    }
};

// Raster-gradient record accumulated by the backward composite (one 64-byte
// line per Gaussian so a warp's reduced partial sums land in two sectors).
//   [0..1] dL/dmean2D.xy   [2..4] dL/dconic (x,y,w)   [5] dL/dopacity
//   [6..8] dL/dcolour      [9]    dL/ddepth           [10..12] dL/d(aux colour)   [13..15] pad
constexpr int GRAD_REC = 16;

// host launch-parameter helpers
struct TileGrid {
    int x, y;
    __host__ __device__ int count() const { return x * y; }
};
__host__ __device__ inline TileGrid tile_grid(int W, int H) {
    return TileGrid{(W + TILE_X - 1) / TILE_X, (H + TILE_Y - 1) / TILE_Y};
}

}  // namespace s3g
