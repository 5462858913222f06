// deform_tc_bwd.cuh - the deformation decoder BACKWARD on the 5th-generation tensor cores.
//
// STATUS: DRAFT.  Written at the end of round 1 after the GPU budget was spent: it compiles for sm_100a
// but HAS NOT RUN ON HARDWARE.  Nothing calls it unless S3G_TC_BWD=1 is set (deformation.py); the validated
// backward is deform_backward_kernel (mma.sync, deform.cuh).  Design notes: docs/round2_tc_backward.md.
//
// Supported configuration: 4 HexPlane levels, heads pos + shs + dino on, scales / rotation / opacity heads off
// (the shipped defaults).  Everything else keeps the mma.sync kernel.
//
// CTA = 128 threads = one tile of 128 Gaussians; thread r owns Gaussian r = TMEM lane r, like the forward.
//   T1  forward layers        Y  = X  W^T   A = activation tile (K-major), B = prepared W           (K-major)
//   T2  delta propagation     dX = dY W     A = delta tile (K-major),      B = prepared W^T         (K-major)
//   T3  weight gradients      dW = dY^T X   A = delta tile, B = activation tile, BOTH read MN-major: the
//                             reduction runs over the tile's rows (the 128 Gaussians)
// all as 3xTF32 (lo*hi + hi*lo + hi*hi).  Shared memory: three 64 KB operand tiles (hi+lo) + one 32 KB weight
// buffer; TMEM: 512 columns.
#pragma once
#include "deform.cuh"
#include "umma.cuh"

namespace s3g {

constexpr int BWM = 128;                 // rows per tile
constexpr int BW_TMEM_COLS = 512;
constexpr int BW_TILE_FLOATS = 2 * BWM * 64;     // one operand buffer: [hi | lo], up to 128 x 64 each

// prepared (tf32 hi/lo, canonical K-major) matrices E[n][k] the kernel consumes as B operands
enum { BW_FA = 0, BW_FB,                 // W0[:, 0:64], W0[:, 64:128]                 [64][64]   (T1)
       BW_D0, BW_D2, BW_S1, BW_S2, BW_P1,               //                             (T1)
       BW_D2T, BW_D0T, BW_S2T, BW_S1T, BW_P1T,          // transposes                  (T2)
       BW_FAT, BW_FBT,                   // (W0[:, 0:64])^T, (W0[:, 64:128])^T          [64][64]   (T2)
       BW_COUNT };
struct BwTable {
    int off[BW_COUNT];     // float offset of the [hi | lo] block
    int n[BW_COUNT];       // rows (multiple of 16)
    int k[BW_COUNT];       // columns (multiple of 8)
    int total;
};
struct BwPrepArgs {
    const float* src[BW_COUNT];   // source matrix, row-major
    int stride[BW_COUNT];         // its row stride
    int col0[BW_COUNT];           // first source column used
    int transpose[BW_COUNT];      // E[n][k] = transpose ? S[k][col0 + n] : S[n][col0 + k]
    int n_valid[BW_COUNT];        // rows of E that exist in the source (rest zero)
    BwTable tab;
    float* dst;
};
__global__ void __launch_bounds__(256) bw_prep_weights_kernel(const __grid_constant__ BwPrepArgs p) {
    const int e = blockIdx.y;
    const int N = p.tab.n[e], K = p.tab.k[e];
    float* dst = p.dst + p.tab.off[e];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * K; i += gridDim.x * blockDim.x) {
        const int n = i / K, k = i - n * K;
        float v = 0.f;
        if (n < p.n_valid[e])
            v = p.transpose[e] ? __ldg(p.src[e] + (size_t)k * p.stride[e] + p.col0[e] + n)
                               : __ldg(p.src[e] + (size_t)n * p.stride[e] + p.col0[e] + k);
        uint32_t hi, lo;
        split_tf32_rna(v, hi, lo);
        const int ci = umma::canon_idx(n, k, K);
        dst[ci] = __uint_as_float(hi);
        dst[N * K + ci] = __uint_as_float(lo);
    }
}

struct DeformTcBwdArgs {
    DNet net;
    int P;
    const float *xyz, *scales, *rot, *opacity, *shs, *campos;
    int sh_degree;
    const float* features;                                  // [P][128] saved by the forward
    const float *g_means, *g_scales, *g_rot, *g_opacity, *g_colors, *g_dx, *g_dshs, *g_feat;   // NULL = zero
    float *d_scales, *d_rot, *d_opacity, *d_shs;            // per-Gaussian outputs written in full
    float* dxyz_direct;                                     // [P][3]: g_means + colour-direction term
    float* dfeatures;                                       // [P][128] for hexplane_scatter_kernel
    float* partial;                                         // [grid][off.total] per-CTA Linear gradients
    GradOff off;
    const float* wprep;
    BwTable tab;
};

struct BwCtx {
    float *xh, *x2, *dl, *w_sm;          // operand tiles ([hi | lo] each) and the weight buffer
    uint64_t *w_bar, *mma_bar;
    uint32_t wph, mph, tmem;
    const float* wprep;
    const BwTable* tab;
};

__device__ __forceinline__ void bw_fetch(BwCtx& c, int e) {      // thread 0
    const uint32_t bytes = (uint32_t)(2 * c.tab->n[e] * c.tab->k[e]) * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(c.w_bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(umma::smem_u32(c.w_sm)), "l"(c.wprep + c.tab->off[e]), "r"(bytes), "r"(umma::smem_u32(c.w_bar)) : "memory");
}

// make the owners' tile stores visible to the tensor core and line the CTA up
__device__ __forceinline__ void bw_publish() {
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
}
__device__ __forceinline__ void bw_wait_mma(BwCtx& c) {           // all threads
    umma::mbar_wait(c.mma_bar, c.mph);
    umma::fence_after_sync();
    c.mph ^= 1;
}

// T1 / T2: D[128 x N] (+)= A_tile[128 x K] * E^T with E = prepared entry `e` ([N][K]) already requested.
// After completion the weights of `next` (or -1) start streaming in.
__device__ __forceinline__ void bw_layer(BwCtx& c, const float* a_tile, int K, int e, uint32_t col, bool accumulate, int next) {
    bw_publish();
    if (threadIdx.x == 0) {
        umma::mbar_wait(c.w_bar, c.wph);
        umma::fence_after_sync();
        const int N = c.tab->n[e];
        const uint32_t idesc = umma::make_idesc_tf32(BWM, N);
        const uint32_t sbo = (uint32_t)(K / 4) * 128u;
        const uint32_t a_hi = umma::smem_u32(a_tile), a_lo = a_hi + (uint32_t)(BWM * K) * 4u;
        const uint32_t b_hi = umma::smem_u32(c.w_sm), b_lo = b_hi + (uint32_t)(N * K) * 4u;
        for (int k0 = 0; k0 < K; k0 += 8) {
            const uint32_t off = (uint32_t)(k0 / 4) * 128u;
            const uint64_t dah = umma::make_smem_desc(a_hi + off, 128, sbo), dal = umma::make_smem_desc(a_lo + off, 128, sbo);
            const uint64_t dbh = umma::make_smem_desc(b_hi + off, 128, sbo), dbl = umma::make_smem_desc(b_lo + off, 128, sbo);
            umma::mma_tf32(c.tmem + col, dal, dbh, idesc, accumulate || k0 > 0);
            umma::mma_tf32(c.tmem + col, dah, dbl, idesc, true);
            umma::mma_tf32(c.tmem + col, dah, dbh, idesc, true);
        }
        umma::commit(c.mma_bar);
    }
    c.wph ^= 1;
    bw_wait_mma(c);
    if (threadIdx.x == 0 && next >= 0) bw_fetch(c, next);
}

// T3: D[128 x N] = A_tile^T * B_tile, A_tile: [128 g][Ka], B_tile: [128 g][N] (canonical K-major tiles whose rows
// are the Gaussians), both read through MN-major descriptors (umma.cuh).  Rows >= Ka of D are garbage.
__device__ __forceinline__ void bw_wgrad(BwCtx& c, const float* a_tile, int Ka, const float* b_tile, int N, uint32_t col) {
    bw_publish();
    if (threadIdx.x == 0) {
        umma::fence_after_sync();
        const uint32_t idesc = umma::make_idesc_tf32_major(BWM, N, true, true);
        const uint32_t stepA = (uint32_t)(Ka / 4) * 128u, stepB = (uint32_t)(N / 4) * 128u;
        const uint32_t a_hi = umma::smem_u32(a_tile), a_lo = a_hi + (uint32_t)(BWM * Ka) * 4u;
        const uint32_t b_hi = umma::smem_u32(b_tile), b_lo = b_hi + (uint32_t)(BWM * N) * 4u;
        for (int g0 = 0; g0 < BWM; g0 += 8) {
            const uint32_t oa = (uint32_t)(g0 / 8) * stepA, ob = (uint32_t)(g0 / 8) * stepB;
            const uint64_t dah = umma::make_smem_desc(a_hi + oa, stepA, 128), dal = umma::make_smem_desc(a_lo + oa, stepA, 128);
            const uint64_t dbh = umma::make_smem_desc(b_hi + ob, stepB, 128), dbl = umma::make_smem_desc(b_lo + ob, stepB, 128);
            umma::mma_tf32(c.tmem + col, dal, dbh, idesc, g0 > 0);
            umma::mma_tf32(c.tmem + col, dah, dbl, idesc, true);
            umma::mma_tf32(c.tmem + col, dah, dbh, idesc, true);
        }
        umma::commit(c.mma_bar);
    }
    bw_wait_mma(c);
}

// this thread's row of a tile of width K: values -> hi / lo halves of the canonical tile
template <int K, int NV, bool RELU>
__device__ __forceinline__ void bw_store_row(float* tile, int row, const float (&v)[NV]) {
    static_assert(NV <= K && NV % 4 == 0, "row width");
    float* hi = tile;
    float* lo = tile + BWM * K;
#pragma unroll
    for (int k = 0; k < K; k += 4) {
        float a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = (k + i < NV) ? v[(k + i < NV) ? k + i : 0] : 0.f;
            if (RELU) a[i] = fmaxf(a[i], 0.f);
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(a[i], h[i], l[i]);
        const int ci = umma::canon_idx(row, k, K);
        *reinterpret_cast<uint4*>(hi + ci) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(lo + ci) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// this thread's row of TMEM columns [col, col + 64)
__device__ __forceinline__ void bw_load_row64(const BwCtx& c, uint32_t col, float (&v)[64]) {
    const uint32_t lane_base = (uint32_t)((threadIdx.x >> 5) * 32) << 16;
    float t[32];
    umma::tmem_ld32(c.tmem + lane_base + col, t);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = t[i];
    umma::tmem_ld32(c.tmem + lane_base + col + 32, t);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[32 + i] = t[i];
}

// weight gradient in TMEM columns [col, col + 64): rows < m_valid -> RED into the CTA's partial buffer,
// dst[m * dst_stride + n]
__device__ __forceinline__ void bw_flush_wgrad(const BwCtx& c, uint32_t col, int m_valid, float* dst, int dst_stride) {
    if ((int)threadIdx.x < ((m_valid + 31) & ~31)) {      // whole warps take part in the TMEM load
        float v[64];
        bw_load_row64(c, col, v);
        if ((int)threadIdx.x < m_valid) {
            float* d = dst + (size_t)threadIdx.x * dst_stride;
#pragma unroll
            for (int n = 0; n < 64; ++n) atomicAdd(d + n, v[n]);
        }
    }
    umma::fence_before_sync();
}

// column sums of a [32 lanes][64] register tile: after the butterfly lane l holds the sums of columns 2l, 2l+1
// (bias gradients; each warp keeps its own running totals and REDs them once at the end of the kernel)
__device__ __forceinline__ void bw_colsum64(const float (&v)[64], float& s0, float& s1) {
    const uint32_t F = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float a[32];
    {   // 64 -> 32: lanes with bit 4 set keep the upper half
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float send = up ? v[i] : v[i + 32], keep = up ? v[i + 32] : v[i];
            a[i] = keep + __shfl_xor_sync(F, send, 16);
        }
    }
    float b[16];
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float send = up ? a[i] : a[i + 16], keep = up ? a[i + 16] : a[i];
            b[i] = keep + __shfl_xor_sync(F, send, 8);
        }
    }
    float d[8];
    {
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float send = up ? b[i] : b[i + 8], keep = up ? b[i + 8] : b[i];
            d[i] = keep + __shfl_xor_sync(F, send, 4);
        }
    }
    float e[4];
    {
        const bool up = lane & 2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = up ? d[i] : d[i + 4], keep = up ? d[i + 4] : d[i];
            e[i] = keep + __shfl_xor_sync(F, send, 2);
        }
    }
    {
        const bool up = lane & 1;
        const float send0 = up ? e[0] : e[2], keep0 = up ? e[2] : e[0];
        const float send1 = up ? e[1] : e[3], keep1 = up ? e[3] : e[1];
        s0 += keep0 + __shfl_xor_sync(F, send0, 1);
        s1 += keep1 + __shfl_xor_sync(F, send1, 1);
    }
}
// column owned by (lane, slot) after bw_colsum64: bit4 -> +32, bit3 -> +16, bit2 -> +8, bit1 -> +4, bit0 -> +2, slot -> +1
__device__ __forceinline__ int bw_colsum_column(int lane, int slot) {
    return ((lane & 16) ? 32 : 0) + ((lane & 8) ? 16 : 0) + ((lane & 4) ? 8 : 0) + ((lane & 2) ? 4 : 0) + ((lane & 1) ? 2 : 0) + slot;
}

enum { BB_FEAT = 0, BB_D0, BB_D2, BB_S1, BB_S2, BB_P1, BB_COUNT };     // bias-gradient slots kept per warp

__global__ void __launch_bounds__(BWM, 1) deform_backward_tc_kernel(const __grid_constant__ DeformTcBwdArgs a) {
    extern __shared__ __align__(128) float s_dyn[];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[6 * 64];         // b_feat, b_d0, b_d2, b_s1, b_s2 (48), b_p1
    const DNet& n = a.net;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    BwCtx c;
    c.xh = s_dyn;
    c.x2 = s_dyn + BW_TILE_FLOATS;
    c.dl = s_dyn + 2 * BW_TILE_FLOATS;
    c.w_sm = s_dyn + 3 * BW_TILE_FLOATS;
    c.w_bar = &s_bar[0]; c.mma_bar = &s_bar[1];
    c.wph = 0; c.mph = 0;
    c.wprep = a.wprep; c.tab = &a.tab;

    {
        const float* bp[6] = {n.b_feat, n.b_d0, n.b_d2, n.shs.b1, n.shs.b2, n.pos.b1};
        const int bn[6] = {64, 64, 64, 64, 48, 64};
#pragma unroll
        for (int l = 0; l < 6; ++l)
            if (tid < 64) s_bias[l * 64 + tid] = tid < bn[l] ? bp[l][tid] : 0.f;
    }
    float* part = a.partial + (size_t)blockIdx.x * a.off.total;
    for (int i = tid; i < a.off.total; i += BWM) part[i] = 0.f;
    if (warp == 0) umma::tmem_alloc(&s_tmem, BW_TMEM_COLS);
    if (tid == 0) { umma::mbar_init(&s_bar[0], 1); umma::mbar_init(&s_bar[1], 1); }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    c.tmem = s_tmem;
    if (tid == 0) bw_fetch(c, BW_FA);

    // TMEM columns
    constexpr uint32_t C_H = 0, C_O1 = 64, C_O2 = 128, C_DHR = 192 /* relu(H) consumers */, C_DHD = 256 /* dino */,
                       C_WG = 320 /* weight gradients */;
    float bsum[BB_COUNT][2];                 // running bias-gradient column sums of this warp
#pragma unroll
    for (int i = 0; i < BB_COUNT; ++i) bsum[i][0] = bsum[i][1] = 0.f;
    float b2_pos[3] = {0.f, 0.f, 0.f}, b4_dino[3] = {0.f, 0.f, 0.f};   // tiny-head bias gradients (lane-local partials)

    const int ntiles = (a.P + BWM - 1) / BWM;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int gi = tile * BWM + tid;
        const bool valid = gi < a.P;
        const float* frow = a.features + (size_t)(valid ? gi : 0) * 128;

        // ---- 1. H = F W0^T + b0, two 64-column halves of F through X2 ---------------------------------
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float f[64];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 v = valid ? __ldg(reinterpret_cast<const float4*>(frow + 64 * half) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
            }
            bw_store_row<64, 64, false>(c.x2, tid, f);
            bw_layer(c, c.x2, 64, half == 0 ? BW_FA : BW_FB, C_H, half == 1, half == 0 ? BW_FB : BW_D0);
        }
        uint64_t mask_h = 0;                  // H > 0 (ReLU mask of the pos / shs heads' input)
        {
            float h[64];
            bw_load_row64(c, C_H, h);
#pragma unroll
            for (int i = 0; i < 64; ++i) { h[i] += s_bias[0 * 64 + i]; if (h[i] > 0.f) mask_h |= 1ull << i; }
            // ---- 2. dino head on H (no leading ReLU) ---------------------------------------------------
            bw_store_row<64, 64, false>(c.xh, tid, h);
        }
        bw_layer(c, c.xh, 64, BW_D0, C_O1, false, BW_D2);
        uint64_t mask_a0 = 0;
        {
            float a0[64];
            bw_load_row64(c, C_O1, a0);
#pragma unroll
            for (int i = 0; i < 64; ++i) { a0[i] += s_bias[1 * 64 + i]; if (a0[i] > 0.f) mask_a0 |= 1ull << i; }
            bw_store_row<64, 64, true>(c.x2, tid, a0);                   // relu(A0)
        }
        bw_layer(c, c.x2, 64, BW_D2, C_O2, false, BW_D2T);
        float gf[3] = {0.f, 0.f, 0.f};
        if (valid && a.g_feat) { gf[0] = a.g_feat[(size_t)gi * 3]; gf[1] = a.g_feat[(size_t)gi * 3 + 1]; gf[2] = a.g_feat[(size_t)gi * 3 + 2]; }
        {
            float a2[64];
            bw_load_row64(c, C_O2, a2);
            float d2[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                a2[j] += s_bias[2 * 64 + j];
                const float s = gf[0] * __ldg(n.w_d4 + j) + gf[1] * __ldg(n.w_d4 + 64 + j) + gf[2] * __ldg(n.w_d4 + 128 + j);
                d2[j] = a2[j] > 0.f ? s : 0.f;
            }
            // dW_d4 = g_feat^T relu(A2): a 16-wide delta tile (3 columns used) against relu(A2) - needs X2, which still
            // holds relu(A0) for dW_d2: first dW_d2 and the push to A0, then relu(A2) replaces relu(A0)
            bw_store_row<64, 64, false>(c.dl, tid, d2);                  // delta A2
            bw_colsum64(d2, bsum[BB_D2][0], bsum[BB_D2][1]);
            bw_wgrad(c, c.dl, 64, c.x2, 64, C_WG);                       // dW_d2 = dA2^T relu(A0)
            bw_flush_wgrad(c, C_WG, 64, part + a.off.d2w, 64);
            bw_layer(c, c.dl, 64, BW_D2T, C_O1, false, BW_D0T);          // dA0 (pre-mask) = dA2 W_d2
            // now X2 <- relu(A2), DL <- g_feat (16 wide)
            bw_store_row<64, 64, true>(c.x2, tid, a2);
            float g16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) g16[i] = i < 3 ? gf[i] : 0.f;
            bw_store_row<16, 16, false>(c.dl, tid, g16);
            bw_wgrad(c, c.dl, 16, c.x2, 64, C_WG);                       // dW_d4 = g_feat^T relu(A2)
            bw_flush_wgrad(c, C_WG, 3, part + a.off.d4w, 64);
#pragma unroll
            for (int i = 0; i < 3; ++i) b4_dino[i] += gf[i];
        }
        {
            float d0[64];
            bw_load_row64(c, C_O1, d0);
#pragma unroll
            for (int j = 0; j < 64; ++j) d0[j] = (mask_a0 >> j) & 1ull ? d0[j] : 0.f;
            bw_store_row<64, 64, false>(c.dl, tid, d0);
            bw_colsum64(d0, bsum[BB_D0][0], bsum[BB_D0][1]);
        }
        bw_wgrad(c, c.dl, 64, c.xh, 64, C_WG);                           // dW_d0 = dA0^T H
        bw_flush_wgrad(c, C_WG, 64, part + a.off.d0w, 64);
        bw_layer(c, c.dl, 64, BW_D0T, C_DHD, false, BW_S1);              // dH (dino part) = dA0 W_d0

        // ---- 3. shs head on relu(H) ---------------------------------------------------------------
        {
            float h[64];
            bw_load_row64(c, C_H, h);
#pragma unroll
            for (int i = 0; i < 64; ++i) h[i] += s_bias[0 * 64 + i];
            bw_store_row<64, 64, true>(c.xh, tid, h);                    // XH <- relu(H)
        }
        bw_layer(c, c.xh, 64, BW_S1, C_O1, false, BW_S2);
        uint64_t mask_as = 0;
        {
            float as[64];
            bw_load_row64(c, C_O1, as);
#pragma unroll
            for (int i = 0; i < 64; ++i) { as[i] += s_bias[3 * 64 + i]; if (as[i] > 0.f) mask_as |= 1ull << i; }
            bw_store_row<64, 64, true>(c.x2, tid, as);                   // relu(A_shs)
        }
        bw_layer(c, c.x2, 64, BW_S2, C_O2, false, BW_S2T);
        float dxyz[3] = {0.f, 0.f, 0.f};
        {
            // dshs -> shs_final -> colour clamp mask, dL/dshs_final = basis (x) g_colour, direction term of d_xyz
            float sf[64];
            bw_load_row64(c, C_O2, sf);                                  // columns 48..63 unused
            float bs[16], gr[3] = {0.f, 0.f, 0.f};
            int nb = 0;
            float vx = 0.f, vy = 0.f, vz = 0.f, s2 = 1.f, x = 0.f, y = 0.f, z = 1.f;
            if (valid) {
                vx = a.xyz[(size_t)gi * 3] - a.campos[0]; vy = a.xyz[(size_t)gi * 3 + 1] - a.campos[1];
                vz = a.xyz[(size_t)gi * 3 + 2] - a.campos[2];
                s2 = vx * vx + vy * vy + vz * vz;
                const float inv = 1.0f / sqrtf(s2);
                x = vx * inv; y = vy * inv; z = vz * inv;
                nb = sh_basis16(a.sh_degree, x, y, z, bs);
#pragma unroll
                for (int j = 0; j < 48; ++j) sf[j] += s_bias[4 * 64 + j] + a.shs[(size_t)gi * 48 + j];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    float r = 0.f;
                    for (int k = 0; k < nb; ++k) r = fmaf(bs[k], sf[3 * k + ch], r);
                    gr[ch] = (r + 0.5f > 0.0f) ? (a.g_colors ? a.g_colors[(size_t)gi * 3 + ch] : 0.f) : 0.f;
                }
                if (a.sh_degree > 0) {
                    const float C1 = 0.4886025119029199f;
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    float ddir[3] = {0.f, 0.f, 0.f};
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
#define SHF(i) sf[3 * (i) + ch]
                        float dx_ = -C1 * SHF(3), dy_ = -C1 * SHF(1), dz_ = C1 * SHF(2);
                        if (a.sh_degree > 1) {
                            dx_ += 1.0925484305920792f * y * SHF(4) + 0.31539156525252005f * -2.f * x * SHF(6) +
                                   -1.0925484305920792f * z * SHF(7) + 0.5462742152960396f * 2.f * x * SHF(8);
                            dy_ += 1.0925484305920792f * x * SHF(4) + -1.0925484305920792f * z * SHF(5) +
                                   0.31539156525252005f * -2.f * y * SHF(6) + 0.5462742152960396f * -2.f * y * SHF(8);
                            dz_ += -1.0925484305920792f * y * SHF(5) + 0.31539156525252005f * 4.f * z * SHF(6) +
                                   -1.0925484305920792f * x * SHF(7);
                        }
                        if (a.sh_degree > 2) {
                            dx_ += -0.5900435899266435f * SHF(9) * 6.f * xy + 2.890611442640554f * SHF(10) * yz +
                                   -0.4570457994644658f * SHF(11) * -2.f * xy + 0.3731763325901154f * SHF(12) * -6.f * xz +
                                   -0.4570457994644658f * SHF(13) * (-3.f * xx + 4.f * zz - yy) +
                                   1.445305721320277f * SHF(14) * 2.f * xz + -0.5900435899266435f * SHF(15) * 3.f * (xx - yy);
                            dy_ += -0.5900435899266435f * SHF(9) * 3.f * (xx - yy) + 2.890611442640554f * SHF(10) * xz +
                                   -0.4570457994644658f * SHF(11) * (-3.f * yy + 4.f * zz - xx) +
                                   0.3731763325901154f * SHF(12) * -6.f * yz + -0.4570457994644658f * SHF(13) * -2.f * xy +
                                   1.445305721320277f * SHF(14) * -2.f * yz + -0.5900435899266435f * SHF(15) * -6.f * xy;
                            dz_ += 2.890611442640554f * SHF(10) * xy + -0.4570457994644658f * SHF(11) * 8.f * yz +
                                   0.3731763325901154f * SHF(12) * 3.f * (2.f * zz - xx - yy) +
                                   -0.4570457994644658f * SHF(13) * 8.f * xz + 1.445305721320277f * SHF(14) * (xx - yy);
                        }
#undef SHF
                        ddir[0] += dx_ * gr[ch]; ddir[1] += dy_ * gr[ch]; ddir[2] += dz_ * gr[ch];
                    }
                    const float inv32 = 1.0f / sqrtf(s2 * s2 * s2);
                    const float dot = vx * ddir[0] + vy * ddir[1] + vz * ddir[2];
                    dxyz[0] = (s2 * ddir[0] - vx * dot) * inv32;
                    dxyz[1] = (s2 * ddir[1] - vy * dot) * inv32;
                    dxyz[2] = (s2 * ddir[2] - vz * dot) * inv32;
                }
            }
            float dsh[48];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float b = k < nb ? bs[k] : 0.f;
                dsh[3 * k] = b * gr[0]; dsh[3 * k + 1] = b * gr[1]; dsh[3 * k + 2] = b * gr[2];
            }
            if (valid) {
#pragma unroll
                for (int j = 0; j < 48; ++j) {
                    a.d_shs[(size_t)gi * 48 + j] = dsh[j];
                    dsh[j] += a.g_dshs ? a.g_dshs[(size_t)gi * 48 + j] : 0.f;
                }
            }
            bw_store_row<48, 48, false>(c.dl, tid, dsh);                 // delta dshs, 48 wide
            {   // bias gradient of the 48-wide layer through the 64-wide butterfly
                float p64[64];
#pragma unroll
                for (int j = 0; j < 64; ++j) p64[j] = j < 48 ? dsh[j] : 0.f;
                bw_colsum64(p64, bsum[BB_S2][0], bsum[BB_S2][1]);
            }
        }
        bw_wgrad(c, c.dl, 48, c.x2, 64, C_WG);                           // dW_s2 = d(dshs)^T relu(A_shs)
        bw_flush_wgrad(c, C_WG, 48, part + a.off.shs[2], 64);
        bw_layer(c, c.dl, 48, BW_S2T, C_O1, false, BW_S1T);              // dA_shs (pre-mask) = d(dshs) W_s2
        {
            float ds[64];
            bw_load_row64(c, C_O1, ds);
#pragma unroll
            for (int j = 0; j < 64; ++j) ds[j] = (mask_as >> j) & 1ull ? ds[j] : 0.f;
            bw_store_row<64, 64, false>(c.dl, tid, ds);
            bw_colsum64(ds, bsum[BB_S1][0], bsum[BB_S1][1]);
        }
        bw_wgrad(c, c.dl, 64, c.xh, 64, C_WG);                           // dW_s1 = dA_shs^T relu(H)
        bw_flush_wgrad(c, C_WG, 64, part + a.off.shs[0], 64);
        bw_layer(c, c.dl, 64, BW_S1T, C_DHR, false, BW_P1);              // dH (relu part) = dA_shs W_s1

        // ---- 4. pos head on relu(H) ---------------------------------------------------------------
        bw_layer(c, c.xh, 64, BW_P1, C_O1, false, BW_P1T);
        float ddx[3] = {0.f, 0.f, 0.f};
        if (valid) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
                ddx[i] = (a.g_means ? a.g_means[(size_t)gi * 3 + i] : 0.f) + (a.g_dx ? a.g_dx[(size_t)gi * 3 + i] : 0.f);
        }
        {
            float ap[64];
            bw_load_row64(c, C_O1, ap);
            float dp[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                ap[j] += s_bias[5 * 64 + j];
                const float s = ddx[0] * __ldg(n.pos.w2 + j) + ddx[1] * __ldg(n.pos.w2 + 64 + j) + ddx[2] * __ldg(n.pos.w2 + 128 + j);
                dp[j] = ap[j] > 0.f ? s : 0.f;
            }
            // dW_p2 = ddx^T relu(A_pos): relu(A_pos) -> X2, ddx (16 wide) -> DL, then DL <- dA_pos
            bw_store_row<64, 64, true>(c.x2, tid, ap);
            float g16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) g16[i] = i < 3 ? ddx[i] : 0.f;
            bw_store_row<16, 16, false>(c.dl, tid, g16);
            bw_wgrad(c, c.dl, 16, c.x2, 64, C_WG);
            bw_flush_wgrad(c, C_WG, 3, part + a.off.pos[2], 64);
#pragma unroll
            for (int i = 0; i < 3; ++i) b2_pos[i] += ddx[i];
            bw_store_row<64, 64, false>(c.dl, tid, dp);
            bw_colsum64(dp, bsum[BB_P1][0], bsum[BB_P1][1]);
        }
        bw_wgrad(c, c.dl, 64, c.xh, 64, C_WG);                           // dW_p1 = dA_pos^T relu(H)
        bw_flush_wgrad(c, C_WG, 64, part + a.off.pos[0], 64);
        bw_layer(c, c.dl, 64, BW_P1T, C_DHR, true, BW_FAT);              // dH (relu part) += dA_pos W_p1

        // ---- 5. dH, then the feature layer: dW0 = dH^T F, dF = dH W0 (two halves of F through X2) ------
        {
            float dh[64], dd[64];
            bw_load_row64(c, C_DHR, dh);
            bw_load_row64(c, C_DHD, dd);
#pragma unroll
            for (int j = 0; j < 64; ++j) dh[j] = ((mask_h >> j) & 1ull ? dh[j] : 0.f) + dd[j];
            bw_store_row<64, 64, false>(c.dl, tid, dh);
            bw_colsum64(dh, bsum[BB_FEAT][0], bsum[BB_FEAT][1]);
        }
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float f[64];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 v = valid ? __ldg(reinterpret_cast<const float4*>(frow + 64 * half) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
            }
            bw_store_row<64, 64, false>(c.x2, tid, f);
            bw_wgrad(c, c.dl, 64, c.x2, 64, C_WG);                       // dW0[:, half] = dH^T F_half
            bw_flush_wgrad(c, C_WG, 64, part + a.off.w_feat + 64 * half, 128);
            bw_layer(c, c.dl, 64, half == 0 ? BW_FAT : BW_FBT, C_O1, false, half == 0 ? BW_FBT : BW_FA);
            float df[64];
            bw_load_row64(c, C_O1, df);
            if (valid) {
                float4* out = reinterpret_cast<float4*>(a.dfeatures + (size_t)gi * 128 + 64 * half);
#pragma unroll
                for (int j = 0; j < 16; ++j) out[j] = make_float4(df[4 * j], df[4 * j + 1], df[4 * j + 2], df[4 * j + 3]);
            }
            umma::fence_before_sync();
        }

        // ---- 6. per-Gaussian outputs of the raw parameters (heads off: plain activation backward) ----------
        if (valid) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                a.dxyz_direct[(size_t)gi * 3 + i] = (a.g_means ? a.g_means[(size_t)gi * 3 + i] : 0.f) + dxyz[i];
                a.d_scales[(size_t)gi * 3 + i] = (a.g_scales ? a.g_scales[(size_t)gi * 3 + i] : 0.f) * expf(a.scales[(size_t)gi * 3 + i]);
            }
            float q[4], gq[4], nn = 0.f, dot = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) { q[i] = a.rot[(size_t)gi * 4 + i]; gq[i] = a.g_rot ? a.g_rot[(size_t)gi * 4 + i] : 0.f; nn += q[i] * q[i]; }
            const float nrm = sqrtf(nn);
            if (nrm > 1e-12f) {
                const float inv = 1.f / nrm;
#pragma unroll
                for (int i = 0; i < 4; ++i) dot += q[i] * inv * gq[i];
#pragma unroll
                for (int i = 0; i < 4; ++i) a.d_rot[(size_t)gi * 4 + i] = (gq[i] - q[i] * inv * dot) * inv;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) a.d_rot[(size_t)gi * 4 + i] = gq[i] * 1e12f;
            }
            const float sg = 1.0f / (1.0f + expf(-a.opacity[gi]));
            a.d_opacity[gi] = (a.g_opacity ? a.g_opacity[gi] : 0.f) * sg * (1.f - sg);
        }
    }

    // ---- bias gradients: each warp adds its column sums, tiny heads reduce over the warp first -----------
    {
        const int boff[BB_COUNT] = {a.off.b_feat, a.off.d0b, a.off.d2b, a.off.shs[1], a.off.shs[3], a.off.pos[1]};
        const int bn[BB_COUNT] = {64, 64, 64, 64, 48, 64};
#pragma unroll
        for (int l = 0; l < BB_COUNT; ++l) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int col = bw_colsum_column(lane, s);
                if (col < bn[l]) atomicAdd(part + boff[l] + col, bsum[l][s]);
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float p = b2_pos[i], d = b4_dino[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { p += __shfl_xor_sync(0xffffffffu, p, o); d += __shfl_xor_sync(0xffffffffu, d, o); }
            if (lane == 0) { atomicAdd(part + a.off.pos[3] + i, p); atomicAdd(part + a.off.d4b + i, d); }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(c.tmem, BW_TMEM_COLS);
}

}  // namespace s3g
