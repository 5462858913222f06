// composite.cuh - per-tile front-to-back alpha composite, forward and backward
// (colour + depth), for sm_100a.
//
// Replaces renderCUDA forward (DGR/cuda_rasterizer/forward.cu:261-379) and
// backward (backward.cu:415-590).  Same tiles, same sorted lists, same per-pixel
// arithmetic and thresholds (power > 0, alpha < 1/255, T < 1e-4), so colour,
// depth, final_T and n_contrib reproduce the reference.  What changes is how the
// work is organised:
//
//  * a 16x16 tile is 8 warps, each owning an 8x4 pixel rectangle;
//  * batches of 256 instances are gathered global->shared with cp.async
//    (16-byte records, double buffered) while the previous batch is composited;
//  * per 32 instances, every lane tests ONE instance against the warp's
//    rectangle (exact minimum of the conic's quadratic form over the rectangle
//    vs. the Gaussian's 1/255 cut-off, see cull_tau()) and the warp only walks
//    the ballot's survivors - the all-pairs loop of the reference
//    (forward.cu:328-366) is what makes it issue-bound;
//  * a warp leaves the loop as soon as all its pixels are saturated, the block
//    as soon as all warps are;
//  * backward: per surviving instance the 32 pixel partials of the 10 gradient
//    components are reduced with a 12-shuffle split butterfly and land as ONE
//    10-lane RED on a 64-byte record, instead of 10 atomics per (pixel,
//    Gaussian) pair (backward.cu:550-587).
#pragma once
#include "common.cuh"

namespace s3g {

constexpr int CB = 256;          // instances per staged batch
constexpr int SUB_W = 8;         // warp rectangle
constexpr int SUB_H = 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Minimum over the rectangle dx in [dxl,dxh], dy in [dyl,dyh] of
// q(dx,dy) = 0.5*(A dx^2 + C dy^2) + B dx dy  (= -power of forward.cu:338),
// for a positive-definite conic: the minimiser is the centre if it lies inside,
// otherwise it sits on one of the (at most two) edges facing the centre; the two
// clamped 1-D minimisations below cover every case.
__device__ __forceinline__ float rect_min_q(float A, float B, float C, float dxl, float dxh,
                                            float dyl, float dyh) {
    const float ddx = fminf(fmaxf(0.f, dxl), dxh);
    const float ddy = fminf(fmaxf(0.f, dyl), dyh);
    // edge dx = ddx, free dy
    const float y1 = fminf(fmaxf(__fdividef(-B * ddx, C), dyl), dyh);
    const float q1 = 0.5f * (A * ddx * ddx + C * y1 * y1) + B * ddx * y1;
    // edge dy = ddy, free dx
    const float x2 = fminf(fmaxf(__fdividef(-B * ddy, A), dxl), dxh);
    const float q2 = 0.5f * (A * x2 * x2 + C * ddy * ddy) + B * x2 * ddy;
    return fminf(q1, q2);
}

struct RenderFwdArgs {
    const uint2* ranges;
    const uint32_t* point_list;
    int W, H;
    const float4* xyAB;
    const float4* Cod;
    const float4* rgb;
    const float* bg;
    float* final_T;
    uint32_t* n_contrib;
    float* out_color;
    float* out_depth;
};

// One staged instance: 48 bytes = three 16-byte gathers.  Lane j of the cull
// test reads element j: LDS.128 at a 12-word stride is conflict-free per
// quarter-warp.
struct __align__(16) Rec {
    float4 a;   // pix.x, pix.y, conic.x, conic.y
    float4 b;   // conic.z, opacity, depth, tau_cull
    float4 c;   // r, g, b, -
};

__global__ void __launch_bounds__(TILE_PIX) render_forward_kernel(RenderFwdArgs a) {
    __shared__ Rec s_rec[2][CB];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int X0 = blockIdx.x * TILE_X + (warp & 1) * SUB_W;
    const int Y0 = blockIdx.y * TILE_Y + (warp >> 1) * SUB_H;
    const int px = X0 + (lane & 7), py = Y0 + (lane >> 3);
    const bool inside = px < a.W && py < a.H;
    const float pixx = (float)px, pixy = (float)py;
    const uint32_t pix_id = (uint32_t)a.W * py + px;
    // rectangle of d = xy - pix over the warp's pixels
    const float rx0 = (float)X0, rx1 = (float)(X0 + SUB_W - 1);
    const float ry0 = (float)Y0, ry1 = (float)(Y0 + SUB_H - 1);

    const uint2 range = a.ranges[tile];
    const int n = (int)(range.y - range.x);
    const int nb = (n + CB - 1) / CB;

    float T = 1.0f;
    uint32_t last_contributor = 0;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
    bool alive = inside;

    // prologue: gather batch 0, fetch the id of batch 1
    uint32_t id_next = 0;
    if (tid < n) {
        const uint32_t id = a.point_list[range.x + tid];
        cp_async16(&s_rec[0][tid].a, &a.xyAB[id]);
        cp_async16(&s_rec[0][tid].b, &a.Cod[id]);
        cp_async16(&s_rec[0][tid].c, &a.rgb[id]);
    }
    cp_async_commit();
    if (CB + tid < n) id_next = a.point_list[range.x + CB + tid];
    cp_async_wait_all();
    __syncthreads();

    for (int b = 0; b < nb; ++b) {
        const int buf = b & 1;
        // prefetch batch b+1 into the other buffer, and the ids of batch b+2
        if ((b + 1) * CB + tid < n) {
            cp_async16(&s_rec[buf ^ 1][tid].a, &a.xyAB[id_next]);
            cp_async16(&s_rec[buf ^ 1][tid].b, &a.Cod[id_next]);
            cp_async16(&s_rec[buf ^ 1][tid].c, &a.rgb[id_next]);
        }
        cp_async_commit();
        if ((b + 2) * CB + tid < n) id_next = a.point_list[range.x + (b + 2) * CB + tid];

        const int cnt = min(CB, n - b * CB);
        const Rec* rec = s_rec[buf];
        bool warp_done = !__any_sync(0xffffffffu, alive);
        if (!warp_done) {
            for (int c0 = 0; c0 < cnt; c0 += 32) {
                const int j = c0 + lane;
                bool pass = false;
                if (j < cnt) {
                    const float4 g0 = rec[j].a;
                    const float4 g1 = rec[j].b;
                    const float q = rect_min_q(g0.z, g0.w, g1.x, g0.x - rx1, g0.x - rx0,
                                               g0.y - ry1, g0.y - ry0);
                    pass = !(q > g1.w);
                }
                uint32_t mask = __ballot_sync(0xffffffffu, pass);
                const uint32_t idx_base = (uint32_t)(b * CB + c0 + 1);
                // Branch-free body: the kernel is issue-bound and most lanes of a
                // surviving instance do blend, so predication beats divergence.
                while (mask) {
                    const int bit = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const Rec& r = rec[c0 + bit];
                    const float4 g0 = r.a;
                    const float4 g1 = r.b;
                    const float4 col = r.c;
                    const float dx = g0.x - pixx, dy = g0.y - pixy;
                    const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                    const float alpha = min(0.99f, g1.y * expf(power));
                    const float test_T = T * (1 - alpha);
                    // forward.cu:339-354: skip power > 0, skip alpha < 1/255, stop at T < 1e-4
                    const bool ok = alive && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                    const bool stop = ok && (test_T < 0.0001f);
                    const bool blend = ok && !stop;
                    const float ae = blend ? alpha : 0.0f;
                    C0 += col.x * ae * T;
                    C1 += col.y * ae * T;
                    C2 += col.z * ae * T;
                    D += g1.z * ae * T;
                    T = blend ? test_T : T;
                    last_contributor = blend ? idx_base + (uint32_t)bit : last_contributor;
                    alive = alive && !stop;
                }
                if (!__any_sync(0xffffffffu, alive)) {
                    warp_done = true;
                    break;
                }
            }
        }
        cp_async_wait_all();
        if (__syncthreads_and(warp_done)) break;
    }

    if (inside) {
        const size_t HW = (size_t)a.H * a.W;
        a.final_T[pix_id] = T;
        a.n_contrib[pix_id] = last_contributor;
        a.out_color[0 * HW + pix_id] = C0 + T * a.bg[0];
        a.out_color[1 * HW + pix_id] = C1 + T * a.bg[1];
        a.out_color[2 * HW + pix_id] = C2 + T * a.bg[2];
        a.out_depth[pix_id] = D;
    }
}

struct RenderBwdArgs {
    const uint2* ranges;
    const uint32_t* point_list;
    int W, H;
    const float* bg;
    const float4* xyAB;
    const float4* Cod;
    const float4* rgb;
    const float* final_T;
    const uint32_t* n_contrib;
    const float* dL_dpix;        // [3,H,W]
    const float* dL_dpix_depth;  // [1,H,W]
    float* grad_rec;             // [P][GRAD_REC], zeroed
};

// 10 per-lane partials -> 10 warp sums, each left in one even lane:
// lanes {0,2,4,8,10} hold v0..v4, lanes {16,18,20,24,26} hold v5..v9.
__device__ __forceinline__ float warp_reduce10(const float (&v)[10], int lane) {
    const uint32_t F = 0xffffffffu;
    float r[5];
    const bool h4 = lane & 16;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float send = h4 ? v[i] : v[i + 5];
        const float keep = h4 ? v[i + 5] : v[i];
        r[i] = keep + __shfl_xor_sync(F, send, 16);
    }
    const bool h3 = lane & 8;
    float s0, s1, s2;
    {
        float send = h3 ? r[0] : r[3], keep = h3 ? r[3] : r[0];
        s0 = keep + __shfl_xor_sync(F, send, 8);
        send = h3 ? r[1] : r[4]; keep = h3 ? r[4] : r[1];
        s1 = keep + __shfl_xor_sync(F, send, 8);
        send = h3 ? r[2] : 0.f; keep = h3 ? 0.f : r[2];
        s2 = keep + __shfl_xor_sync(F, send, 8);
    }
    const bool h2 = lane & 4;
    float t0, t1;
    {
        float send = h2 ? s0 : s2, keep = h2 ? s2 : s0;
        t0 = keep + __shfl_xor_sync(F, send, 4);
        send = h2 ? s1 : 0.f; keep = h2 ? 0.f : s1;
        t1 = keep + __shfl_xor_sync(F, send, 4);
    }
    const bool h1 = lane & 2;
    float u;
    {
        const float send = h1 ? t0 : t1, keep = h1 ? t1 : t0;
        u = keep + __shfl_xor_sync(F, send, 2);
    }
    u += __shfl_xor_sync(F, u, 1);
    return u;
}
// slot (0..9) owned by `lane` after warp_reduce10, or -1
__device__ __forceinline__ int reduce10_slot(int lane) {
    if (lane & 1) return -1;
    const int code = (lane >> 1) & 7;   // (b3,b2,b1)
    int r;
    switch (code) {
        case 0: r = 0; break;
        case 1: r = 1; break;
        case 2: r = 2; break;
        case 4: r = 3; break;
        case 5: r = 4; break;
        default: return -1;
    }
    return r + ((lane & 16) ? 5 : 0);
}

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__global__ void __launch_bounds__(TILE_PIX) render_backward_kernel(RenderBwdArgs a) {
    __shared__ Rec s_rec[2][CB];
    __shared__ uint32_t s_id[2][CB];
    __shared__ uint32_t s_max;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int X0 = blockIdx.x * TILE_X + (warp & 1) * SUB_W;
    const int Y0 = blockIdx.y * TILE_Y + (warp >> 1) * SUB_H;
    const int px = X0 + (lane & 7), py = Y0 + (lane >> 3);
    const bool inside = px < a.W && py < a.H;
    const float pixx = (float)px, pixy = (float)py;
    const uint32_t pix_id = (uint32_t)a.W * py + px;
    const float rx0 = (float)X0, rx1 = (float)(X0 + SUB_W - 1);
    const float ry0 = (float)Y0, ry1 = (float)(Y0 + SUB_H - 1);
    const size_t HW = (size_t)a.H * a.W;

    const uint2 range = a.ranges[tile];

    const float T_final = inside ? a.final_T[pix_id] : 0.f;
    float T = T_final;
    const int last_contributor = inside ? (int)a.n_contrib[pix_id] : 0;
    float dpix0 = 0.f, dpix1 = 0.f, dpix2 = 0.f, dpixd = 0.f;
    if (inside) {
        dpix0 = a.dL_dpix[0 * HW + pix_id];
        dpix1 = a.dL_dpix[1 * HW + pix_id];
        dpix2 = a.dL_dpix[2 * HW + pix_id];
        dpixd = a.dL_dpix_depth[pix_id];
    }
    // -(T_final * sum_c bg_c dL/dpix_c): the background term of backward.cu:564-567
    const float nbg = -T_final * (a.bg[0] * dpix0 + a.bg[1] * dpix1 + a.bg[2] * dpix2);

    // only instances below the block's deepest contributor matter (backward.cu:513)
    if (tid == 0) s_max = 0;
    __syncthreads();
    const int warp_max = (int)__reduce_max_sync(0xffffffffu, (uint32_t)last_contributor);
    if (lane == 0 && warp_max > 0) atomicMax(&s_max, (uint32_t)warp_max);
    __syncthreads();
    const int m = (int)s_max;   // instances [0, m) are walked back to front
    const int nb = (m + CB - 1) / CB;
    if (nb == 0) return;

    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accd = 0.f;
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_depth = 0.f;
    const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;
    const int my_slot = reduce10_slot(lane);
    float* const my_grad = a.grad_rec + (my_slot >= 0 ? my_slot : 0);

    // slot t of batch b holds list index m-1-(b*CB+t)
    uint32_t id_next = 0;
    {
        const int i0 = m - 1 - tid;
        if (i0 >= 0) {
            const uint32_t id = a.point_list[range.x + i0];
            s_id[0][tid] = id;
            cp_async16(&s_rec[0][tid].a, &a.xyAB[id]);
            cp_async16(&s_rec[0][tid].b, &a.Cod[id]);
            cp_async16(&s_rec[0][tid].c, &a.rgb[id]);
        }
        cp_async_commit();
        const int i1 = m - 1 - (CB + tid);
        if (i1 >= 0) id_next = a.point_list[range.x + i1];
        cp_async_wait_all();
        __syncthreads();
    }

    for (int b = 0; b < nb; ++b) {
        const int buf = b & 1;
        {
            const int i1 = m - 1 - ((b + 1) * CB + tid);
            if (i1 >= 0) {
                s_id[buf ^ 1][tid] = id_next;
                cp_async16(&s_rec[buf ^ 1][tid].a, &a.xyAB[id_next]);
                cp_async16(&s_rec[buf ^ 1][tid].b, &a.Cod[id_next]);
                cp_async16(&s_rec[buf ^ 1][tid].c, &a.rgb[id_next]);
            }
            cp_async_commit();
            const int i2 = m - 1 - ((b + 2) * CB + tid);
            if (i2 >= 0) id_next = a.point_list[range.x + i2];
        }
        const int top = m - 1 - b * CB;           // list index of slot 0
        const int cnt = min(CB, top + 1);
        const Rec* rec = s_rec[buf];
        // the warp has nothing to do for instances at or above warp_max
        if (top - (cnt - 1) < warp_max) {
            for (int c0 = 0; c0 < cnt; c0 += 32) {
                const int j = c0 + lane;
                bool pass = false;
                if (j < cnt && top - j < warp_max) {
                    const float4 g0 = rec[j].a;
                    const float4 g1 = rec[j].b;
                    const float q = rect_min_q(g0.z, g0.w, g1.x, g0.x - rx1, g0.x - rx0,
                                               g0.y - ry1, g0.y - ry0);
                    pass = !(q > g1.w);
                }
                uint32_t mask = __ballot_sync(0xffffffffu, pass);
                while (mask) {
                    const int jj = c0 + __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int contributor = top - jj;   // 0-based list index
                    const Rec& r = rec[jj];
                    const float4 g0 = r.a;
                    const float4 g1 = r.b;
                    const float dx = g0.x - pixx, dy = g0.y - pixy;
                    const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                    const float G = expf(power);
                    const float alpha = min(0.99f, g1.y * G);
                    const bool contrib = contributor < last_contributor && !(power > 0.0f) &&
                                         !(alpha < 1.0f / 255.0f);
                    if (!__any_sync(0xffffffffu, contrib)) continue;
                    float v[10];
#pragma unroll
                    for (int i = 0; i < 10; ++i) v[i] = 0.f;
                    if (contrib) {
                        const float4 col = r.c;
                        // 1/(1-alpha): alpha <= 0.99, approx reciprocal (1 ulp) instead of
                        // the two IEEE divisions of backward.cu:529,567
                        const float inv = rcp_approx(1.f - alpha);
                        T = T * inv;
                        const float w = alpha * T;
                        const float om = 1.f - last_alpha;
                        acc0 = last_alpha * lc0 + om * acc0;
                        acc1 = last_alpha * lc1 + om * acc1;
                        acc2 = last_alpha * lc2 + om * acc2;
                        accd = last_alpha * last_depth + om * accd;
                        lc0 = col.x; lc1 = col.y; lc2 = col.z; last_depth = g1.z;
                        float dL_dalpha = (col.x - acc0) * dpix0 + (col.y - acc1) * dpix1 +
                                          (col.z - acc2) * dpix2 + (g1.z - accd) * dpixd;
                        v[6] = w * dpix0;
                        v[7] = w * dpix1;
                        v[8] = w * dpix2;
                        v[9] = w * dpixd;
                        dL_dalpha = dL_dalpha * T + nbg * inv;
                        last_alpha = alpha;
                        const float dL_dG = g1.y * dL_dalpha;
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = -gdx * g0.z - gdy * g0.w;
                        const float dG_ddely = -gdy * g1.x - gdx * g0.w;
                        const float h = -0.5f * dL_dG;
                        v[0] = dL_dG * dG_ddelx * ddelx_dx;
                        v[1] = dL_dG * dG_ddely * ddely_dy;
                        v[2] = h * gdx * dx;
                        v[3] = h * gdx * dy;
                        v[4] = h * gdy * dy;
                        v[5] = G * dL_dalpha;
                    }
                    const float sum = warp_reduce10(v, lane);
                    if (my_slot >= 0) atomicAdd(my_grad + (size_t)s_id[buf][jj] * GRAD_REC, sum);
                }
            }
        }
        cp_async_wait_all();
        __syncthreads();
    }
}

}  // namespace s3g
