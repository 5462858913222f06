// api.cu - extern "C" entry points of libs3g_b200.so (see include/s3g_b200.h).
//
// Host orchestration of the forward / backward passes; mirrors the control flow
// of CudaRasterizer::Rasterizer::{markVisible,forward,backward}
// (DGR/cuda_rasterizer/rasterizer_impl.cu:141-153,198-339,343-444) behind a flat
// C ABI.  No torch, no CPU fallback: every path ends in a kernel launch on the
// caller's stream or in an error code.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>

#include "../../include/s3g_b200.h"
#include "binning.cuh"
#include "common.cuh"
#include "composite.cuh"
#include "deform.cuh"
#include "deform_tc.cuh"
#include "preprocess.cuh"
#include "radix_sort.cuh"
#include "train_step.cuh"

using namespace s3g;

namespace {
thread_local std::string g_last_error;

int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    g_last_error = what;
    if (e != cudaSuccess) {
        g_last_error += ": ";
        g_last_error += cudaGetErrorString(e);
    }
    return code;
}

#define S3G_CUDA(call, what)                                        \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return fail(S3G_ERR_CUDA, what, e__); \
    } while (0)

// debug mode: synchronise + check after every stage (CHECK_CUDA, auxiliary.h:166-173)
#define S3G_STAGE(what)                                                            \
    do {                                                                           \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ == cudaSuccess && debug) e__ = cudaStreamSynchronize(stream);      \
        if (e__ != cudaSuccess) return fail(S3G_ERR_CUDA, what, e__);              \
    } while (0)

// ---- optional per-stage timing -------------------------------------------
constexpr int kMaxStages = 12;
struct StageTimer {
    bool on = false;
    int n[2] = {0, 0};
    cudaEvent_t ev[2][kMaxStages + 1] = {};
    const char* names[2][kMaxStages] = {};
    void mark(int which, cudaStream_t s, const char* name_of_stage_that_starts) {
        if (!on) return;
        int i = n[which];
        if (i > kMaxStages) return;
        if (!ev[which][i]) cudaEventCreate(&ev[which][i]);
        cudaEventRecord(ev[which][i], s);
        if (name_of_stage_that_starts && i < kMaxStages) names[which][i] = name_of_stage_that_starts;
        n[which] = i + 1;
    }
    void begin(int which) { n[which] = 0; }
};
StageTimer g_timer;
#define S3G_MARK(which, name) g_timer.mark(which, stream, name)

// number of key bits that cover every tile id (getHigherMsb, rasterizer_impl.cu:35-50)
int tile_key_bits(uint32_t tiles) {
    int b = 0;
    while (b < 32 && (tiles >> b) != 0) ++b;
    return b < 1 ? 1 : b;
}
}  // namespace

extern "C" {

int s3g_abi_version(void) { return S3G_ABI_VERSION; }
const char* s3g_last_error(void) { return g_last_error.c_str(); }
const char* s3g_build_arch(void) { return "sm_100a"; }

int s3g_profile_enable(int on) {
    g_timer.on = on != 0;
    return S3G_OK;
}
int s3g_profile_read(int which, float* ms, int capacity) {
    if (which < 0 || which > 1 || !ms) return fail(S3G_ERR_ARG, "profile_read: bad argument");
    int stages = g_timer.n[which] - 1;
    if (stages < 0) stages = 0;
    if (stages > capacity) stages = capacity;
    for (int i = 0; i < stages; ++i) {
        cudaError_t e = cudaEventSynchronize(g_timer.ev[which][i + 1]);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms[i], g_timer.ev[which][i], g_timer.ev[which][i + 1]);
        if (e != cudaSuccess) return fail(S3G_ERR_CUDA, "profile_read", e);
    }
    return stages;
}
const char* s3g_profile_stage_name(int which, int index) {
    if (which < 0 || which > 1 || index < 0 || index >= kMaxStages || !g_timer.names[which][index]) return "";
    return g_timer.names[which][index];
}

size_t s3g_geom_bytes(int64_t P) {
    size_t t = 0;
    GeomState::carve(nullptr, P, &t);
    return t;
}
size_t s3g_binning_bytes(int64_t R) {
    size_t t = 0;
    BinningState::carve(nullptr, R, &t);
    return t;
}
size_t s3g_image_bytes(int width, int height) {
    size_t t = 0;
    ImageState::carve(nullptr, width, height, &t);
    return t;
}
size_t s3g_sort_temp_bytes(int64_t n) {
    // look-back state + the two ping-pong arrays
    size_t m = (size_t)(n > 0 ? n : 1);
    return SortTemp::bytes(n) + 2 * (m * sizeof(uint32_t) + 128);
}

int s3g_state_field(int buffer, const char* name, int64_t P, int64_t R, int width, int height,
                    size_t* offset, size_t* elem_bytes, size_t* count) {
    if (!name || !offset || !elem_bytes || !count) return fail(S3G_ERR_ARG, "state_field: null argument");
    // carve on a fake 128-aligned base so that pointers are offsets + base
    char* base = reinterpret_cast<char*>((uintptr_t)1 << 40);
    auto set = [&](const void* p, size_t eb, size_t n) {
        *offset = (size_t)(reinterpret_cast<const char*>(p) - base);
        *elem_bytes = eb;
        *count = n;
        return S3G_OK;
    };
    const std::string f(name);
    if (buffer == 0) {
        GeomState g = GeomState::carve(base, P);
        if (f == "xyAB") return set(g.xyAB, 16, P);
        if (f == "Cod") return set(g.Cod, 16, P);
        if (f == "rgb") return set(g.rgb, 16, P);
        if (f == "depth_key") return set(g.depth_key, 4, P);
        if (f == "tiles_touched") return set(g.tiles_touched, 4, P);
        if (f == "rect") return set(g.rect, 8, P);
        if (f == "clamped") return set(g.clamped, 1, P);
        if (f == "order") return set(g.order_a, 4, P);
        if (f == "offsets") return set(g.offsets, 4, P);
        if (f == "grad_rec") return set(g.grad_rec, 4, (size_t)P * GRAD_REC);
    } else if (buffer == 1) {
        BinningState b = BinningState::carve(base, R);
        if (f == "point_list") return set(b.point_list, 4, R);
        if (f == "point_list_tiles") return set(b.point_list_tiles, 4, R);
    } else if (buffer == 2) {
        ImageState s = ImageState::carve(base, width, height);
        TileGrid tg = tile_grid(width, height);
        if (f == "final_T") return set(s.final_T, 4, (size_t)width * height);
        if (f == "n_contrib") return set(s.n_contrib, 4, (size_t)width * height);
        if (f == "ranges") return set(s.ranges, 8, (size_t)tg.count());
    }
    return fail(S3G_ERR_ARG, "state_field: unknown buffer/field");
}

int s3g_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (P < 0) return fail(S3G_ERR_ARG, "mark_visible: P < 0");
    if (P == 0) return S3G_OK;
    if (!means3D || !viewmatrix || !projmatrix || !present)
        return fail(S3G_ERR_ARG, "mark_visible: null pointer");
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, viewmatrix, projmatrix,
                                                             present);
    S3G_CUDA(cudaGetLastError(), "mark_visible launch");
    return S3G_OK;
}

int s3g_sort_pairs_u32(int64_t n, uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out,
                       uint32_t* vals_out, int begin_bit, int end_bit, void* temp, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n < 0 || n >= (1ll << 30)) return fail(S3G_ERR_ARG, "sort: n out of range");
    if (begin_bit < 0 || end_bit > 32 || end_bit <= begin_bit)
        return fail(S3G_ERR_ARG, "sort: bad bit range");
    if (n == 0) return S3G_OK;
    // scratch ping-pong lives in temp after the SortTemp block
    Carver c(static_cast<char*>(temp));
    SortTemp st = SortTemp::carve(c, n);
    uint32_t* kt = c.take<uint32_t>((size_t)n);
    uint32_t* vt = c.take<uint32_t>((size_t)n);
    S3G_CUDA(radix_sort_pairs((uint32_t)n, keys_in, vals_in, kt, vt, keys_out, vals_out, begin_bit,
                              end_bit, st, stream),
             "radix sort");
    return S3G_OK;
}

int64_t s3g_rasterize_forward(s3g_alloc_fn geom_alloc, void* geom_user, s3g_alloc_fn binning_alloc,
                              void* binning_user, s3g_alloc_fn image_alloc, void* image_user, int P,
                              int D, int M, const float* background, int width, int height,
                              const float* means3D, const float* shs, const float* colors_precomp,
                              const float* opacities, const float* scales, float scale_modifier,
                              const float* rotations, const float* cov3D_precomp,
                              const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                              float tan_fovx, float tan_fovy, int prefiltered, float* out_color,
                              float* out_depth, int* radii, int debug, void* stream_) {
    (void)prefiltered;   // the reference only uses it for a device-side trap (auxiliary.h:156-160)
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (P < 0 || width <= 0 || height <= 0) return fail(S3G_ERR_ARG, "forward: bad sizes");
    if (!geom_alloc || !binning_alloc || !image_alloc) return fail(S3G_ERR_ARG, "forward: null allocator");
    if (!out_color || !out_depth || !background) return fail(S3G_ERR_ARG, "forward: null output/background");
    const size_t HW = (size_t)width * height;
    if (P == 0) {   // rasterize_points.cu:82: outputs stay zero-filled
        S3G_CUDA(cudaMemsetAsync(out_color, 0, 3 * HW * sizeof(float), stream), "memset color");
        S3G_CUDA(cudaMemsetAsync(out_depth, 0, HW * sizeof(float), stream), "memset depth");
        return 0;
    }
    if (!means3D || !opacities || !viewmatrix || !projmatrix)
        return fail(S3G_ERR_ARG, "forward: null input");
    if (!colors_precomp && !shs)
        return fail(S3G_ERR_ARG, "forward: need SHs or precomputed colours");
    if (!colors_precomp && !cam_pos) return fail(S3G_ERR_ARG, "forward: SH path needs cam_pos");
    if (!cov3D_precomp && (!scales || !rotations))
        return fail(S3G_ERR_ARG, "forward: need scales+rotations or precomputed cov3D");
    if (!colors_precomp && (D < 0 || D > 3 || M < (D + 1) * (D + 1)))
        return fail(S3G_ERR_ARG, "forward: SH degree/coefficient count mismatch");
    if ((int64_t)P >= (1ll << 30)) return fail(S3G_ERR_ARG, "forward: P too large");

    const TileGrid tg = tile_grid(width, height);
    if (tg.x > 65535 || tg.y > 65535) return fail(S3G_ERR_ARG, "forward: image too large");

    char* gptr = geom_alloc(geom_user, s3g_geom_bytes(P));
    if (!gptr) return fail(S3G_ERR_ALLOC, "forward: geometry allocator returned NULL");
    GeomState geom = GeomState::carve(gptr, P);
    char* iptr = image_alloc(image_user, s3g_image_bytes(width, height));
    if (!iptr) return fail(S3G_ERR_ALLOC, "forward: image allocator returned NULL");
    ImageState img = ImageState::carve(iptr, width, height);
    if (!radii) radii = geom.internal_radii;

    // ---- per-Gaussian preprocess -----------------------------------------
    PreFwdArgs pa;
    pa.P = P; pa.D = D; pa.M = M;
    pa.means3D = means3D; pa.scales = scales; pa.scale_modifier = scale_modifier;
    pa.rotations = rotations; pa.opacities = opacities; pa.shs = shs;
    pa.cov3D_precomp = cov3D_precomp; pa.colors_precomp = colors_precomp;
    pa.view = viewmatrix; pa.proj = projmatrix; pa.campos = cam_pos;
    pa.W = width; pa.H = height;
    pa.tan_fovx = tan_fovx; pa.tan_fovy = tan_fovy;
    pa.focal_y = height / (2.0f * tan_fovy);   // rasterizer_impl.cu:223-224
    pa.focal_x = width / (2.0f * tan_fovx);
    pa.radii = radii; pa.xyAB = geom.xyAB; pa.Cod = geom.Cod; pa.rgb = geom.rgb;
    pa.depth_key = geom.depth_key; pa.tiles_touched = geom.tiles_touched; pa.rect = geom.rect;
    pa.clamped = geom.clamped; pa.order = geom.order_a;
    pa.grid_x = tg.x; pa.grid_y = tg.y;
    g_timer.begin(0);
    S3G_MARK(0, "preprocess_forward");
    {
        const bool use_sh = colors_precomp == nullptr;
        const size_t smem = use_sh ? (size_t)(PRE_THREADS / 32) * 32 * row_stride(3 * M) * sizeof(float) : 0;
        if (smem > 200 * 1024) return fail(S3G_ERR_ARG, "forward: too many SH coefficients");
        const int grid = (P + PRE_THREADS - 1) / PRE_THREADS;
        if (use_sh && M == 16) {
            preprocess_forward_kernel<16><<<grid, PRE_THREADS, smem, stream>>>(pa);
        } else {
            if (smem > 48 * 1024)
                S3G_CUDA(cudaFuncSetAttribute(preprocess_forward_kernel<0>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                         "smem attribute");
            preprocess_forward_kernel<0><<<grid, PRE_THREADS, smem, stream>>>(pa);
        }
    }
    S3G_STAGE("preprocess_forward");
    S3G_MARK(0, "depth_sort");

    // ---- depth digits of the LSD sort, over Gaussians --------------------
    S3G_CUDA(radix_sort_pairs((uint32_t)P, geom.depth_key, geom.order_a, geom.key_b, geom.order_b,
                              nullptr, geom.order_a, 0, 32, geom.sort, stream),
             "depth sort");
    S3G_STAGE("depth sort");

    S3G_MARK(0, "scan+readback");
    // ---- offsets in depth order + total ----------------------------------
    {
        const size_t zb = (size_t)(reinterpret_cast<char*>(geom.scan_misc + 32) -
                                   reinterpret_cast<char*>(geom.scan_status));
        S3G_CUDA(cudaMemsetAsync(geom.scan_status, 0, zb, stream), "memset scan");
        const uint32_t nblk = (uint32_t)div_up64(P, SCAN_TILE);
        scan_tiles_kernel<<<nblk, SCAN_THREADS, 0, stream>>>(geom.order_a, geom.tiles_touched,
                                                             (uint32_t)P, geom.offsets,
                                                             geom.scan_status, geom.scan_misc);
        S3G_STAGE("scan");
    }
    // rasterizer_impl.cu:281-282: the one read-back that sizes the binning arena.  Pinned
    // destination + event spin-wait: a blocking cudaStreamSynchronize() after a long queue
    // sleeps in the OS and wakes up milliseconds late, which idles the GPU.
    uint64_t total = 0;
    {
        static thread_local uint64_t* h_total = nullptr;
        static thread_local cudaEvent_t ev = nullptr;
        if (!h_total) {
            S3G_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_total), sizeof(uint64_t), cudaHostAllocDefault),
                     "pinned readback alloc");
            S3G_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "readback event");
        }
        S3G_CUDA(cudaMemcpyAsync(h_total, geom.scan_misc + 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream),
                 "num_rendered copy");
        S3G_CUDA(cudaEventRecord(ev, stream), "num_rendered event");
        cudaError_t q;
        while ((q = cudaEventQuery(ev)) == cudaErrorNotReady) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        if (q != cudaSuccess) return fail(S3G_ERR_CUDA, "num_rendered wait", q);
        total = *h_total;
    }
    if (total >= (1ull << 30)) return fail(S3G_ERR_ARG, "forward: more than 2^30 tile instances");
    const int64_t R = (int64_t)total;

    char* bptr = binning_alloc(binning_user, s3g_binning_bytes(R));
    if (!bptr) return fail(S3G_ERR_ALLOC, "forward: binning allocator returned NULL");
    BinningState bin = BinningState::carve(bptr, R);

    S3G_CUDA(cudaMemsetAsync(img.ranges, 0, (size_t)tg.count() * sizeof(uint2), stream),
             "memset ranges");
    S3G_MARK(0, "emit");
    if (R > 0) {
        emit_instances_kernel<<<(P + 255) / 256, 256, 0, stream>>>(
            (uint32_t)P, geom.order_a, geom.offsets, geom.tiles_touched, geom.rect, tg.x, bin.tile_a,
            bin.idx_a);
        S3G_STAGE("emit");
        S3G_MARK(0, "tile_sort");
        S3G_CUDA(radix_sort_pairs((uint32_t)R, bin.tile_a, bin.idx_a, bin.tile_b, bin.idx_b,
                                  bin.point_list_tiles, bin.point_list, 0,
                                  tile_key_bits((uint32_t)tg.count()), bin.sort, stream),
                 "tile sort");
        S3G_STAGE("tile sort");
        S3G_MARK(0, "tile_ranges");
        tile_ranges_kernel<<<(uint32_t)((R + 255) / 256), 256, 0, stream>>>(
            (uint32_t)R, bin.point_list_tiles, img.ranges);
        S3G_STAGE("ranges");
    }

    S3G_MARK(0, "render_forward");
    RenderFwdArgs ra;
    ra.ranges = img.ranges; ra.point_list = bin.point_list; ra.W = width; ra.H = height;
    ra.xyAB = geom.xyAB; ra.Cod = geom.Cod; ra.rgb = geom.rgb; ra.bg = background;
    ra.final_T = img.final_T; ra.n_contrib = img.n_contrib;
    ra.out_color = out_color; ra.out_depth = out_depth;
    render_forward_kernel<<<dim3(tg.x, tg.y), TILE_PIX, 0, stream>>>(ra);
    S3G_STAGE("render_forward");
    S3G_MARK(0, nullptr);
    return R;
}

int s3g_rasterize_backward(int P, int D, int M, int64_t R, const float* background, int width,
                           int height, const float* means3D, const float* shs,
                           const float* colors_precomp, const float* scales, float scale_modifier,
                           const float* rotations, const float* cov3D_precomp,
                           const float* viewmatrix, const float* projmatrix, const float* campos,
                           float tan_fovx, float tan_fovy, const int* radii, char* geom_buffer,
                           char* binning_buffer, char* image_buffer, const float* dL_dpix,
                           const float* dL_dpix_depth, float* dL_dmean2D, float* dL_dconic,
                           float* dL_dopacity, float* dL_dcolor, float* dL_ddepth, float* dL_dmean3D,
                           float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                           int debug, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (P < 0 || width <= 0 || height <= 0 || R < 0) return fail(S3G_ERR_ARG, "backward: bad sizes");
    if (P == 0) return S3G_OK;   // rasterize_points.cu:165
    if (!geom_buffer || !image_buffer || (R > 0 && !binning_buffer))
        return fail(S3G_ERR_STATE, "backward: missing state buffer");
    if (!means3D || !viewmatrix || !projmatrix || !background || !dL_dpix || !dL_dpix_depth)
        return fail(S3G_ERR_ARG, "backward: null input");
    if (!dL_dmean2D || !dL_dopacity || !dL_dcolor || !dL_dmean3D || !dL_dcov3D || !dL_dscale ||
        !dL_drot)
        return fail(S3G_ERR_ARG, "backward: null gradient output");
    const bool use_sh = (colors_precomp == nullptr);
    if (use_sh && (!shs || !dL_dsh || !campos)) return fail(S3G_ERR_ARG, "backward: SH path needs shs, dL_dsh, campos");
    if (!cov3D_precomp && (!scales || !rotations))
        return fail(S3G_ERR_ARG, "backward: need scales+rotations or precomputed cov3D");

    GeomState geom = GeomState::carve(geom_buffer, P);
    BinningState bin = BinningState::carve(binning_buffer, R);
    ImageState img = ImageState::carve(image_buffer, width, height);
    if (!radii) radii = geom.internal_radii;
    const TileGrid tg = tile_grid(width, height);

    g_timer.begin(1);
    S3G_MARK(1, "zero_grad_rec");
    S3G_CUDA(cudaMemsetAsync(geom.grad_rec, 0, (size_t)P * GRAD_REC * sizeof(float), stream),
             "memset grad_rec");
    S3G_MARK(1, "render_backward");
    if (R > 0) {
        RenderBwdArgs ra;
        ra.ranges = img.ranges; ra.point_list = bin.point_list; ra.W = width; ra.H = height;
        ra.bg = background; ra.xyAB = geom.xyAB; ra.Cod = geom.Cod; ra.rgb = geom.rgb;
        ra.final_T = img.final_T; ra.n_contrib = img.n_contrib;
        ra.dL_dpix = dL_dpix; ra.dL_dpix_depth = dL_dpix_depth; ra.grad_rec = geom.grad_rec;
        render_backward_kernel<<<dim3(tg.x, tg.y), TILE_PIX, 0, stream>>>(ra);
        S3G_STAGE("render_backward");
    }

    S3G_MARK(1, "preprocess_backward");
    PreBwdArgs pb;
    pb.P = P; pb.D = D; pb.M = use_sh ? M : 0;
    pb.means3D = means3D; pb.radii = radii;
    pb.shs = use_sh ? shs : nullptr;   // backward.cu:406
    pb.clamped = geom.clamped;
    pb.scales = cov3D_precomp ? nullptr : scales;   // backward.cu:410 (the reference keys on scales)
    pb.rotations = rotations; pb.scale_modifier = scale_modifier;
    pb.cov3D_precomp = cov3D_precomp;
    pb.view = viewmatrix; pb.proj = projmatrix; pb.campos = campos;
    pb.focal_y = height / (2.0f * tan_fovy);
    pb.focal_x = width / (2.0f * tan_fovx);
    pb.tan_fovx = tan_fovx; pb.tan_fovy = tan_fovy;
    pb.grad_rec = geom.grad_rec;
    pb.dL_dmean2D = dL_dmean2D; pb.dL_dconic = dL_dconic; pb.dL_dopacity = dL_dopacity;
    pb.dL_dcolor = dL_dcolor; pb.dL_ddepth = dL_ddepth; pb.dL_dmean3D = dL_dmean3D;
    pb.dL_dcov3D = dL_dcov3D; pb.dL_dsh = use_sh ? dL_dsh : nullptr;
    pb.dL_dscale = dL_dscale; pb.dL_drot = dL_drot;
    {
        const size_t smem = use_sh ? (size_t)(PRE_THREADS / 32) * 32 * row_stride(3 * M) * sizeof(float) : 0;
        if (smem > 200 * 1024) return fail(S3G_ERR_ARG, "backward: too many SH coefficients");
        const int grid = (P + PRE_THREADS - 1) / PRE_THREADS;
        if (use_sh && M == 16) {
            preprocess_backward_kernel<16><<<grid, PRE_THREADS, smem, stream>>>(pb);
        } else {
            if (smem > 48 * 1024)
                S3G_CUDA(cudaFuncSetAttribute(preprocess_backward_kernel<0>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                         "smem attribute");
            preprocess_backward_kernel<0><<<grid, PRE_THREADS, smem, stream>>>(pb);
        }
    }
    S3G_STAGE("preprocess_backward");
    S3G_MARK(1, nullptr);
    return S3G_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// HexPlane + decoder
// ---------------------------------------------------------------------------
namespace {
int to_dnet(const s3g_deform_net* n, DNet& d) {
    if (!n) return fail(S3G_ERR_ARG, "deform: null net");
    if (n->feat_dim != FD) return fail(S3G_ERR_UNSUPPORTED, "deform: output_coordinate_dim must be 32");
    if (n->width != HWID) return fail(S3G_ERR_UNSUPPORTED, "deform: net_width must be 64");
    if (!(n->num_levels == 1 || n->num_levels == 2 || n->num_levels == 3 || n->num_levels == 4 ||
          n->num_levels == 8))
        return fail(S3G_ERR_UNSUPPORTED, "deform: number of HexPlane levels must be 1, 2, 3, 4 or 8");
    d.L = n->num_levels;
    for (int l = 0; l < d.L; ++l) {
        for (int c = 0; c < 4; ++c) {
            d.reso[l][c] = n->reso[l][c];
            if (d.reso[l][c] < 2) return fail(S3G_ERR_ARG, "deform: plane resolution < 2");
        }
        for (int k = 0; k < 6; ++k) {
            d.planes[l][k] = n->planes[l][k];
            if (!d.planes[l][k]) return fail(S3G_ERR_ARG, "deform: null plane");
        }
    }
    for (int c = 0; c < 3; ++c) {
        d.aabb0[c] = n->aabb[c];
        d.inv_span2[c] = 2.0f / (n->aabb[3 + c] - n->aabb[c]);   // hexplane.py:19-20
    }
    if (!n->w_feat || !n->b_feat) return fail(S3G_ERR_ARG, "deform: null feature_out");
    d.w_feat = n->w_feat; d.b_feat = n->b_feat;
    d.pos = {n->w_pos1, n->b_pos1, n->w_pos2, n->b_pos2};
    d.scl = {n->w_scl1, n->b_scl1, n->w_scl2, n->b_scl2};
    d.rot = {n->w_rot1, n->b_rot1, n->w_rot2, n->b_rot2};
    d.opa = {n->w_opa1, n->b_opa1, n->w_opa2, n->b_opa2};
    d.shs = {n->w_shs1, n->b_shs1, n->w_shs2, n->b_shs2};
    d.w_d0 = n->w_dino0; d.b_d0 = n->b_dino0; d.w_d2 = n->w_dino2; d.b_d2 = n->b_dino2;
    d.w_d4 = n->w_dino4; d.b_d4 = n->b_dino4;
    const Head2* hs[5] = {&d.pos, &d.scl, &d.rot, &d.opa, &d.shs};
    for (const Head2* h : hs)
        if (h->w1 && !(h->b1 && h->w2 && h->b2)) return fail(S3G_ERR_ARG, "deform: incomplete head");
    if (d.w_d0 && !(d.b_d0 && d.w_d2 && d.b_d2 && d.w_d4 && d.b_d4))
        return fail(S3G_ERR_ARG, "deform: incomplete dino head");
    return S3G_OK;
}
// weight matrices in the order one tile consumes them (see WPipe)
void build_wseq(const DNet& d, bool backward, WSeq& q) {
    q.count = 0;
    auto add = [&](const float* W, int N, int K) { q.W[q.count] = W; q.N[q.count] = (short)N; q.K[q.count] = (short)K; ++q.count; };
    const int KF = FD * d.L;
    add(d.w_feat, 64, KF);
    const Head2* small[4] = {&d.pos, &d.scl, &d.rot, &d.opa};
    for (const Head2* h : small)
        if (h->w1) { add(h->w1, 64, 64); if (backward) add(h->w1, 64, 64); }
    if (d.shs.w1) {
        add(d.shs.w1, 64, 64); add(d.shs.w2, 48, 64);
        if (backward) { add(d.shs.w2, 48, 64); add(d.shs.w1, 64, 64); }
    }
    if (d.w_d0) {
        add(d.w_d0, 64, 64); add(d.w_d2, 64, 64);
        if (backward) { add(d.w_d2, 64, 64); add(d.w_d0, 64, 64); }
    }
    if (backward) add(d.w_feat, 64, KF);
}
// prepared-weight table of the tcgen05 decoder
void build_tc_table(const DNet& d, TcTable& t, TcPrepArgs* prep) {
    int off = 0;
    auto put = [&](int id, const float* W, int N, int K) {
        if (!W) { t.off[id] = -1; t.npad[id] = 0; t.k[id] = 0; if (prep) { prep->W[id] = nullptr; prep->n[id] = 0; } return; }
        const int np = (N + 15) & ~15;
        t.off[id] = off; t.npad[id] = np; t.k[id] = K;
        off += 2 * np * K;
        if (prep) { prep->W[id] = W; prep->n[id] = N; }
    };
    put(TL_FEAT, d.w_feat, 64, FD * d.L);
    put(TL_POS1, d.pos.w1, 64, 64); put(TL_POS2, d.pos.w1 ? d.pos.w2 : nullptr, 3, 64);
    put(TL_SCL1, d.scl.w1, 64, 64); put(TL_SCL2, d.scl.w1 ? d.scl.w2 : nullptr, 3, 64);
    put(TL_ROT1, d.rot.w1, 64, 64); put(TL_ROT2, d.rot.w1 ? d.rot.w2 : nullptr, 4, 64);
    put(TL_OPA1, d.opa.w1, 64, 64); put(TL_OPA2, d.opa.w1 ? d.opa.w2 : nullptr, 1, 64);
    put(TL_SHS1, d.shs.w1, 64, 64); put(TL_SHS2, d.shs.w1 ? d.shs.w2 : nullptr, 48, 64);
    put(TL_D0, d.w_d0, 64, 64); put(TL_D2, d.w_d0 ? d.w_d2 : nullptr, 64, 64); put(TL_D4, d.w_d0 ? d.w_d4 : nullptr, 3, 64);
    t.total = off;
}
int deform_grid(int ntiles) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int g = sms;      // one persistent CTA per SM (shared memory: weights double-buffered)
    return ntiles < g ? ntiles : g;
}
}  // namespace

extern "C" {

int s3g_deform_forward(const s3g_deform_net* net, int P, const float* xyz, const float* scales,
                       const float* rotations, const float* opacity, const float* shs, float time,
                       const float* campos, int sh_degree, float* means3D, float* scales_act,
                       float* rot_act, float* opacity_act, float* colors, float* dx, float* dshs,
                       float* feat, float* features, void* workspace, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (P < 0) return fail(S3G_ERR_ARG, "deform_forward: P < 0");
    if (P == 0) return S3G_OK;
    DeformFwdArgs a;
    int rc = to_dnet(net, a.net);
    if (rc != S3G_OK) return rc;
    if (!xyz || !scales || !rotations || !opacity || !shs || !campos)
        return fail(S3G_ERR_ARG, "deform_forward: null input");
    if (!means3D || !scales_act || !rot_act || !opacity_act || !colors || !features)
        return fail(S3G_ERR_ARG, "deform_forward: null output");
    if (sh_degree < 0 || sh_degree > 3) return fail(S3G_ERR_ARG, "deform_forward: sh_degree must be 0..3");
    a.P = P; a.xyz = xyz; a.scales = scales; a.rot = rotations; a.opacity = opacity; a.shs = shs;
    a.campos = campos; a.time = time; a.sh_degree = sh_degree;
    a.o_means = means3D; a.o_scales = scales_act; a.o_rot = rot_act; a.o_opacity = opacity_act;
    a.o_colors = colors; a.o_dx = dx; a.o_dshs = dshs; a.o_feat = feat; a.features = features;
    build_wseq(a.net, false, a.wseq);
    {
        SampleArgs sa;
        sa.net = a.net; sa.P = P; sa.xyz = xyz; sa.time = time; sa.features = features;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int blocks = std::min((P + 7) / 8, sms * 12);
        if (a.net.L == 4) hexplane_sample_kernel<4><<<blocks, 256, 0, stream>>>(sa);
        else hexplane_sample_kernel<0><<<blocks, 256, 0, stream>>>(sa);
        S3G_CUDA(cudaGetLastError(), "hexplane_sample launch");
    }
    if (a.net.L <= 4) {
        // ---- decoder on the 5th-gen tensor cores (tcgen05 / TMEM) ------------------------------
        if (!workspace) return fail(S3G_ERR_ARG, "deform_forward: null workspace");
        DeformTcArgs t;
        t.net = a.net; t.P = P; t.xyz = xyz; t.scales = scales; t.rot = rotations; t.opacity = opacity; t.shs = shs;
        t.campos = campos; t.sh_degree = sh_degree;
        t.o_means = means3D; t.o_scales = scales_act; t.o_rot = rot_act; t.o_opacity = opacity_act; t.o_colors = colors;
        t.o_dx = dx; t.o_dshs = dshs; t.o_feat = feat; t.features = features;
        TcPrepArgs pp;
        build_tc_table(a.net, t.tab, &pp);
        pp.tab = t.tab;
        pp.dst = reinterpret_cast<float*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
        t.wprep = pp.dst;
        tc_prep_weights_kernel<<<dim3(8, TL_COUNT), 256, 0, stream>>>(pp);
        S3G_CUDA(cudaGetLastError(), "tc_prep_weights launch");
        const size_t smem = (size_t)(2 * TCM * 128 + 2 * 64 * 128) * sizeof(float);
        S3G_CUDA(cudaFuncSetAttribute(deform_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "deform tc smem attr");
        const int ntiles = (P + TCM - 1) / TCM;
        deform_forward_tc_kernel<<<deform_grid(ntiles), TCM, smem, stream>>>(t);
    } else {
        const size_t smem = DeformSmem::floats(a.net.L) * sizeof(float);
        const int ntiles = (P + DT - 1) / DT;
        S3G_CUDA(cudaFuncSetAttribute(deform_forward_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "deform smem attr");
        deform_forward_kernel<0><<<deform_grid(ntiles), DTHREADS, smem, stream>>>(a);
    }
    S3G_CUDA(cudaGetLastError(), "deform_forward launch");
    return S3G_OK;
}


namespace {
// layout of one CTA's partial-gradient buffer
int make_offsets(const DNet& d, GradOff& o) {
    int t = 0;
    auto take = [&](int n) { int r = t; t += (n + 3) & ~3; return r; };
    auto head = [&](const Head2& h, int k, int (&dst)[4]) {
        if (h.w1) { dst[0] = take(64 * 64); dst[1] = take(64); dst[2] = take(k * 64); dst[3] = take(k); }
        else { dst[0] = dst[1] = dst[2] = dst[3] = -1; }
    };
    o.w_feat = take(64 * FD * d.L); o.b_feat = take(64);
    head(d.pos, 3, o.pos); head(d.scl, 3, o.scl); head(d.rot, 4, o.rot); head(d.opa, 1, o.opa); head(d.shs, 48, o.shs);
    if (d.w_d0) { o.d0w = take(4096); o.d0b = take(64); o.d2w = take(4096); o.d2b = take(64); o.d4w = take(192); o.d4b = take(3); }
    else { o.d0w = o.d0b = o.d2w = o.d2b = o.d4w = o.d4b = -1; }
    o.total = t;
    return t;
}
int bwd_grid(int ntiles) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return ntiles < sms ? ntiles : sms;
}
constexpr int kMaxBwdGrid = 256;
}  // namespace

size_t s3g_deform_forward_workspace_bytes(const s3g_deform_net* net) {
    DNet d;
    if (to_dnet(net, d) != S3G_OK) return 0;
    TcTable t;
    build_tc_table(d, t, nullptr);
    return (size_t)t.total * sizeof(float) + 512;
}

size_t s3g_deform_workspace_bytes(const s3g_deform_net* net, int P) {
    DNet d;
    if (to_dnet(net, d) != S3G_OK) return 0;
    GradOff o;
    make_offsets(d, o);
    // per-CTA partial Linear gradients + dL/d(features) [P][32L]
    return (size_t)kMaxBwdGrid * o.total * sizeof(float) + 512 + (size_t)(P > 0 ? P : 0) * FD * d.L * sizeof(float);
}

int s3g_deform_backward(const s3g_deform_net* net, int P, const float* xyz, const float* scales,
                        const float* rotations, const float* opacity, const float* shs, float time,
                        const float* campos, int sh_degree, const float* features,
                        const float* g_means3D, const float* g_scales_act, const float* g_rot_act, const float* g_opacity_act,
                        const float* g_colors, const float* g_dx, const float* g_dshs, const float* g_feat,
                        float* d_xyz, float* d_scales, float* d_rotations, float* d_opacity, float* d_shs,
                        const s3g_deform_net_grads* grads, void* workspace, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (P < 0) return fail(S3G_ERR_ARG, "deform_backward: P < 0");
    DeformBwdArgs a;
    int rc = to_dnet(net, a.net);
    if (rc != S3G_OK) return rc;
    if (!grads || !workspace) return fail(S3G_ERR_ARG, "deform_backward: null grads/workspace");
    if (P > 0 && (!xyz || !scales || !rotations || !opacity || !shs || !campos || !features))
        return fail(S3G_ERR_ARG, "deform_backward: null input");
    if (P > 0 && (!d_xyz || !d_scales || !d_rotations || !d_opacity || !d_shs))
        return fail(S3G_ERR_ARG, "deform_backward: null output");
    if (sh_degree < 0 || sh_degree > 3) return fail(S3G_ERR_ARG, "deform_backward: sh_degree must be 0..3");
    const DNet& d = a.net;
    a.P = P; a.xyz = xyz; a.scales = scales; a.rot = rotations; a.opacity = opacity; a.shs = shs;
    a.campos = campos; a.time = time; a.sh_degree = sh_degree;
    a.g_means = g_means3D; a.g_scales = g_scales_act; a.g_rot = g_rot_act; a.g_opacity = g_opacity_act;
    a.g_colors = g_colors; a.g_dx = g_dx; a.g_dshs = g_dshs; a.g_feat = g_feat;
    a.d_xyz = d_xyz; a.d_scales = d_scales; a.d_rot = d_rotations; a.d_opacity = d_opacity; a.d_shs = d_shs;
    for (int l = 0; l < d.L; ++l)
        for (int k = 0; k < 6; ++k) {
            a.gplanes[l][k] = grads->planes[l][k];
            if (!a.gplanes[l][k]) return fail(S3G_ERR_ARG, "deform_backward: null plane gradient");
        }
    make_offsets(d, a.off);
    build_wseq(d, true, a.wseq);
    a.partial = reinterpret_cast<float*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    a.features = features;
    a.dfeatures = a.partial + (size_t)kMaxBwdGrid * a.off.total;
    const int ntiles = (P + DT - 1) / DT;
    int grid = bwd_grid(ntiles);
    if (grid > kMaxBwdGrid) grid = kMaxBwdGrid;
    // destination table of the reduction
    ReduceArgs r;
    r.nseg = 0; r.partial = a.partial; r.stride = a.off.total; r.nparts = grid > 0 ? grid : 0;
    auto seg = [&](float* dst, int off, int count) -> bool {
        if (off < 0) return true;
        if (!dst) return false;
        r.seg[r.nseg++] = ReduceSeg{dst, off, count};
        return true;
    };
    bool ok = seg(grads->w_feat, a.off.w_feat, 64 * FD * d.L) && seg(grads->b_feat, a.off.b_feat, 64);
    auto hseg = [&](const int (&o)[4], float* w1, float* b1, float* w2, float* b2, int k) {
        return seg(w1, o[0], 4096) && seg(b1, o[1], 64) && seg(w2, o[2], k * 64) && seg(b2, o[3], k);
    };
    ok = ok && hseg(a.off.pos, grads->w_pos1, grads->b_pos1, grads->w_pos2, grads->b_pos2, 3);
    ok = ok && hseg(a.off.scl, grads->w_scl1, grads->b_scl1, grads->w_scl2, grads->b_scl2, 3);
    ok = ok && hseg(a.off.rot, grads->w_rot1, grads->b_rot1, grads->w_rot2, grads->b_rot2, 4);
    ok = ok && hseg(a.off.opa, grads->w_opa1, grads->b_opa1, grads->w_opa2, grads->b_opa2, 1);
    ok = ok && hseg(a.off.shs, grads->w_shs1, grads->b_shs1, grads->w_shs2, grads->b_shs2, 48);
    ok = ok && seg(grads->w_dino0, a.off.d0w, 4096) && seg(grads->b_dino0, a.off.d0b, 64) &&
         seg(grads->w_dino2, a.off.d2w, 4096) && seg(grads->b_dino2, a.off.d2b, 64) &&
         seg(grads->w_dino4, a.off.d4w, 192) && seg(grads->b_dino4, a.off.d4b, 3);
    if (!ok) return fail(S3G_ERR_ARG, "deform_backward: null Linear gradient for an enabled layer");
    if (P > 0) {
        const size_t smem = DeformBwdSmem::floats(d.L) * sizeof(float);
        if (smem > 227 * 1024) return fail(S3G_ERR_UNSUPPORTED, "deform_backward: too many levels for shared memory");
        if (d.L == 4) {
            S3G_CUDA(cudaFuncSetAttribute(deform_backward_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "deform bwd smem attr");
            deform_backward_kernel<4><<<grid, DTHREADS, smem, stream>>>(a);
        } else {
            S3G_CUDA(cudaFuncSetAttribute(deform_backward_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "deform bwd smem attr");
            deform_backward_kernel<0><<<grid, DTHREADS, smem, stream>>>(a);
        }
        S3G_CUDA(cudaGetLastError(), "deform_backward launch");
        ScatterArgs sc;
        sc.net = a.net; sc.P = P; sc.xyz = xyz; sc.time = time; sc.dfeatures = a.dfeatures; sc.d_xyz = d_xyz;
        for (int l = 0; l < S3G_MAX_LEVELS; ++l)
            for (int k = 0; k < 6; ++k) sc.gplanes[l][k] = l < d.L ? a.gplanes[l][k] : nullptr;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int blocks = std::min((P + 7) / 8, sms * 8);
        if (d.L == 4) hexplane_scatter_kernel<4><<<blocks, 256, 0, stream>>>(sc);
        else hexplane_scatter_kernel<0><<<blocks, 256, 0, stream>>>(sc);
        S3G_CUDA(cudaGetLastError(), "hexplane_scatter launch");
    }
    {
        int maxc = 1;
        for (int i = 0; i < r.nseg; ++i) maxc = std::max(maxc, r.seg[i].count);
        deform_reduce_kernel<<<dim3((maxc + 255) / 256, r.nseg), 256, 0, stream>>>(r);
    }
    S3G_CUDA(cudaGetLastError(), "deform_reduce launch");
    return S3G_OK;
}

// ---- training-step kernels (train_step.cuh) --------------------------------------------------
int s3g_adam_step(int n, const s3g_adam_tensor* tensors, double beta1, double beta2, double eps, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n < 0 || (n > 0 && !tensors)) return fail(S3G_ERR_ARG, "adam_step: bad tensor table");
    AdamArgs a;
    a.w1 = (float)(1.0 - beta1);
    a.beta2 = (float)beta2;
    a.omb2 = (float)(1.0 - beta2);
    a.eps = (float)eps;
    int i = 0;
    while (i < n) {
        a.count = 0;
        int blocks = 0;
        for (; i < n && a.count < ADAM_MAX_TENSORS; ++i) {
            const s3g_adam_tensor& t = tensors[i];
            if (t.numel == 0) continue;
            if (t.numel < 0 || !t.param || !t.grad || !t.exp_avg || !t.exp_avg_sq || t.step < 1)
                return fail(S3G_ERR_ARG, "adam_step: null pointer, negative numel or step < 1");
            const long long nb = (t.numel + ADAM_CHUNK - 1) / ADAM_CHUNK;
            if (nb + blocks > 0x7fffffffLL) return fail(S3G_ERR_ARG, "adam_step: tensor too large");
            AdamTensor& d = a.t[a.count];
            d.p = t.param; d.g = t.grad; d.m = t.exp_avg; d.v = t.exp_avg_sq; d.n = t.numel;
            const double bc1 = 1.0 - std::pow(beta1, (double)t.step);
            const double bc2 = 1.0 - std::pow(beta2, (double)t.step);
            d.step_size = (float)(t.lr / bc1);
            d.bc2_sqrt = (float)std::sqrt(bc2);
            a.block_start[a.count] = blocks;
            blocks += (int)nb;
            ++a.count;
        }
        a.block_start[a.count] = blocks;
        if (blocks > 0) {
            adam_multi_tensor_kernel<<<blocks, ADAM_THREADS, 0, stream>>>(a);
            S3G_CUDA(cudaGetLastError(), "adam_step launch");
        }
    }
    return S3G_OK;
}

int s3g_densify_stats(int P, const float* viewspace_grad, const int* radii, float* xyz_gradient_accum,
                      float* denom, float* max_radii2D, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (P < 0) return fail(S3G_ERR_ARG, "densify_stats: P < 0");
    if (P == 0) return S3G_OK;
    if (!viewspace_grad || !radii || !xyz_gradient_accum || !denom || !max_radii2D)
        return fail(S3G_ERR_ARG, "densify_stats: null pointer");
    densify_stats_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, viewspace_grad, radii, xyz_gradient_accum, denom, max_radii2D);
    S3G_CUDA(cudaGetLastError(), "densify_stats launch");
    return S3G_OK;
}

namespace {
constexpr int kDepthBlocks = 592;    // 4 per SM
struct LossPlan {
    size_t map_floats, img_blocks;
    dim3 grid;
};
LossPlan loss_plan(int B, int C, int H, int W) {
    LossPlan p;
    p.map_floats = (size_t)B * C * H * W;
    p.grid = dim3((W + LOSS_T - 1) / LOSS_T, (H + LOSS_T - 1) / LOSS_T, B * C);
    p.img_blocks = (size_t)p.grid.x * p.grid.y * p.grid.z;
    return p;
}
}  // namespace

size_t s3g_image_loss_workspace_bytes(int B, int C, int H, int W) {
    if (B <= 0 || C < 0 || H <= 0 || W <= 0) return 0;
    const LossPlan p = loss_plan(B, C, H, W);
    return 256 + sizeof(float) * (3 * p.map_floats + 2 * p.img_blocks + 2 * kDepthBlocks);
}

namespace {
struct LossBufs { float *m0, *m1, *m2, *part_img, *part_dep; };
LossBufs loss_bufs(const LossPlan& p, const void* workspace) {
    LossBufs b;
    float* ws = reinterpret_cast<float*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    b.m0 = ws; b.m1 = b.m0 + p.map_floats; b.m2 = b.m1 + p.map_floats;
    b.part_img = b.m2 + p.map_floats; b.part_dep = b.part_img + 2 * p.img_blocks;
    return b;
}
LossWin loss_window() {
    // gaussian(11, 1.5) of loss_utils.py:56-58: float32 exp values normalised by their float32 sum
    LossWin win;
    float g[11], s = 0.f;
    for (int x = 0; x < 11; ++x) g[x] = (float)std::exp(-(double)((x - 5) * (x - 5)) / (2.0 * 1.5 * 1.5));
    for (int x = 0; x < 11; ++x) s += g[x];
    for (int x = 0; x < 11; ++x) win.w[x] = g[x] / s;
    return win;
}
int loss_check(int B, int C, int H, int W, const void* image, const void* gt, const void* depth, const void* gt_depth) {
    if (B <= 0 || C < 0 || H <= 0 || W <= 0) return fail(S3G_ERR_ARG, "image_loss: empty image");
    if ((long long)B * C > 65535) return fail(S3G_ERR_ARG, "image_loss: too many image planes");
    if (C > 0 && (!image || !gt)) return fail(S3G_ERR_ARG, "image_loss: null image pointer");
    if (C == 0 && !depth) return fail(S3G_ERR_ARG, "image_loss: neither image nor depth");
    if ((depth == nullptr) != (gt_depth == nullptr)) return fail(S3G_ERR_ARG, "image_loss: depth and gt_depth go together");
    return S3G_OK;
}
}  // namespace

int s3g_image_loss_forward(int B, int C, int H, int W, const float* image, const float* gt_image, const float* depth,
                           const float* gt_depth, float max_depth, double* sums, void* workspace, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = loss_check(B, C, H, W, image, gt_image, depth, gt_depth)) return rc;
    if (!sums || !workspace) return fail(S3G_ERR_ARG, "image_loss_forward: null sums/workspace");
    const LossPlan p = loss_plan(B, C, H, W);
    const LossBufs b = loss_bufs(p, workspace);
    if (C > 0) {
        loss_stats_kernel<<<p.grid, LOSS_THREADS, 0, stream>>>(H, W, image, gt_image, loss_window(), b.m0, b.m1, b.m2, b.part_img);
        S3G_CUDA(cudaGetLastError(), "loss_stats launch");
    }
    int ndep = 0;
    if (depth) {
        ndep = kDepthBlocks;
        loss_depth_stats_kernel<<<kDepthBlocks, LOSS_THREADS, 0, stream>>>((size_t)B * H * W, depth, gt_depth, max_depth, b.part_dep);
        S3G_CUDA(cudaGetLastError(), "loss_depth_stats launch");
    }
    loss_reduce_kernel<<<1, 256, 0, stream>>>((int)p.img_blocks, b.part_img, ndep, b.part_dep, sums);
    S3G_CUDA(cudaGetLastError(), "loss_reduce launch");
    return S3G_OK;
}

int s3g_image_loss_backward(int B, int C, int H, int W, const float* image, const float* gt_image, const float* depth,
                            const float* gt_depth, float max_depth, const float* weights, const double* sums,
                            const void* workspace, float* g_image, float* g_depth, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (int rc = loss_check(B, C, H, W, image, gt_image, depth, gt_depth)) return rc;
    if (!weights || !sums || !workspace || (C > 0 && !g_image) || (depth && !g_depth))
        return fail(S3G_ERR_ARG, "image_loss_backward: null pointer");
    const LossPlan p = loss_plan(B, C, H, W);
    const LossBufs b = loss_bufs(p, workspace);
    if (C > 0) {
        loss_grad_kernel<<<p.grid, LOSS_THREADS, 0, stream>>>(H, W, image, gt_image, loss_window(), b.m0, b.m1, b.m2, weights,
                                                             1.0f / (float)p.map_floats, g_image);
        S3G_CUDA(cudaGetLastError(), "loss_grad launch");
    }
    if (depth) {
        const size_t nd = (size_t)B * H * W;
        loss_depth_grad_kernel<<<(unsigned)((nd + 255) / 256), 256, 0, stream>>>(nd, depth, gt_depth, max_depth, weights, sums, g_depth);
        S3G_CUDA(cudaGetLastError(), "loss_depth_grad launch");
    }
    return S3G_OK;
}

// ---- HexPlane regularisers -------------------------------------------------------------------
namespace {
int reg_table(int n, const s3g_plane_desc* planes, bool need_grad, RegArgs& a) {
    if (n <= 0 || n > REG_MAX_PLANES || !planes) return fail(S3G_ERR_ARG, "plane_reg: 1..48 planes expected");
    int blocks = 0;
    a.count = n;
    for (int i = 0; i < n; ++i) {
        const s3g_plane_desc& d = planes[i];
        if (!d.plane || (need_grad && !d.grad)) return fail(S3G_ERR_ARG, "plane_reg: null plane / grad pointer");
        if (d.H < 3 || d.W < 1 || d.C < 4 || d.C % 4) return fail(S3G_ERR_ARG, "plane_reg: need H >= 3, W >= 1, C % 4 == 0");
        RegPlane& p = a.p[i];
        p.t = d.plane; p.g = d.grad; p.H = d.H; p.W = d.W; p.C = d.C;
        p.k_smooth = (float)((double)d.w_smooth / ((double)d.C * (d.H - 2) * d.W));
        p.k_l1 = (float)((double)d.w_l1 / ((double)d.C * d.H * d.W));
        const long long n4 = (long long)d.H * d.W * d.C / 4;
        a.block_start[i] = blocks;
        const long long nb = (n4 + REG_CHUNK4 - 1) / REG_CHUNK4;
        if (blocks + nb > 0x7fffffffLL) return fail(S3G_ERR_ARG, "plane_reg: planes too large");
        blocks += (int)nb;
    }
    a.block_start[n] = blocks;
    return blocks;
}
}  // namespace

size_t s3g_plane_reg_workspace_bytes(int n, const s3g_plane_desc* planes) {
    RegArgs a;
    const int blocks = reg_table(n, planes, false, a);
    return blocks < 0 ? 0 : 256 + sizeof(double) * (size_t)blocks;
}

int s3g_plane_reg_forward(int n, const s3g_plane_desc* planes, double* total, void* workspace, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    RegArgs a;
    const int blocks = reg_table(n, planes, false, a);
    if (blocks < 0) return blocks;
    if (!total || !workspace) return fail(S3G_ERR_ARG, "plane_reg_forward: null total/workspace");
    double* partial = reinterpret_cast<double*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    plane_reg_forward_kernel<<<blocks, REG_THREADS, 0, stream>>>(a, partial);
    S3G_CUDA(cudaGetLastError(), "plane_reg_forward launch");
    plane_reg_reduce_kernel<<<1, 256, 0, stream>>>(blocks, partial, total);
    S3G_CUDA(cudaGetLastError(), "plane_reg_reduce launch");
    return S3G_OK;
}

int s3g_plane_reg_backward(int n, const s3g_plane_desc* planes, const float* gscale, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    RegArgs a;
    const int blocks = reg_table(n, planes, true, a);
    if (blocks < 0) return blocks;
    if (!gscale) return fail(S3G_ERR_ARG, "plane_reg_backward: null gscale");
    plane_reg_backward_kernel<<<blocks, REG_THREADS, 0, stream>>>(a, gscale);
    S3G_CUDA(cudaGetLastError(), "plane_reg_backward launch");
    return S3G_OK;
}

// ---- densify / prune row gather ----------------------------------------------------------------
int s3g_gather_rows(int n, const s3g_row_tensor* tensors, int64_t n_out, int64_t n_kept, const int64_t* src_index,
                    void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n <= 0 || n > ROWS_MAX_TENSORS || !tensors) return fail(S3G_ERR_ARG, "gather_rows: 1..32 tensors expected");
    if (n_out < 0 || n_kept < 0 || n_kept > n_out) return fail(S3G_ERR_ARG, "gather_rows: need 0 <= n_kept <= n_out");
    if (n_out == 0) return S3G_OK;
    if (!src_index) return fail(S3G_ERR_ARG, "gather_rows: null src_index");
    RowArgs a;
    a.count = n; a.n_out = n_out; a.n_kept = n_kept;
    a.src_index = reinterpret_cast<const long long*>(src_index);
    long long most = 0;
    for (int i = 0; i < n; ++i) {
        const s3g_row_tensor& t = tensors[i];
        if (!t.src || !t.dst || t.row_floats <= 0) return fail(S3G_ERR_ARG, "gather_rows: null pointer or row_floats <= 0");
        a.t[i] = RowTensor{t.src, t.dst, t.row_floats, t.zero_new};
        most = std::max(most, (long long)n_out * t.row_floats);
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (most + 255) / 256;
    const int gx = (int)std::min<long long>(want, (long long)sms * 8);
    gather_rows_kernel<<<dim3(gx, n), 256, 0, stream>>>(a);
    S3G_CUDA(cudaGetLastError(), "gather_rows launch");
    return S3G_OK;
}

}  // extern "C"
