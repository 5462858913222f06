// api_deform.cu - extern "C" entry points of the fused HexPlane + decoder stage (see include/s3g_b200.h).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>

#include "api_common.cuh"
#include "common.cuh"
#include "deform.cuh"
#include "deform_tc.cuh"

using namespace s3g;

namespace {
inline int fail(int code, const char* what, cudaError_t e = cudaSuccess) { return s3g::api_fail(code, what, e); }
}  // namespace

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

}  // extern "C"
