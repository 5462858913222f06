// api_deform.cu - extern "C" entry points of the fused HexPlane + decoder stage, forward (see include/s3g_b200.h).
#include "deform_host.cuh"
#include "deform_tc.cuh"

using namespace s3g;

namespace {
// prepared-weight table of the tcgen05 decoder
void build_tc_table(const DNet& d, TcTable& t, TcPrepArgs* prep) {
    int off = 0;
    auto put = [&](int id, const float* W, int N, int K, int ld = 0) {
        if (!W || K <= 0) { t.off[id] = -1; t.npad[id] = 0; t.k[id] = 0; if (prep) { prep->W[id] = nullptr; prep->n[id] = 0; prep->ld[id] = 0; } return; }
        const int np = (N + 15) & ~15;
        t.off[id] = off; t.npad[id] = np; t.k[id] = K;
        off += 2 * np * K;
        if (prep) { prep->W[id] = W; prep->n[id] = N; prep->ld[id] = ld ? ld : K; }
    };
    {   // the feature layer as two K-halves (deform_tc.cuh: TL_FEAT / TL_FEATB)
        const int KF = FD * d.L, KA = KF < 64 ? KF : 64;
        put(TL_FEAT, d.w_feat, 64, KA, KF);
        put(TL_FEATB, d.w_feat ? d.w_feat + KA : nullptr, 64, KF - KA, KF);
    }
    put(TL_POS1, d.pos.w1, 64, 64); put(TL_POS2, d.pos.w1 ? d.pos.w2 : nullptr, 3, 64);
    put(TL_SCL1, d.scl.w1, 64, 64); put(TL_SCL2, d.scl.w1 ? d.scl.w2 : nullptr, 3, 64);
    put(TL_ROT1, d.rot.w1, 64, 64); put(TL_ROT2, d.rot.w1 ? d.rot.w2 : nullptr, 4, 64);
    put(TL_OPA1, d.opa.w1, 64, 64); put(TL_OPA2, d.opa.w1 ? d.opa.w2 : nullptr, 1, 64);
    put(TL_SHS1, d.shs.w1, 64, 64); put(TL_SHS2, d.shs.w1 ? d.shs.w2 : nullptr, 48, 64);
    put(TL_D0, d.w_d0, 64, 64); put(TL_D2, d.w_d0 ? d.w_d2 : nullptr, 64, 64); put(TL_D4, d.w_d0 ? d.w_d4 : nullptr, 3, 64);
    t.total = off;
}
int deform_grid(int ntiles, int ctas_per_sm = 1) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int g = sms * ctas_per_sm;      // persistent CTAs, as many as are resident at once
    return ntiles < g ? ntiles : g;
}
}  // namespace

extern "C" {

int s3g_deform_forward(const s3g_deform_net* net, int P, const float* xyz, const float* scales,
                       const float* rotations, const float* opacity, const float* shs, float time,
                       const float* campos, int sh_degree, float* means3D, float* scales_act,
                       float* rot_act, float* opacity_act, float* colors, float* dx, float* dshs,
                       float* feat, float* features, void* workspace, void* stream_) {
    return s3g_deform_forward_save(net, P, xyz, scales, rotations, opacity, shs, time, campos, sh_degree, means3D, scales_act,
                                   rot_act, opacity_act, colors, dx, dshs, feat, features, nullptr, workspace, stream_);
}

int s3g_deform_forward_save(const s3g_deform_net* net, int P, const float* xyz, const float* scales,
                            const float* rotations, const float* opacity, const float* shs, float time,
                            const float* campos, int sh_degree, float* means3D, float* scales_act,
                            float* rot_act, float* opacity_act, float* colors, float* dx, float* dshs,
                            float* feat, float* features, float* acts, void* workspace, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (P < 0) return fail(S3G_ERR_ARG, "deform_forward: P < 0");
    if (P == 0) return S3G_OK;
    struct { DNet net; } a;
    int rc = to_dnet(net, a.net);
    if (rc != S3G_OK) return rc;
    if (!xyz || !scales || !rotations || !opacity || !shs || !campos)
        return fail(S3G_ERR_ARG, "deform_forward: null input");
    if (!means3D || !scales_act || !rot_act || !opacity_act || !colors || !features)
        return fail(S3G_ERR_ARG, "deform_forward: null output");
    if (sh_degree < 0 || sh_degree > 3) return fail(S3G_ERR_ARG, "deform_forward: sh_degree must be 0..3");
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
    {
        // ---- decoder on the 5th-gen tensor cores (tcgen05 / TMEM) ------------------------------
        if (!workspace) return fail(S3G_ERR_ARG, "deform_forward: null workspace");
        DeformTcArgs t;
        t.net = a.net; t.P = P; t.xyz = xyz; t.scales = scales; t.rot = rotations; t.opacity = opacity; t.shs = shs;
        t.campos = campos; t.sh_degree = sh_degree;
        t.o_means = means3D; t.o_scales = scales_act; t.o_rot = rot_act; t.o_opacity = opacity_act; t.o_colors = colors;
        t.o_dx = dx; t.o_dshs = dshs; t.o_feat = feat; t.features = features;
        t.acts = acts;
        t.act_stride = (size_t)((P + 127) / 128) * 128 * 64;
        {
            int slot[AK_COUNT];
            act_slots(a.net, slot);
            for (int i = 0; i < AK_COUNT; ++i) t.act_slot[i] = slot[i];
        }
        TcPrepArgs pp;
        build_tc_table(a.net, t.tab, &pp);
        pp.tab = t.tab;
        pp.dst = reinterpret_cast<float*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
        t.wprep = pp.dst;
        tc_prep_weights_kernel<<<dim3(8, TL_COUNT), 256, 0, stream>>>(pp);
        S3G_CUDA(cudaGetLastError(), "tc_prep_weights launch");
        // operands [128][64] hi + lo, weights [64][64] hi + lo: 96 KB, two CTAs (and 2 x 256 TMEM columns) per SM
        const size_t smem = (size_t)(2 * TCM * 64 + 2 * 64 * 64) * sizeof(float);
        S3G_CUDA(cudaFuncSetAttribute(deform_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "deform tc smem attr");
        const int ntiles = (P + TCM - 1) / TCM;
        deform_forward_tc_kernel<<<deform_grid(ntiles, 2), TCM, smem, stream>>>(t);
    }
    S3G_CUDA(cudaGetLastError(), "deform_forward launch");
    return S3G_OK;
}


size_t s3g_deform_forward_workspace_bytes(const s3g_deform_net* net) {
    DNet d;
    if (to_dnet(net, d) != S3G_OK) return 0;
    TcTable t;
    build_tc_table(d, t, nullptr);
    return (size_t)t.total * sizeof(float) + 512;
}

}  // extern "C"
