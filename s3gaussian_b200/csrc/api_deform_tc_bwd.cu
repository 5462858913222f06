// api_deform_tc_bwd.cu - launcher of the tcgen05 backward decoder DRAFT (deform_tc_bwd.cuh).  Its own translation
// unit so that the validated kernels of api_deform_bwd.cu stay byte-identical.  NEVER RUN ON HARDWARE; reachable
// only with S3G_TC_BWD=1 in the environment.
#include "deform_host.cuh"
#include "deform_tc_bwd.cuh"

namespace s3g {
namespace {
bool tc_bwd_requested() {
    static const bool on = [] { const char* e = std::getenv("S3G_TC_BWD"); return e && e[0] == '1'; }();
    return on;
}
bool tc_bwd_supported(const DNet& d) {
    return d.L == 4 && d.pos.w1 && d.shs.w1 && d.w_d0 && !d.scl.w1 && !d.rot.w1 && !d.opa.w1;
}
void tc_bwd_table(const DNet& d, BwPrepArgs& p) {
    int t = 0;
    auto add = [&](int e, const float* src, int stride, int col0, bool tr, int n, int k, int n_valid) {
        p.src[e] = src; p.stride[e] = stride; p.col0[e] = col0; p.transpose[e] = tr ? 1 : 0; p.n_valid[e] = n_valid;
        p.tab.off[e] = t; p.tab.n[e] = n; p.tab.k[e] = k;
        t += 2 * n * k;
    };
    add(BW_FA, d.w_feat, 128, 0, false, 64, 64, 64);   add(BW_FB, d.w_feat, 128, 64, false, 64, 64, 64);
    add(BW_D0, d.w_d0, 64, 0, false, 64, 64, 64);      add(BW_D2, d.w_d2, 64, 0, false, 64, 64, 64);
    add(BW_S1, d.shs.w1, 64, 0, false, 64, 64, 64);    add(BW_S2, d.shs.w2, 64, 0, false, 48, 64, 48);
    add(BW_P1, d.pos.w1, 64, 0, false, 64, 64, 64);
    add(BW_D2T, d.w_d2, 64, 0, true, 64, 64, 64);      add(BW_D0T, d.w_d0, 64, 0, true, 64, 64, 64);
    add(BW_S2T, d.shs.w2, 64, 0, true, 64, 48, 64);    // E[n][k] = W_s2[k][n]: 64 x 48
    add(BW_S1T, d.shs.w1, 64, 0, true, 64, 64, 64);    add(BW_P1T, d.pos.w1, 64, 0, true, 64, 64, 64);
    add(BW_FAT, d.w_feat, 128, 0, true, 64, 64, 64);   add(BW_FBT, d.w_feat, 128, 64, true, 64, 64, 64);
    p.tab.total = t;
}
}  // namespace

bool tc_bwd_launch(const DeformBwdArgs& a, int max_grid, cudaStream_t stream, int* grid_out) {
    const DNet& d = a.net;
    if (!tc_bwd_requested() || !tc_bwd_supported(d) || a.P <= 0) return false;
    static_assert(BW_COUNT == 14, "workspace reservation in api_deform_bwd.cu assumes 14 prepared entries");
    BwPrepArgs prep;
    tc_bwd_table(d, prep);
    // prepared weights live behind dfeatures in the caller's workspace (s3g_deform_workspace_bytes reserves the room)
    float* wprep = a.dfeatures + (size_t)a.P * FD * d.L;
    wprep = reinterpret_cast<float*>(((uintptr_t)wprep + 255) & ~(uintptr_t)255);
    prep.dst = wprep;
    bw_prep_weights_kernel<<<dim3(8, BW_COUNT), 256, 0, stream>>>(prep);
    DeformTcBwdArgs t;
    t.net = a.net; t.P = a.P; t.xyz = a.xyz; t.scales = a.scales; t.rot = a.rot; t.opacity = a.opacity; t.shs = a.shs;
    t.campos = a.campos; t.sh_degree = a.sh_degree; t.features = a.features;
    t.g_means = a.g_means; t.g_scales = a.g_scales; t.g_rot = a.g_rot; t.g_opacity = a.g_opacity;
    t.g_colors = a.g_colors; t.g_dx = a.g_dx; t.g_dshs = a.g_dshs; t.g_feat = a.g_feat;
    t.d_scales = a.d_scales; t.d_rot = a.d_rot; t.d_opacity = a.d_opacity; t.d_shs = a.d_shs;
    t.dxyz_direct = a.d_xyz; t.dfeatures = a.dfeatures; t.partial = a.partial; t.off = a.off;
    t.wprep = wprep; t.tab = prep.tab;
    const size_t tsmem = (size_t)(3 * BW_TILE_FLOATS + 2 * 64 * 64) * sizeof(float);
    if (cudaFuncSetAttribute(deform_backward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem) != cudaSuccess) {
        *grid_out = fail(S3G_ERR_CUDA, "deform tc bwd smem attr");
        return true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = (a.P + BWM - 1) / BWM;
    int grid = ntiles < sms ? ntiles : sms;
    if (grid > max_grid) grid = max_grid;
    deform_backward_tc_kernel<<<grid, BWM, tsmem, stream>>>(t);
    *grid_out = cudaGetLastError() == cudaSuccess ? grid : fail(S3G_ERR_CUDA, "deform tc bwd launch");
    return true;
}
}  // namespace s3g
