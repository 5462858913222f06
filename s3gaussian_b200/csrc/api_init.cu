// api_init.cu - extern "C" entry points of the one-time initialisation kernels (see include/s3g_b200.h).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "api_common.cuh"
#include "common.cuh"
#include "knn.cuh"
#include "radix_sort.cuh"

using namespace s3g;

namespace {
inline int fail(int code, const char* what, cudaError_t e = cudaSuccess) { return s3g::api_fail(code, what, e); }

struct KnnWork {
    uint32_t* bounds;
    uint32_t *keys_a, *vals_a, *keys_b, *vals_b, *order;
    float4* sorted;
    KnnBox *boxes, *super;
    SortTemp st;
    int nboxes, nsuper;
};
KnnWork knn_carve(Carver& c, int P) {
    KnnWork w;
    w.nboxes = (P + KNN_BOX - 1) / KNN_BOX;
    w.nsuper = (w.nboxes + KNN_SUPER - 1) / KNN_SUPER;
    w.bounds = c.take<uint32_t>(8);
    w.keys_a = c.take<uint32_t>((size_t)P); w.vals_a = c.take<uint32_t>((size_t)P);
    w.keys_b = c.take<uint32_t>((size_t)P); w.vals_b = c.take<uint32_t>((size_t)P);
    w.order = c.take<uint32_t>((size_t)P);
    w.sorted = c.take<float4>((size_t)P);
    w.boxes = c.take<KnnBox>((size_t)w.nboxes);
    w.super = c.take<KnnBox>((size_t)w.nsuper);
    w.st = SortTemp::carve(c, P);
    return w;
}
}  // namespace

extern "C" {

size_t s3g_knn_workspace_bytes(int P) {
    if (P <= 0) return 0;
    char* z = nullptr;
    Carver c(z);
    knn_carve(c, P);
    return c.off + 256;
}

int s3g_knn_mean_dist2(int P, const float* points, float* mean_dist2, void* workspace, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (P < 0) return fail(S3G_ERR_ARG, "knn: P < 0");
    if (P == 0) return S3G_OK;
    if (!points || !mean_dist2 || !workspace) return fail(S3G_ERR_ARG, "knn: null pointer");
    Carver c(static_cast<char*>(workspace));
    KnnWork w = knn_carve(c, P);
    S3G_CUDA(cudaMemsetAsync(w.bounds, 0xff, 3 * sizeof(uint32_t), stream), "knn memset");
    S3G_CUDA(cudaMemsetAsync(w.bounds + 3, 0, 3 * sizeof(uint32_t), stream), "knn memset");
    const int grid = (P + 255) / 256;
    knn_bounds_kernel<<<grid < 1184 ? grid : 1184, 256, 0, stream>>>(P, points, w.bounds);
    knn_morton_kernel<<<grid, 256, 0, stream>>>(P, points, w.bounds, w.keys_a, w.vals_a);
    S3G_CUDA(cudaGetLastError(), "knn morton launch");
    S3G_CUDA(radix_sort_pairs((uint32_t)P, w.keys_a, w.vals_a, w.keys_b, w.vals_b, nullptr, w.order, 0, 30, w.st, stream),
             "knn sort");
    knn_box_kernel<<<w.nboxes, KNN_BOX, 0, stream>>>(P, points, w.order, w.sorted, w.boxes);
    knn_super_kernel<<<w.nsuper, 32, 0, stream>>>(w.nboxes, w.boxes, w.super);
    knn_search_kernel<<<grid, 256, 0, stream>>>(P, w.sorted, w.boxes, w.nboxes, w.super, w.nsuper, mean_dist2);
    S3G_CUDA(cudaGetLastError(), "knn search launch");
    return S3G_OK;
}

}  // extern "C"
