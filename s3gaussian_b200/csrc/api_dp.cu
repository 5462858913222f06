// api_dp.cu - extern "C" entry points of the peer-memory gradient all-reduce (see include/s3g_b200.h).
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#include "api_common.cuh"
#include "peer.cuh"

using namespace s3g;

namespace {
inline int fail(int code, const char* what, cudaError_t e = cudaSuccess) { return s3g::api_fail(code, what, e); }

int peer_args(int world, int rank, const void* const* bufs, int64_t numel, PeerArgs& a) {
    if (world < 2 || world > PEER_MAX_RANKS || rank < 0 || rank >= world) return fail(S3G_ERR_ARG, "peer: 2 <= world <= 16, 0 <= rank < world");
    if (!bufs || numel <= 0 || (numel & 3)) return fail(S3G_ERR_ARG, "peer: numel must be a positive multiple of 4");
    for (int p = 0; p < world; ++p) {
        if (!bufs[p] || ((uintptr_t)bufs[p] & 15)) return fail(S3G_ERR_ARG, "peer: null or misaligned buffer pointer");
        a.buf[p] = static_cast<float*>(const_cast<void*>(bufs[p]));
    }
    a.world = world; a.rank = rank;
    a.n4 = numel / 4;
    a.chunk4 = (a.n4 + world - 1) / world;
    return S3G_OK;
}
int peer_grid(long long work4) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (work4 + PEER_THREADS - 1) / PEER_THREADS;
    return (int)std::max<long long>(1, std::min<long long>(want, (long long)sms * 4));
}
}  // namespace

extern "C" {

int s3g_peer_reduce_scatter(int world, int rank, const void* const* bufs, int64_t numel, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PeerArgs a;
    if (int rc = peer_args(world, rank, bufs, numel, a)) return rc;
    const int grid = peer_grid(a.chunk4);
    switch (world) {
        case 2: peer_reduce_scatter_kernel<2><<<grid, PEER_THREADS, 0, stream>>>(a); break;
        case 4: peer_reduce_scatter_kernel<4><<<grid, PEER_THREADS, 0, stream>>>(a); break;
        case 8: peer_reduce_scatter_kernel<8><<<grid, PEER_THREADS, 0, stream>>>(a); break;
        default: peer_reduce_scatter_kernel<0><<<grid, PEER_THREADS, 0, stream>>>(a); break;
    }
    S3G_CUDA(cudaGetLastError(), "peer_reduce_scatter launch");
    return S3G_OK;
}

int s3g_peer_all_gather(int world, int rank, const void* const* bufs, int64_t numel, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PeerArgs a;
    if (int rc = peer_args(world, rank, bufs, numel, a)) return rc;
    const int per = std::max(1, peer_grid(a.chunk4) / (world - 1) * 2);
    peer_all_gather_kernel<<<dim3(per, world - 1), PEER_THREADS, 0, stream>>>(a);
    S3G_CUDA(cudaGetLastError(), "peer_all_gather launch");
    return S3G_OK;
}

int s3g_peer_nvls_all_reduce(int world, int rank, void* multicast_ptr, int64_t numel, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (world < 2 || world > PEER_MAX_RANKS || rank < 0 || rank >= world) return fail(S3G_ERR_ARG, "peer: 2 <= world <= 16, 0 <= rank < world");
    if (!multicast_ptr || ((uintptr_t)multicast_ptr & 15) || numel <= 0 || (numel & 3))
        return fail(S3G_ERR_ARG, "peer_nvls: multicast pointer must be 16-byte aligned, numel a positive multiple of 4");
    const long long n4 = numel / 4, chunk4 = (n4 + world - 1) / world;
    peer_nvls_all_reduce_kernel<<<peer_grid(chunk4), PEER_THREADS, 0, stream>>>(static_cast<float*>(multicast_ptr), rank, n4, chunk4);
    S3G_CUDA(cudaGetLastError(), "peer_nvls_all_reduce launch");
    return S3G_OK;
}

int s3g_peer_reduce_gather(int world, int rank, void* stage_local, const void* const* bucket_ptrs,
                           void* bucket_multicast, int64_t numel, int64_t chunk, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (world < 2 || world > PEER_MAX_RANKS || rank < 0 || rank >= world)
        return fail(S3G_ERR_ARG, "peer: 2 <= world <= 16, 0 <= rank < world");
    if (!stage_local || ((uintptr_t)stage_local & 15) || numel <= 0 || (numel & 3) || chunk <= 0 || (chunk & 3) ||
        chunk * world < numel)
        return fail(S3G_ERR_ARG, "peer_reduce_gather: staging / sizes (numel, chunk multiples of 4, world * chunk >= numel)");
    if (!bucket_multicast && !bucket_ptrs) return fail(S3G_ERR_ARG, "peer_reduce_gather: no destination");
    ReduceGatherArgs a;
    a.stage = static_cast<float*>(stage_local);
    a.mc = static_cast<float*>(bucket_multicast);
    for (int p = 0; p < PEER_MAX_RANKS; ++p) a.bucket[p] = nullptr;
    if (bucket_ptrs)
        for (int p = 0; p < world; ++p) {
            if (!bucket_ptrs[p] || ((uintptr_t)bucket_ptrs[p] & 15)) return fail(S3G_ERR_ARG, "peer_reduce_gather: bucket pointer");
            a.bucket[p] = static_cast<float*>(const_cast<void*>(bucket_ptrs[p]));
        }
    a.world = world; a.rank = rank;
    a.chunk4 = chunk / 4;
    a.n4 = numel / 4;
    const int grid = peer_grid(a.chunk4);
    switch (world) {
        case 2: peer_reduce_gather_kernel<2><<<grid, PEER_THREADS, 0, stream>>>(a); break;
        case 4: peer_reduce_gather_kernel<4><<<grid, PEER_THREADS, 0, stream>>>(a); break;
        case 8: peer_reduce_gather_kernel<8><<<grid, PEER_THREADS, 0, stream>>>(a); break;
        default: peer_reduce_gather_kernel<0><<<grid, PEER_THREADS, 0, stream>>>(a); break;
    }
    S3G_CUDA(cudaGetLastError(), "peer_reduce_gather launch");
    return S3G_OK;
}

}  // extern "C"
