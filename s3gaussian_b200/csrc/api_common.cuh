// api_common.cuh - what the translation units behind include/s3g_b200.h share: the thread-local error
// message and the CUDA-call check.  (The entry points are split over api.cu, api_deform.cu and
// api_train.cu; one translation unit holding every kernel made cicc 12.9 crash.)
#pragma once
#include <cuda_runtime.h>

#include "../../include/s3g_b200.h"

namespace s3g {
// records `what` (+ the CUDA error string) as the thread's last error and returns `code`
int api_fail(int code, const char* what, cudaError_t e = cudaSuccess);
}  // namespace s3g

#define S3G_CUDA(call, what)                                                  \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return s3g::api_fail(S3G_ERR_CUDA, what, e__); \
    } while (0)
