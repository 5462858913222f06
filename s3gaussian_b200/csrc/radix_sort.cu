// radix_sort.cu - translation unit of the onesweep radix sort (kernels + host driver of radix_sort.cuh).
#define S3G_RADIX_SORT_IMPL
#include <cstdlib>
#include "common.cuh"
#include "radix_sort.cuh"
