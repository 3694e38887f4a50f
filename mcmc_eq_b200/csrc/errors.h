// errors.h -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
namespace mq {
void set_error(const char* fmt, ...);
}
#define MQ_CUDA(x)                                                                      \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            mq::set_error("%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            return MQ_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)
