// Shared helpers for libzsg_b200 (sm_100a).  No torch types anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/zsg_b200.h"

namespace zsg {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return ZSG_ECUDA;
  }
  return ZSG_OK;
}

#define ZSG_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      zsg::set_error(__VA_ARGS__);        \
      return ZSG_EINVAL;                  \
    }                                     \
  } while (0)

inline cudaStream_t as_stream(zsg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace zsg
