// common.cuh -- shared helpers for libiea_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/iea_b200.h"

namespace iea {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define IEA_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      iea::set_error(__VA_ARGS__);               \
      return -2;                                 \
    }                                            \
  } while (0)

#define IEA_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      iea::set_error("%s failed: %s", #call, cudaGetErrorString(e_));         \
      return -3;                                                              \
    }                                                                         \
  } while (0)

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float ld_act(const void* p, int dtype, int64_t i) {
  return dtype == IEA_BF16 ? __bfloat162float(((const bf16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void st_act(void* p, int dtype, int64_t i, float v) {
  if (dtype == IEA_BF16) ((bf16*)p)[i] = __float2bfloat16_rn(v);
  else ((float*)p)[i] = v;
}
// value as the consumer will read it back (rounded through the storage type)
__device__ __forceinline__ float round_act(int dtype, float v) {
  return dtype == IEA_BF16 ? __bfloat162float(__float2bfloat16_rn(v)) : v;
}
static inline size_t dtype_size(int dtype) { return dtype == IEA_BF16 ? 2 : 4; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum, result valid in every thread; blockDim.x multiple of 32, <= 1024.
// `red` is shared scratch of >= 33 floats.  Deterministic (fixed tree).
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < nw ? red[lane] : -3.4e38f;
    t = warp_max(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace iea
