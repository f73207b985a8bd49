// Shared helpers for the volt_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include "../../include/volt_b200.h"
#include <cstdint>
#include <cstdio>

namespace volt {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define VOLT_CUDA(call)                                   \
  do {                                                    \
    int _s = ::volt::check_cuda((call), #call);           \
    if (_s != VOLT_OK) return _s;                 \
  } while (0)

#define VOLT_REQUIRE(cond, ...)                           \
  do {                                                    \
    if (!(cond)) {                                        \
      ::volt::set_error(__VA_ARGS__);                     \
      return VOLT_ERR_ARG;                        \
    }                                                     \
  } while (0)

// Cached workspace, one arena per (device, slot, stream): grown on demand, released by volt_release_workspaces().
int get_workspace(size_t bytes, void** ptr, int slot, cudaStream_t stream, int* created = nullptr);
int sm_count();
int device_slot();   // current CUDA device clamped to [0, 16): index of the per-device caches (function attributes, streams)

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

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  r = red[0];
  __syncthreads();
  return r;
}

}  // namespace volt
