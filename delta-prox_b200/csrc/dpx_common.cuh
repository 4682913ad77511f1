// dpx_common.cuh — shared helpers for libdprox_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dprox_b200.h"
#include "dpx_types.cuh"

namespace dpx {

// ---- error plumbing (thread-local message, int status across the ABI) ------------------------
void set_error(const char* fmt, ...);

#define DPX_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      dpx::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DPX_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define DPX_CUFFT(call)                                                                       \
  do {                                                                                        \
    cufftResult r__ = (call);                                                                 \
    if (r__ != CUFFT_SUCCESS) {                                                               \
      dpx::set_error("%s failed: cufftResult %d (%s:%d)", #call, (int)r__, __FILE__, __LINE__);   \
      return DPX_ERR_CUFFT;                                                                   \
    }                                                                                         \
  } while (0)

#define DPX_REQUIRE(cond, ...)                                                                \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      dpx::set_error(__VA_ARGS__);                                                            \
      return DPX_ERR_INVALID;                                                                 \
    }                                                                                         \
  } while (0)

extern unsigned long long g_launches;   // number of library kernels launched (bench.py's gpu_launches)

#define DPX_LAUNCH_CHECK()                                                                    \
  do {                                                                                        \
    ++dpx::g_launches;                                                                        \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess) {                                                                 \
      dpx::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DPX_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- device helpers -----------------------------------------------------------------------------
#ifdef __CUDACC__

// Streaming 128-bit accesses: data touched once per kernel, keep it out of L1.
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of K values; result valid in thread 0.  `red` is K*32 floats of shared memory.
template <int K>
__device__ __forceinline__ void block_sum(float (&vals)[K], float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float s = warp_sum(vals[k]);
    if (lane == 0) red[k * 32 + warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float s = lane < nwarp ? red[k * 32 + lane] : 0.f;
      vals[k] = warp_sum(s);
    }
  }
}

#endif  // __CUDACC__

}  // namespace dpx
